"""Lane imbalance of the flat candidate scan, simulated on the CPU (development aid; scipy kd-tree, no GPU): for the settled
bench pair, per-lane and per-warp trips / rows of the ball walk, and what running K queries of a lane as one loop would save."""
import sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import importlib
synth = importlib.import_module('multi-view-registration_b200.synth')
from scipy.spatial import cKDTree
V=24; n=200000
t,Tt = synth.turntable_view(0,V,n)
s,Ts = synth.turntable_view(1,V,n)
P = np.linalg.inv(Tt) @ Ts
sp = (s[:,:3].astype(np.float64) @ P[:3,:3].T + P[:3,3])
tp = t[:,:3].astype(np.float64)
tree = cKDTree(tp)
u = np.array([0.577,0.577,0.577])
r1,r2 = 0.3,0.27
q1 = sp + u*r1; q = sp + u*r2
d1,j1 = tree.query(q1)
gate=4.0
matched = d1 < gate
d = np.where(matched, np.linalg.norm(q - tp[j1],axis=1), gate)
d = np.minimum(d, gate)
c = np.array((1.0,2.0,2.0))
lo = tp.min(0) - 2
dims = np.ceil((tp.max(0)+2-lo)/c).astype(int)+1
ti = np.floor((tp-lo)/c).astype(np.int64)
key = (ti[:,2]*dims[1]+ti[:,1])*dims[0]+ti[:,0]
H = np.bincount(key, minlength=int(dims.prod()))
cs = np.concatenate([[0],np.cumsum(H)])
# source sorted by its own cells (grid over source box, same cell shape)
slo = q.min(0)-2
sdims = np.ceil((q.max(0)+2-slo)/c).astype(int)+1
si = np.floor((q-slo)/c).astype(np.int64)
skey = (si[:,2]*sdims[1]+si[:,1])*sdims[0]+si[:,0]
order = np.argsort(skey, kind='stable')
m=0.02
N = 32*2000
sel = order[40000:40000+N]
def clampi(v,a): return np.minimum(np.maximum(v,0),dims[a]-1)
trips = np.zeros(N,int); cands=np.zeros(N,int); rows=np.zeros(N,int); nseg=np.zeros(N,int)
for ii,i in enumerate(sel):
    qq=q[i]; dd=d[i]
    tq = (qq-lo)/c
    rc = dd/c
    a = [clampi(int(np.floor(tq[k]-rc[k]-m)),k) for k in range(3)]
    b = [clampi(int(np.floor(tq[k]+rc[k]+m)),k) for k in range(3)]
    for z in range(a[2],b[2]+1):
        ez = max(0, max(z-tq[2], tq[2]-(z+1)) - m) if not (z<=tq[2]<z+1) else 0
        for y in range(a[1],b[1]+1):
            ey = max(0, max(y-tq[1], tq[1]-(y+1)) - m) if not (y<=tq[1]<y+1) else 0
            eyz = (ez*2.0)**2+(ey*2.0)**2
            rows[ii]+=1
            if eyz > dd*dd: continue
            rx = np.sqrt(max(dd*dd-eyz,0))/c[0] + 2*m
            xa = max(a[0], clampi(int(np.floor(tq[0]-rx)),0)); xb=min(b[0], clampi(int(np.floor(tq[0]+rx)),0))
            r = (z*dims[1]+y)*dims[0]
            k = cs[r+xb+1]-cs[r+xa]
            if k>0:
                cands[ii]+=k; trips[ii]+= (k+7)//8; nseg[ii]+=1
tw = trips.reshape(-1,32)
print("matched frac", matched[sel].mean())
print("per-lane: cands mean %.1f  trips mean %.2f  rows %.2f segs %.2f"%(cands.mean(), trips.mean(), rows.mean(), nseg.mean()))
print("per-warp max trips mean %.2f ; mean-of-lane-mean %.2f"%(tw.max(1).mean(), tw.mean()))
print("percentiles cands", np.percentile(cands,[50,90,99,99.9]))
print("percentiles trips", np.percentile(trips,[50,90,99,99.9]))
print("warp max trips hist", np.bincount(tw.max(1))[:40])
rw = rows.reshape(-1,32)
print("per-warp max rows mean %.2f"%rw.max(1).mean())
# unmatched
um = ~matched[sel]
print("unmatched lanes: cands mean %.1f trips %.1f rows %.1f"%(cands[um].mean(), trips[um].mean(), rows[um].mean()))
mm = matched[sel]
twm = np.where(mm.reshape(-1,32), tw, 0)
print("warp max trips over matched lanes only: %.2f"%twm.max(1).mean())
print("---- concatenation")
for K in (1,2,4):
    W = (N//32)//K
    t3 = trips[:W*K*32].reshape(K, W, 32)
    r3 = rows[:W*K*32].reshape(K, W, 32)
    seq = t3.max(2).sum(0).mean()      # K sequential searches: sum of per-item warp maxima
    cat = t3.sum(0).max(1).mean()      # one loop over the K lists of a lane
    rseq = r3.max(2).sum(0).mean(); rcat = r3.sum(0).max(1).mean()
    print("K=%d trips: sequential %.2f  concatenated %.2f (%.0f%%)   rows: %.2f -> %.2f (%.0f%%)"%(K, seq, cat, 100*(cat/seq-1), rseq, rcat, 100*(rcat/rseq-1)))
