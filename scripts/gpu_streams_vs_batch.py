"""P ring pairs on one GPU: ONE lock-step batch (one launch per iteration half for all pairs) against G concurrent groups, each a
lock-step batch on its own stream (development aid): python scripts/gpu_streams_vs_batch.py P G [G ...]"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvr_b200, mvr_b200.synth as synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Gs = [int(x) for x in sys.argv[2:]] or [1, P]
V, n = 24, 200_000
views, poses = zip(*[synth.turntable_view(v % V, V, n) for v in range(P + 1)])
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(P + 1)]
dv = [torch.from_numpy(p).cuda() for p in views]
ctxs = [mvr_b200.Context(0) for _ in range(P)]
guesses = []
for k, c in enumerate(ctxs):
    c.set_target_device(dv[k].data_ptr(), n); c.set_source_device(dv[k + 1].data_ptr(), n)
    guesses.append((np.linalg.inv(init[k]) @ init[k + 1]).astype(np.float32))
prm = mvr_b200.default_params(max_iterations=30, max_dist=4.0, reciprocal=1, fixed_iterations=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run_groups(G):
    out = [None] * P
    bounds = [(g * P) // G for g in range(G + 1)]
    def run(g):
        a, b = bounds[g], bounds[g + 1]
        res = mvr_b200.icp_align_batch(ctxs[a:b], prm, guesses[a:b])
        out[a:b] = res
    th = [threading.Thread(target=run, args=(g,)) for g in range(1, G)]
    for t in th: t.start()
    run(0)
    for t in th: t.join()
    return out
ref = None
for G in Gs + Gs:
    ts = []
    for rep in range(5):
        flush.fill_(1); torch.cuda.synchronize(); t0 = time.perf_counter(); r = run_groups(G); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    sig = [(x["n_corr"], x["final"].tobytes()) for x in r]
    if ref is None: ref = sig
    print("%2d pairs in %2d concurrent group(s): %.3f ms (min of 5)  identical results: %s" % (P, G, 1e3 * min(ts), sig == ref), flush=True)
