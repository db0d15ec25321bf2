"""Registration time of the 24 x 200k ring against the iteration count (development aid): where in the align the time goes.
python scripts/gpu_iter_profile.py [views] [n]   (MVR_B200_LIB selects a library variant)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvr_b200, mvr_b200.synth as synth
V = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
dv = [torch.from_numpy(p).cuda() for p in views]
dl = [(t.data_ptr(), n) for t in dv]
reg = mvr_b200.Registrator(0, 1)
if os.environ.get("MVR_GROUP"):   # pairs per launch (0 / unset: automatic, a quarter of the batch; the groups run concurrently)
    reg.context(0).set_batch_group(int(os.environ["MVR_GROUP"]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prev = 0.0
for iters in [int(x) for x in os.environ.get("MVR_ITERS", "1,2,3,4,6,8,12,16,20,30").split(",")]:
    icp = mvr_b200.default_params(max_iterations=iters, max_dist=4.0, reciprocal=int(os.environ.get("MVR_RECIP", "1")), fixed_iterations=1)
    tp = mvr_b200.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr_b200.RING_PAIRS, loop_closure=1, lum_iterations=16,
                                   pair_begin=0, pair_end=int(os.environ.get("MVR_PAIRS", "0")))
    ts = []
    for rep in range(4):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); got, reps = reg.register_turntable(dl, tp, init_poses=init); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = 1e3 * min(ts)
    c = reg.context(0)
    dbg = [c.debug_value(k) for k in range(8)]
    print("iters %2d: %.3f ms (+%.3f)  pair0 ncorr %d  missed %d" % (iters, t, t - prev, reps[0]["n_corr"], dbg[2]), flush=True)
    prev = t
