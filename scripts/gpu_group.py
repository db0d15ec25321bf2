"""Batched ring registration vs pairs per launch (development aid): python scripts/gpu_group.py [views] [n]"""
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
icp = mvr_b200.default_params(max_iterations=30, max_dist=4.0, reciprocal=1, fixed_iterations=1)
tp = mvr_b200.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr_b200.RING_PAIRS, loop_closure=1, lum_iterations=16)
reg = mvr_b200.Registrator(0, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
groups = [int(x) for x in os.environ.get("MVR_GROUPS", "1,2,3,4,6,8").split(",")]
cells = [float(x) for x in os.environ.get("MVR_CELLS", "0").split(",")]
reg.register_turntable(dl, tp, init_poses=init)   # creates the contexts
for cell in cells:
  for k in range(reg.streams()):
    reg.context(k).set_index_options(cell, 9)
  for g in groups:
    reg.context(0).set_batch_group(g)
    ts = []
    for rep in range(4):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); got, reps = reg.register_turntable(dl, tp, init_poses=init); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("cell %.2f group %d: registration %.2f ms (min of 5; all %s) pair0 ncorr %d gpu_ms %.2f" % (cell, g, 1e3 * min(ts), ["%.1f" % (1e3 * t) for t in ts], reps[0]["n_corr"], reps[0]["gpu_ms"]), flush=True)
