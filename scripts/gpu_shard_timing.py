"""Where a 3-pair shard (the 8-GPU case) spends its time (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvr_b200, mvr_b200.synth as synth, mvr_b200.ring as ring
V, n = 24, 200_000
need = [0, 1, 2, 3]
views = {v: synth.turntable_view(v, V, n)[0] for v in need}
poses = [synth.view_pose(v, V) for v in range(V)]
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
dv = {v: torch.from_numpy(p).cuda() for v, p in views.items()}
dl = [(dv[v].data_ptr(), n) if v in dv else None for v in range(V)]
icp = mvr_b200.default_params(max_iterations=30, max_dist=4.0, reciprocal=1, fixed_iterations=1)
tp = mvr_b200.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr_b200.RING_PAIRS, loop_closure=1, lum_iterations=16, pair_begin=0, pair_end=3)
reg = mvr_b200.Registrator(0, 1)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got, reps = reg.register_turntable(dl, tp, init_poses=init)
    t1 = time.perf_counter()
    rec = ring.pack_reports(reps, 0, 3)
    allrec = np.concatenate([rec] * 8, axis=0)   # stands in for the gather (no second rank here)
    t2 = time.perf_counter()
    ring.close_ring(allrec, synth.PIVOT, 115.0)
    t3 = time.perf_counter()
    print("register_turntable %.3f ms (gpu_ms %.3f)  pack+gather %.3f ms  close_ring %.3f ms" % (1e3 * (t1 - t0), reps[0]["gpu_ms"], 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
