"""Where the reference-style accumulative registration spends its time (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth
V, n = 24, 200_000
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
reg = mvr_b200.Registrator(0, 1)
icp = mvr_b200.default_params(max_dist=4.0, reciprocal=1, max_iterations=2**31 - 1, euclidean_fitness_epsilon=50.0)
tp = mvr_b200.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=5, mode=mvr_b200.ACCUMULATE, want_fitness=0)
for rep in range(3):
    c = reg.context(0); c.set_profiling(True); c.kernel_stats(reset=True)
    t0 = time.perf_counter(); got, reps = reg.register_turntable(views, tp, init_poses=init); wall = 1e3 * (time.perf_counter() - t0)
    st = c.kernel_stats(reset=True)
    print("wall %.1f ms; kernel ms:" % wall, {k: (v["launches"], round(v["ms"], 2)) for k, v in st.items() if v["launches"]}, flush=True)
