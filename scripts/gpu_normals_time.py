import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
tgt, _ = synth.turntable_view(0, 24, n)
c = mvr_b200.Context(0); c.set_target(tgt); c.set_profiling(True)
for rep in range(3):
    c.kernel_stats(reset=True)
    t0 = time.perf_counter(); nrm = c.estimate_normals(mvr_b200.TARGET, n, 16); t1 = time.perf_counter()
    st = c.kernel_stats(reset=True)
    print("normals %d pts k=16: kernel %.3f ms, index %.3f ms, call %.1f ms" % (n, st["normals"]["ms"], st["sort"]["ms"], 1e3 * (t1 - t0)), flush=True)
