"""One pairwise ICP (development aid for ncu captures): python scripts/prof_icp.py [n] [iters] [recip]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
recip = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = mvr_b200.Context(0)
tgt, _ = synth.turntable_view(0, 24, n)
src, Ts = synth.turntable_view(1, 24, n)
guess = (synth.perturbation() @ Ts).astype(np.float32)
ctx.set_target(tgt); ctx.set_source(src)
p = mvr_b200.default_params(max_iterations=iters, max_dist=4.0, reciprocal=recip, fixed_iterations=1)
for rep in range(2):
    r = ctx.icp_align(p, guess=guess, n_source=n)
print("iters %d ncorr %d mse %.5f gpu_ms %.3f" % (r["iterations"], r["n_corr"], r["mse"], r["gpu_ms"]))
