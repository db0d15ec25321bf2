"""Cell-edge sweep of the fused ICP iteration (development aid): python scripts/gpu_cellsweep.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth
ctx = mvr_b200.Context(0)
ctx.set_profiling(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
tgt, _ = synth.turntable_view(0, 24, n)
src, Ts = synth.turntable_view(1, 24, n)
guess = (synth.perturbation() @ Ts).astype(np.float32)
ctx.set_target(tgt); ctx.set_source(src)
for gate in (4.0,):
    for cell in (0.0, 0.75, 1.0, 1.25, 1.5, 2.0, 2.5, 3.0):
        ctx.set_index_options(cell, 9)
        for recip in (1, 0):
            p = mvr_b200.default_params(max_iterations=30, max_dist=gate, reciprocal=recip, fixed_iterations=1)
            for rep in range(3):
                ctx.kernel_stats(reset=True); r = ctx.icp_align(p, guess=guess, n_source=n); st = ctx.kernel_stats(reset=True)
            print("gate %.1f cell %.2f recip %d: gpu %.2f ms  corr %.1f us/iter  index build %.1f us  ncorr %d" % (
                gate, cell, recip, r["gpu_ms"], 1e3 * st["corr"]["ms"] / 30, 1e3 * st["sort"]["ms"], r["n_corr"]), flush=True)
