"""Cell-edge sweep of the ICP correspondence kernel (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth
ctx = mvr_b200.Context(0)
ctx.set_profiling(True)
n = 200_000
tgt, _ = synth.turntable_view(0, 24, n)
src, Ts = synth.turntable_view(1, 24, n)
guess = (synth.perturbation() @ Ts).astype(np.float32)
ctx.set_target(tgt); ctx.set_source(src)
for gate in (2.0, 4.0):
    for cell in (0.0, 1.0, 1.25, 1.5, 2.0, 2.5):
        ctx.set_index_options(cell, 9)
        for recip in (1, 0):
            p = mvr_b200.default_params(max_iterations=30, max_dist=gate, reciprocal=recip, fixed_iterations=1)
            for rep in range(2):
                ctx.kernel_stats(reset=True); r = ctx.icp_align(p, guess=guess, n_source=n); st = ctx.kernel_stats(reset=True)
            print("gate %.1f cell %.2f recip %d: gpu %.2f ms  corr %.1f us  table %.1f us  sort %.1f us reduce %.1f ncorr %d" % (
                gate, cell, recip, r["gpu_ms"], 1e3*st["corr"]["ms"]/30, 1e3*st["table"]["ms"]/30, 1e3*st["sort"]["ms"]/30, 1e3*st["reduce"]["ms"]/30, r["n_corr"]))
