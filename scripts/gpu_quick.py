"""Quick GPU timing probe (development aid, not the bench)."""
import sys, time, os
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
for recip in (1, 0):
    p = mvr_b200.default_params(max_iterations=30, max_dist=4.0, reciprocal=recip, fixed_iterations=1)
    for rep in range(3):
        ctx.kernel_stats(reset=True)
        t0 = time.time(); r = ctx.icp_align(p, guess=guess, n_source=n); t1 = time.time()
        st = ctx.kernel_stats(reset=True)
    print("ICP n=%d recip=%d: wall %.2f ms gpu %.2f ms iters %d ncorr %d mse %.4f queries %d" % (n, recip, (t1-t0)*1e3, r["gpu_ms"], r["iterations"], r["n_corr"], r["mse"], r["nn_queries"]))
    print("   serial solve: %.0f cycles/iteration (%d solves)" % (ctx.debug_value(0) / max(ctx.debug_value(1), 1), ctx.debug_value(1)))
    for k, v in st.items():
        if v["launches"]: print("   %-10s launches %4d  total %.3f ms  avg %.1f us  GB/s(alg) %.1f" % (k, v["launches"], v["ms"], 1e3*v["ms"]/v["launches"], v["bytes"]/max(v["ms"],1e-9)/1e6))
t0 = time.time(); f = ctx.fitness_score(); print("fitness %.4f in %.2f ms" % (f, (time.time()-t0)*1e3))
# NN sweep
import torch
for m, nq in ((1_000_000, 16_000_000),):
    tgt, q = synth.nn_sweep_case(m, nq, order="random")
    for order in ("random", "morton"):
        if order == "morton":
            _, q = synth.nn_sweep_case(m, nq, order="morton")
        tq = torch.from_numpy(q).cuda(); ti = torch.empty(nq, dtype=torch.int32, device="cuda"); td = torch.empty(nq, dtype=torch.float32, device="cuda")
        ctx.set_target(tgt)
        for rep in range(3):
            ctx.kernel_stats(reset=True)
            ctx.nn_query_device(tq.data_ptr(), nq, ti.data_ptr(), td.data_ptr()); ctx.synchronize()
            st = ctx.kernel_stats(reset=True)
        v = st["nn"]
        print("NN m=%d n=%d %s: %.3f ms  %.2f Gq/s  alg GB/s %.1f ; index: morton %.3f sort %.3f table %.3f ms" % (m, nq, order, v["ms"], nq/v["ms"]/1e6, v["bytes"]/v["ms"]/1e6, st["morton"]["ms"], st["sort"]["ms"], st["table"]["ms"]))
