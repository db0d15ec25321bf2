"""NN sweep quick probe over ppc (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvr_b200, mvr_b200.synth as synth
ctx = mvr_b200.Context(0); ctx.set_profiling(True)
m = 1_000_000
for nq in (1_000_000, 16_000_000):
    tgt, q = synth.nn_sweep_case(m, nq, order="random")
    tq = torch.from_numpy(q).cuda(); ti = torch.empty(nq, dtype=torch.int32, device="cuda"); td = torch.empty(nq, dtype=torch.float32, device="cuda")
    for ppc, ratio in ((6, 0.0), (2, 1e-9), (3, 1e-9), (4, 1e-9), (6, 1e-9), (8, 1e-9), (12, 1e-9), (16, 1e-9)):
        ctx.set_nn_options(ppc, ratio)
        ctx.set_target(tgt)
        for rep in range(3):
            ctx.kernel_stats(reset=True)
            ctx.nn_query_device(tq.data_ptr(), nq, ti.data_ptr(), td.data_ptr()); ctx.synchronize()
            st = ctx.kernel_stats(reset=True)
        print("n=%d ppc=%g %s: nn %.3f ms (%.2f Gq/s)  sort %.3f ms table %.3f" % (nq, ppc, "dense" if ratio else "brick", st["nn"]["ms"], nq / st["nn"]["ms"] / 1e6, st["sort"]["ms"], st["table"]["ms"]), flush=True)
        if not ratio: ref = (ti.clone(), td.clone())
        else: assert torch.equal(ti, ref[0]) and torch.equal(td.view(torch.int32), ref[1].view(torch.int32)), "dense and brick results differ"
    del tq, ti, td
