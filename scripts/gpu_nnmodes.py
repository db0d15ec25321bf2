"""NN kernels against batch size (development aid): python scripts/gpu_nnmodes.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvr_b200, mvr_b200.synth as synth
ctx = mvr_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
m = 1_000_000
tgt, q = synth.nn_sweep_case(m, 16_000_000, order="random")
_, qm = synth.nn_sweep_case(m, 16_000_000, order="morton")
ctx.set_target(tgt)
for order, qq in (("random", q), ("morton", qm)):
    tq = torch.from_numpy(qq).cuda(); ti = torch.empty(len(qq), dtype=torch.int32, device="cuda"); td = torch.empty(len(qq), dtype=torch.float32, device="cuda")
    for nq in [int(x) for x in os.environ.get("MVR_NQ", "10000,100000,1000000,4000000,16000000").split(",")]:
        ref = None
        for mode, ppc in (("WARP", 8), ("THREAD", 8), ("CELL", 8), ("SEEDED", 16), ("SEEDED", 8), ("SEEDED", 4), ("SEEDED", 2)):
            if mode == "WARP" and nq > 1_000_000: continue
            ctx.set_nn_mode(getattr(mvr_b200, "NN_" + mode)); ctx.set_nn_options(ppc, 8.0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            best = 1e9
            for rep in range(4):
                e0.record(); ctx.nn_query_device(tq.data_ptr(), nq, ti.data_ptr(), td.data_ptr()); e1.record(); e1.synchronize()
                if rep: best = min(best, e0.elapsed_time(e1))
            if ref is None: ref = (ti[:nq].clone(), td[:nq].clone())
            else: assert torch.equal(ti[:nq], ref[0]) and torch.equal(td[:nq].view(torch.int32), ref[1].view(torch.int32)), "kernels disagree"
            print("%s n=%8d %-6s ppc=%d: %.3f ms  %.2f Gq/s" % (order, nq, mode, ppc, best, nq / best / 1e6), flush=True)
    del tq, ti, td
