import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, mvr_b200 as mvr, mvr_b200.synth as synth, oracle as orc
orc.build()
V, n = 6, 8000
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
ctx = mvr.Context(0)
p = mvr.default_params(max_iterations=6, max_dist=4.0, reciprocal=1, fixed_iterations=1)
op = orc.make_params(max_iterations=6, max_dist=4.0, reciprocal=True, fixed_iterations=True)
for variant in ("target once + fitness", "target once", "target each"):
    print("==", variant)
    model = orc.apply_pose_double(views[0], init[0])
    for v in (1, 2, 3):
        src = orc.apply_pose_double(views[v], init[v])
        if variant != "target each": ctx.set_target(model)
        for rep in range(2):
            if variant == "target each": ctx.set_target(model)
            ctx.set_source(src)
            r = ctx.icp_align(p, n_source=len(src), want_cloud=True)
            o = orc.icp_align(src, model, op)
            print("view", v, "rep", rep, "ncorr equal", [a["n_corr"] for a in r["log"]] == [a["n_corr"] for a in o["log"]], "cloud equal", np.array_equal(r["cloud"], o["cloud"]), "final diff", np.abs(r["final"]-o["final"]).max())
            src = o["cloud"]
        if variant == "target once + fitness": print("   fitness", ctx.fitness_score(), orc.fitness_score(src, model))
        model = np.concatenate([model, src])
