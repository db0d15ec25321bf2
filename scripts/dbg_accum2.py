import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, mvr_b200 as mvr, mvr_b200.synth as synth, oracle as orc
orc.build()
V, n = 6, 8000
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
icp = mvr.default_params(max_iterations=6, max_dist=4.0, reciprocal=1, fixed_iterations=1)
op = orc.make_params(max_iterations=6, max_dist=4.0, reciprocal=True, fixed_iterations=True)
for reps_n in (1, 2):
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=reps_n, mode=mvr.ACCUMULATE)
    reg = mvr.Registrator(0, 1)
    got, reps = reg.register_turntable(views[:4], tp, init_poses=init[:4])
    pose = [np.array(T, dtype=np.float64) for T in init[:4]]
    model = orc.apply_pose_double(views[0], pose[0])
    for v in (1, 2, 3):
        src = orc.apply_pose_double(views[v], pose[v])
        its = 0
        for _ in range(reps_n):
            o = orc.icp_align(src, model, op); pose[v] = o["final"].astype(np.float64) @ pose[v]; src = o["cloud"]; its += o["iterations"]
        model = np.concatenate([model, src])
        r = reps[v - 1]
        print("repeats", reps_n, "view", v, "driver ncorr/mse/iters", r["n_corr"], r["mse"], r["iterations"], "oracle", o["n_corr"], o["mse"], its,
              "pose diff", np.abs(got[v] - pose[v]).max(), "fitness", r["fitness"], orc.fitness_score(src, model[:-len(src)]))
    reg.close()
