import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mvr_b200 as mvr, mvr_b200.synth as synth
def rot_angle(A, B):
    R = np.asarray(A, dtype=np.float64)[:3, :3] @ np.asarray(B, dtype=np.float64)[:3, :3].T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))
V, n = 6, 8000
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
P = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
reg = mvr.Registrator(0, 1)
one = mvr.default_params(max_iterations=1, max_dist=4.0, reciprocal=1, fixed_iterations=1)
c = np.array([0, 0, 900.0])
for loop in range(8):
    rel, w = [], []
    for i in range(V):
        s, t = i, (i + 1) % V
        guess = (np.linalg.inv(P[t]) @ P[s]).astype(np.float32)
        r = reg.pairwise_align(views[s], views[t], one, guess=guess)
        Z = r["final"].astype(np.float64) @ np.linalg.inv(guess.astype(np.float64))
        Zw = P[t] @ Z @ np.linalg.inv(P[t])
        # displacement of the object centre (world) by Zw
        cw = P[t] @ np.append(c, 1.0)
        if loop < 0: print("loop", loop, "edge", i, "ncorr", r["n_corr"], "mse", round(r["mse"], 3), "Z rot", round(rot_angle(Z, np.eye(4)), 5), "centre shift", round(float(np.linalg.norm((Zw @ cw - cw)[:3])), 3))
        rel.append(np.linalg.inv(Zw)); w.append(r["n_corr"])
    X = mvr.ring_close(rel, w, relax=True, iterations=16, centre=(P[0] @ np.append(c, 1.0))[:3], rot_scale=100.0)
    print("  X rot", [round(rot_angle(x, np.eye(4)), 4) for x in X])
    P = [X[v].astype(np.float64) @ P[v] for v in range(V)]
    print("  err vs truth", [round(rot_angle(np.linalg.inv(P[0]) @ P[v], np.linalg.inv(poses[0]) @ poses[v]), 4) for v in range(V)])
