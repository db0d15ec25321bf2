"""registrationLUM on the bench sequence (development aid): python scripts/gpu_lum.py [views] [n] [outer loops]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mvr_b200, mvr_b200.synth as synth
V = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
loops = int(sys.argv[3]) if len(sys.argv) > 3 else 4
def rot_angle(A, B):
    R = np.asarray(A, dtype=np.float64)[:3, :3] @ np.asarray(B, dtype=np.float64)[:3, :3].T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))
views, poses = synth.turntable_sequence(V, n)
E = synth.perturbation()
init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
icp = mvr_b200.default_params(max_iterations=16 * loops, max_dist=4.0)
tp = mvr_b200.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, mode=mvr_b200.LUM)
reg = mvr_b200.Registrator(0, 1)
for rep in range(3):
    t0 = time.perf_counter(); got, _ = reg.register_turntable(views, tp, init_poses=init); dt = time.perf_counter() - t0
    err = max(rot_angle(np.linalg.inv(got[0]) @ got[v], np.linalg.inv(poses[0]) @ poses[v]) for v in range(V))
    print("LUM %d views x %d points, %d outer loops: %.2f ms (host buffers), worst rotation error %.5f rad (initial %.5f)" % (
        V, n, loops, 1e3 * dt, err, max(rot_angle(np.linalg.inv(init[0]) @ init[v], np.linalg.inv(poses[0]) @ poses[v]) for v in range(V))), flush=True)
