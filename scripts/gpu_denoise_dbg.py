import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, oracle as orc, mvr_b200 as mvr, mvr_b200.synth as synth
from scipy.spatial import cKDTree
rng = np.random.default_rng(21)
scan, _ = synth.turntable_view(2, 12, 40_000)
lo, hi = scan[:, :3].min(axis=0), scan[:, :3].max(axis=0)
specks = []
for k in range(60):
    c = lo + rng.random(3) * (hi - lo) + np.array([0, 0, 60.0])
    specks.append(c + rng.normal(size=(1 + k % 14, 3)) * 0.4)
specks = np.concatenate(specks).astype(np.float32)
pts = np.ones((len(scan) + len(specks), 4), dtype=np.float32)
pts[:len(scan), :3] = scan[:, :3]; pts[len(scan):, :3] = specks
pts = pts[rng.permutation(len(pts))]
c = mvr.Context(0)
ok, on = orc.denoise(pts, 3, 2.5)
tree = cKDTree(pts[:, :3].astype(np.float64))
for rep in range(4):
    k, n = c.denoise(pts, 3, 2.5)
    print("rep", rep, "gpu noise", n, "oracle", on)
    extra = sorted(set(ok.tolist()) - set(k.tolist())); missing = sorted(set(k.tolist()) - set(ok.tolist()))
    print("  dropped by gpu but kept by oracle:", extra, " kept by gpu only:", missing)
    for i in extra[:3]:
        nb = tree.query_ball_point(pts[i, :3].astype(np.float64), 2.5)
        d = np.sqrt(((pts[nb, :3].astype(np.float64) - pts[i, :3].astype(np.float64)) ** 2).sum(axis=1))
        print("   point", i, pts[i, :3], "neighbours", nb, d)
