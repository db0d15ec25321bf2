#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no fixtures of its own and its arithmetic lives in an un-vendored PCL that
cannot be built here, so these vectors are NOT outputs of the reference: they freeze the oracle's answers on small,
seeded cases after those answers were checked against independent implementations (brute force, scipy cKDTree,
numpy SVD -- tests/test_oracle.py), so that neither the oracle nor the CUDA path can drift unnoticed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc          # noqa: E402
import mvr_b200.synth as synth  # noqa: E402  (numpy only; does not load the CUDA library)

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    orc.build()
    # 1. pairwise ICP on a tiny turntable pair (reference settings: reciprocal, gate 4, fixed 8 iterations)
    tgt, _ = synth.turntable_view(0, 12, 1500)
    src, Ts = synth.turntable_view(1, 12, 1500)
    guess = (synth.perturbation() @ Ts).astype(np.float32)
    for recip in (0, 1):
        o = orc.icp_align(src, tgt, orc.make_params(max_iterations=8, max_dist=4.0, reciprocal=bool(recip), fixed_iterations=True), guess=guess)
        q, m, d = orc.correspondences(orc.transform(src, guess), tgt, 4.0, bool(recip))
        np.savez_compressed(os.path.join(HERE, "icp_pair_recip%d.npz" % recip), src=src, tgt=tgt, guess=guess, final=o["final"],
                            n_corr=np.array([r["n_corr"] for r in o["log"]], dtype=np.int32),
                            mse=np.array([r["mse"] for r in o["log"]], dtype=np.float64),
                            corr_q=q, corr_m=m, corr_d=d, cloud=o["cloud"])
    # 2. exact NN with ties (lattice points, every target point three times) and far / outside queries
    rng = np.random.default_rng(5)
    base = np.ones((200, 4), dtype=np.float32)
    base[:, :3] = np.round(rng.normal(size=(200, 3)) * 3.0) + [0, 0, 900]
    tgt = np.concatenate([base, base[::-1], base])
    q = np.ones((400, 4), dtype=np.float32)
    q[:, :3] = np.round(rng.normal(size=(400, 3)) * 6.0) / 2 + [0, 0, 900]
    q[-20:, :3] *= 40.0
    idx, d2 = orc.nn_brute(tgt, q)
    np.savez_compressed(os.path.join(HERE, "nn_ties.npz"), tgt=tgt, q=q, idx=idx, d2=d2)
    # 3. index: Morton keys, stable permutation, cell table
    pts = synth.full_object(2000, seed=77)
    lo = pts[:, :3].min(axis=0)
    ext = float((pts[:, :3].max(axis=0) - lo).max())
    bits = 5
    inv_cell = np.float32(1.0 / (ext * 1.0001 / (1 << bits)))
    keys = orc.morton_keys(pts, lo, inv_cell, bits)
    perm = orc.stable_sort_perm(keys)
    np.savez_compressed(os.path.join(HERE, "index.npz"), pts=pts, origin=lo.astype(np.float32), inv_cell=inv_cell, bits=bits, keys=keys,
                        perm=perm, start=orc.cell_table(keys[perm], bits))
    # 4. pose application in double (getTransformedPoints) and the pinned float transform (transformCloud)
    M = synth.rotation_about_axis(0.7, axis=(0.2, -1.0, 0.1))
    M[:3, 3] += [3.0, -2.0, 1.5]
    np.savez_compressed(os.path.join(HERE, "poses.npz"), pts=pts[:500], M=M, posed=orc.apply_pose_double(pts[:500], M),
                        moved=orc.transform(pts[:500], M.astype(np.float32)))
    # 5. kNN-PCA normals (k = 16)
    nrm, nbr = orc.estimate_normals(pts[:800], 16, viewpoint=(0, 0, 0), want_neighbours=True)
    np.savez_compressed(os.path.join(HERE, "normals.npz"), pts=pts[:800], normals=nrm, neighbours=nbr)
    print("golden vectors written to", HERE)


def lum_fixture():
    """LUM relaxation (host/lum.cpp lumRelax): point pairs of a 5-view graph (ring + one chord), noisy, and the corrections
    that minimise sum |X_s a - X_t b|^2 with X_0 = I, found by scipy.optimize.least_squares on the explicit residuals --
    an implementation that shares nothing with the moment-based Gauss-Newton of the library."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(20261018)
    V = 5

    def rigid(rot, trans):
        T = np.eye(4)
        T[:3, :3] = Rotation.from_rotvec(rot * rng.standard_normal(3)).as_matrix()
        T[:3, 3] = trans * rng.standard_normal(3)
        return T

    def apply(T, p):
        return p @ T[:3, :3].T + T[:3, 3]

    disp = [np.eye(4)] + [rigid(0.02, 1.0) for _ in range(V - 1)]
    edges = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0), (1, 3)]
    A, B = [], []
    for (s, t) in edges:
        p = rng.standard_normal((150, 3)) * 45 + [0, 0, 900]
        A.append(apply(disp[s], p) + 0.25 * rng.standard_normal(p.shape))
        B.append(apply(disp[t], p) + 0.25 * rng.standard_normal(p.shape))

    def pose(x6):
        T = np.eye(4)
        T[:3, :3] = Rotation.from_rotvec(x6[:3]).as_matrix()
        T[:3, 3] = x6[3:]
        return T

    def resid(x):
        Xs = [np.eye(4)] + [pose(x[6 * k:6 * k + 6]) for k in range(V - 1)]
        return np.concatenate([(apply(Xs[s], a) - apply(Xs[t], b)).ravel() for (s, t), a, b in zip(edges, A, B)])

    sol = least_squares(resid, np.zeros(6 * (V - 1)), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    X = np.stack([np.eye(4)] + [pose(sol.x[6 * k:6 * k + 6]) for k in range(V - 1)])
    np.savez_compressed(os.path.join(HERE, "lum.npz"), src=np.array([e[0] for e in edges], dtype=np.int32),
                        tgt=np.array([e[1] for e in edges], dtype=np.int32), a=np.stack(A), b=np.stack(B), X=X, cost=2.0 * sol.cost)


if __name__ == "__main__":
    main()
    lum_fixture()
