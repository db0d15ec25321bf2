"""CPU tests (no GPU): pin the oracle -- the checker every GPU parity test leans on -- against independent
implementations (brute force, scipy cKDTree, numpy SVD / Kabsch) and against the committed golden vectors.
The reference has no tests of its own and its PCL arithmetic cannot be built here (parity unpinned),
so this is the strongest pin available."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def d2_pinned(q, p):
    """((dx*dx + dy*dy) + dz*dz) in float32, no FMA -- numpy evaluates exactly that."""
    d = (q[:, None, :3].astype(np.float32) - p[None, :, :3].astype(np.float32))
    sq = d * d
    return (sq[..., 0] + sq[..., 1]) + sq[..., 2]


def test_nn_brute_matches_numpy_pinned_arithmetic(orc, synth):
    tgt, q = synth.nn_sweep_case(700, 300, seed=3)
    idx, d2 = orc.nn_brute(tgt, q)
    D = d2_pinned(q, tgt)
    ref_d2 = D.min(axis=1)
    ref_idx = np.array([np.flatnonzero(D[i] == ref_d2[i])[0] for i in range(len(q))])   # lowest index among exact ties
    assert np.array_equal(idx, ref_idx)
    assert np.array_equal(d2.view(np.uint32), ref_d2.view(np.uint32))


def test_nn_kdtree_equals_brute_and_scipy(orc, synth):
    from scipy.spatial import cKDTree
    tgt, q = synth.nn_sweep_case(20_000, 5000, seed=11)
    ki, kd = orc.nn_kdtree(tgt, q)
    bi, bd = orc.nn_brute(tgt, q)
    assert np.array_equal(ki, bi) and np.array_equal(kd.view(np.uint32), bd.view(np.uint32))
    sd, si = cKDTree(tgt[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=1)
    same = si == ki
    # cKDTree works in double: where it disagrees the two candidates must tie in float32 d2 (or be within rounding)
    D_alt = d2_pinned(q[~same], tgt[si[~same]][None, :, :].reshape(-1, 4))
    alt = np.array([D_alt[k, k] for k in range(D_alt.shape[0])]) if D_alt.size else np.zeros(0, np.float32)
    assert np.all(np.abs(alt - kd[~same]) <= 4e-7 * np.maximum(kd[~same], 1e-12))
    assert same.mean() > 0.999
    assert np.allclose(np.sqrt(kd[same].astype(np.float64)), sd[same], rtol=2e-6, atol=1e-7)


def test_nn_ties_golden(orc):
    g = gold("nn_ties.npz")
    idx, d2 = orc.nn_brute(g["tgt"], g["q"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2.view(np.uint32), g["d2"].view(np.uint32))
    ki, kd = orc.nn_kdtree(g["tgt"], g["q"])
    assert np.array_equal(ki, g["idx"]) and np.array_equal(kd.view(np.uint32), g["d2"].view(np.uint32))
    # every target point exists three times: the winner is always the first copy
    assert np.all(idx < 200)


def test_correspondences_definition(orc, synth):
    """determineCorrespondences / determineReciprocalCorrespondences restated in numpy."""
    views, _ = synth.turntable_sequence(12, 1200)
    src, tgt = views[1], views[0]
    Dst = d2_pinned(src, tgt)
    j = Dst.argmin(axis=1)
    d = Dst[np.arange(len(src)), j]
    for max_dist in (0.5, 4.0, 1e9):
        keep = ~(d.astype(np.float64) > max_dist * max_dist)
        q, m, dd = orc.correspondences(src, tgt, max_dist, False)
        assert np.array_equal(q, np.flatnonzero(keep)) and np.array_equal(m, j[keep])
        assert np.array_equal(dd.view(np.uint32), d[keep].view(np.uint32))
        back = Dst.T.argmin(axis=1)            # nearest source point of every target point (lowest index on ties)
        rk = keep & (back[j] == np.arange(len(src)))
        q, m, dd = orc.correspondences(src, tgt, max_dist, True)
        assert np.array_equal(q, np.flatnonzero(rk)) and np.array_equal(m, j[rk])


def test_svd3_and_umeyama_against_numpy(orc):
    rng = np.random.default_rng(0)
    for k in range(200):
        A = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-3, 4)
        if k % 5 == 0:
            A[2] = A[0] + A[1]            # rank 2
        U, s, V = orc.svd3(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=1e-12 * max(1.0, np.abs(A).max()))
        # the oracle goes through A^T A: a vanishing singular value is only resolved to sqrt(eps) * s_max
        assert np.allclose(s, np.linalg.svd(A, compute_uv=False), rtol=1e-10, atol=1e-7 * np.abs(A).max())
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-12)
    # Kabsch / Umeyama (with_scaling = false) on exact correspondences recovers the motion
    pts = rng.normal(size=(500, 3)) * 30 + [0, 0, 900]
    ang, ax = 0.3, np.array([0.2, -1.0, 0.4]) / np.linalg.norm([0.2, -1.0, 0.4])
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = np.array([1.0, -2.0, 0.5])
    src = np.ones((500, 4), np.float32)
    dst = np.ones((500, 4), np.float32)
    src[:, :3] = pts
    dst[:, :3] = pts @ R.T + t
    T = orc.estimate_rigid_svd(src, dst, np.arange(500, dtype=np.int32), np.arange(500, dtype=np.int32))
    assert np.allclose(T[:3, :3], R, atol=2e-6) and np.allclose(T[:3, 3], t, atol=2e-3)


def test_icp_golden_and_known_answer(orc, synth):
    for recip in (0, 1):
        g = gold("icp_pair_recip%d.npz" % recip)
        o = orc.icp_align(g["src"], g["tgt"], orc.make_params(max_iterations=8, max_dist=4.0, reciprocal=bool(recip), fixed_iterations=True),
                          guess=g["guess"])
        assert np.array_equal(o["final"], g["final"])
        assert [r["n_corr"] for r in o["log"]] == g["n_corr"].tolist()
        assert np.array_equal(o["cloud"], g["cloud"])
        q, m, d = orc.correspondences(orc.transform(g["src"], g["guess"]), g["tgt"], 4.0, bool(recip))
        assert np.array_equal(q, g["corr_q"]) and np.array_equal(m, g["corr_m"]) and np.array_equal(d, g["corr_d"])
    # zero-noise known answer: same surface points under a known small motion
    tgt = synth.full_object(4000, seed=3)
    T = synth.rotation_about_axis(np.deg2rad(1.0), axis=(0.3, 1.0, -0.2))
    T[:3, 3] += [0.5, -0.3, 0.2]
    Ti = np.linalg.inv(T)
    src = np.ones_like(tgt)
    src[:, :3] = (tgt[:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=40, max_dist=10.0, reciprocal=True, fixed_iterations=True))
    assert np.abs(o["final"][:3, :3] - T[:3, :3]).max() < 1e-5 and o["mse"] < 1e-6


def test_icp_reference_configuration_one_iteration(orc, synth):
    """fitEps 64 forwarded as the relative-MSE threshold ends every align after one iteration."""
    g = gold("icp_pair_recip1.npz")
    p = orc.make_params(max_iterations=2**31 - 1, max_dist=4.0, reciprocal=True, transformation_epsilon=1e-6, euclidean_fitness_epsilon=64.0)
    o = orc.icp_align(g["src"], g["tgt"], p, guess=g["guess"])
    assert o["iterations"] == 1 and o["reason"] == 4 and o["converged"]
    far = g["src"].copy()
    far[:, 0] += 1e4
    o = orc.icp_align(far, g["tgt"], p)
    assert o["status"] != 0 and o["iterations"] == 0 and not o["converged"]      # "Not enough correspondences"


def test_index_golden_and_self_consistency(orc):
    g = gold("index.npz")
    keys = orc.morton_keys(g["pts"], g["origin"], g["inv_cell"], int(g["bits"]))
    assert np.array_equal(keys, g["keys"])
    perm = orc.stable_sort_perm(keys)
    assert np.array_equal(perm, g["perm"]) and np.array_equal(perm, np.argsort(keys, kind="stable"))
    start = orc.cell_table(keys[perm], int(g["bits"]))
    assert np.array_equal(start, g["start"])
    assert np.array_equal(start, np.searchsorted(keys[perm], np.arange((1 << (3 * int(g["bits"]))) + 1)))
    # Morton interleave: x bit 0, y bit 1, z bit 2
    c = np.clip(np.floor((g["pts"][:, :3] - g["origin"]) * g["inv_cell"]), 0, (1 << int(g["bits"])) - 1).astype(np.int64)
    ref = np.zeros(len(c), dtype=np.int64)
    for b in range(int(g["bits"])):
        for a in range(3):
            ref |= ((c[:, a] >> b) & 1) << (3 * b + a)
    assert np.array_equal(keys.astype(np.int64), ref)


def test_poses_golden(orc):
    g = gold("poses.npz")
    assert np.array_equal(orc.apply_pose_double(g["pts"], g["M"]), g["posed"])
    ref = (g["pts"][:, :3].astype(np.float64) @ g["M"][:3, :3].T + g["M"][:3, 3]).astype(np.float32)
    assert np.abs(g["posed"][:, :3] - ref).max() <= 6.2e-5      # one float32 ulp at |p| ~ 1000
    assert np.array_equal(orc.transform(g["pts"], g["M"].astype(np.float32)), g["moved"])


def test_normals_golden_and_plane(orc):
    g = gold("normals.npz")
    nrm, nbr = orc.estimate_normals(g["pts"], 16, viewpoint=(0, 0, 0), want_neighbours=True)
    assert np.array_equal(nbr, g["neighbours"]) and np.allclose(nrm, g["normals"], atol=1e-6)
    rng = np.random.default_rng(4)
    pl = np.ones((400, 4), np.float32)
    pl[:, :2] = rng.uniform(-5, 5, size=(400, 2))
    pl[:, 2] = 900.0
    n = orc.estimate_normals(pl, 12, viewpoint=(0, 0, 0))
    assert np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-6) and np.all(n[:, 2] < 0)        # flipped towards the origin
    assert np.allclose(n[:, 3], 0.0, atol=1e-6)                                          # curvature of a plane


def _xform_pinned(M, p):
    """x' = ((m00*x + m01*y) + m02*z) + m03 in float32 without FMA (ICP's in-place transformCloud)."""
    M = M.astype(np.float32)
    x, y, z = p[:, 0].astype(np.float32), p[:, 1].astype(np.float32), p[:, 2].astype(np.float32)
    out = np.ones_like(p, dtype=np.float32)
    for r in range(3):
        out[:, r] = ((M[r, 0] * x + M[r, 1] * y) + M[r, 2] * z) + M[r, 3]
    return out


def _kabsch(a, b):
    """Eigen::umeyama(with_scaling = false) on double centroids / cross-covariance: b ~ R a + t."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    ma, mb = a.mean(axis=0), b.mean(axis=0)
    S = (b - mb).T @ (a - ma) / len(a)
    U, _, Vt = np.linalg.svd(S)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1.0
    R = U @ D @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mb - R @ ma
    return T


@pytest.mark.parametrize("reciprocal", [True, False])
def test_icp_loop_against_an_independent_numpy_restatement(orc, synth, reciprocal):
    """The WHOLE loop of pcl::IterativeClosestPoint::computeTransformation restated with numpy brute force
    and numpy's SVD -- guess applied in float, exact NN with lowest-index ties, (reciprocal test,) double gate, Kabsch on
    double sums, increment cast to float and applied in place, final = increment * final -- against the oracle's C loop:
    the correspondence count of every iteration equal, every increment and the final pose equal to float rounding."""
    views, poses = synth.turntable_sequence(12, 1500)
    src, tgt = views[1], views[0]
    guess = (synth.perturbation() @ np.linalg.inv(poses[0]) @ poses[1]).astype(np.float32)
    iters, max_dist = 6, 4.0
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=iters, max_dist=max_dist, reciprocal=reciprocal, fixed_iterations=True), guess=guess)
    assert o["status"] == 0 and o["iterations"] == iters
    cur = _xform_pinned(guess, src)
    final = guess.astype(np.float64)
    for it in range(iters):
        D = d2_pinned(cur, tgt)
        j = D.argmin(axis=1)
        d = D[np.arange(len(cur)), j]
        keep = ~(d.astype(np.float64) > max_dist * max_dist)
        if reciprocal:
            keep &= D.T.argmin(axis=1)[j] == np.arange(len(cur))
        rec = o["log"][it]
        assert rec["n_corr"] == int(keep.sum()), "iteration %d" % it
        assert rec["mse"] == pytest.approx(float(d[keep].astype(np.float64).mean()), rel=1e-12)
        T = _kabsch(cur[keep, :3], tgt[j[keep], :3]).astype(np.float32)
        assert np.allclose(rec["delta"], T.astype(np.float64), rtol=0, atol=2e-6 * max(1.0, np.abs(T[:3, 3]).max()))
        Tf = rec["delta"].astype(np.float32)      # continue from the oracle's own increment: rounding does not accumulate
        cur = _xform_pinned(Tf, cur)
        final = Tf.astype(np.float64) @ final
    assert np.array_equal(o["final"], final.astype(np.float32))
    assert np.array_equal(o["cloud"][:, :3], _xform_pinned(o["final"], src)[:, :3])


def test_point_to_plane_and_normals_against_numpy(orc, synth):
    """TransformationEstimationPointToPlaneLLS and the kNN-PCA normals (A13) restated with numpy: the 6-vector
    from numpy's least squares over the rows [cross(s, n), n | n.(d - s)], the pose from the same Rz Ry Rx composition;
    normals from numpy's brute-force k nearest neighbours and numpy's symmetric eigen-solver."""
    rng = np.random.default_rng(11)
    tgt = synth.full_object(3000, seed=5)
    k = 12
    nrm, nbr = orc.estimate_normals(tgt, k, viewpoint=(0, 0, 0), want_neighbours=True)
    # normals: brute-force kNN (self included, (d2, index) order) + smallest eigenvector of the neighbours' covariance
    D = d2_pinned(tgt, tgt)
    order = np.lexsort((np.broadcast_to(np.arange(len(tgt)), D.shape), D), axis=1)[:, :k]
    assert np.array_equal(nbr, order.astype(np.int32))
    for i in rng.integers(0, len(tgt), 200):
        P = tgt[order[i], :3].astype(np.float64)
        C = np.cov(P.T, bias=True)
        w, V = np.linalg.eigh(C)
        n = V[:, 0]
        if n @ (0.0 - tgt[i, :3].astype(np.float64)) < 0:
            n = -n
        assert abs(float(n @ nrm[i, :3].astype(np.float64))) > 1 - 1e-6 and float(n @ nrm[i, :3]) > 0
        assert nrm[i, 3] == pytest.approx(w[0] / w.sum(), rel=1e-4, abs=1e-7)
    # point-to-plane: a small known motion, correspondences by index
    T = synth.rotation_about_axis(np.deg2rad(0.4), axis=(0.2, 1.0, -0.5))
    T[:3, 3] += [0.3, -0.2, 0.1]
    Ti = np.linalg.inv(T)
    src = np.ones_like(tgt)
    src[:, :3] = (tgt[:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    idx = np.arange(len(tgt), dtype=np.int32)
    rc, To = orc.estimate_point_to_plane(src, tgt, nrm, idx, idx)
    assert rc == 1 or rc == 0
    s, d, n = src[:, :3].astype(np.float64), tgt[:, :3].astype(np.float64), nrm[:, :3].astype(np.float64)
    A = np.concatenate([np.cross(s, n), n], axis=1)
    b = np.einsum("ij,ij->i", n, d - s)
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    R = np.array([[cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca],
                  [sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca],
                  [-sb, cb * sa, cb * ca]])
    assert np.allclose(To[:3, :3], R, atol=1e-9) and np.allclose(To[:3, 3], x[3:], atol=1e-7)
    assert np.allclose(To[:3, :3], T[:3, :3], atol=5e-5)          # one linearised step lands close to the motion


def test_denoise_oracle_equals_the_delaunay_construction(orc, synth):
    """PointCloud::denoise takes its edges from a 3-D Delaunay triangulation (mvr/src/point_cloud.cpp:469-497: every edge no
    longer than triangle_length) and labels connected components; the oracle (and the GPU kernel) use the plain radius
    graph.  The two have the same components (the short edges of the Euclidean minimum spanning tree lie in both): checked
    here against scipy's Delaunay triangulation, kept points and their output order."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(21)
    obj = synth.full_object(2500, seed=9)[:, :3].astype(np.float64)
    blobs = np.concatenate([c + rng.normal(size=(m, 3)) * 0.8 for c, m in (((150.0, 0.0, 900.0), 14), ((-150.0, 30.0, 905.0), 6), ((0.0, 160.0, 880.0), 9))])
    lone = rng.uniform(-200, 200, size=(25, 3)) + [0, 0, 900]
    pts = np.concatenate([obj, blobs, lone])
    pts = pts[rng.permutation(len(pts))].astype(np.float32).astype(np.float64)   # the oracle gets float records: the same coordinates
    n = len(pts)
    for length, thr in ((6.0, 10), (4.0, 8), (9.0, 12)):
        tri = Delaunay(pts)
        e = np.concatenate([tri.simplices[:, [a, b]] for a in range(4) for b in range(a + 1, 4)])
        e = e[np.linalg.norm(pts[e[:, 0]] - pts[e[:, 1]], axis=1) <= length]
        _, label = connected_components(coo_matrix((np.ones(len(e), np.int8), (e[:, 0], e[:, 1])), shape=(n, n)), directed=False)
        size = np.bincount(label)
        first = np.full(len(size), n)
        np.minimum.at(first, label, np.arange(n))
        keep = np.flatnonzero(size[label] >= thr)
        keep = keep[np.lexsort((keep, first[label[keep]]))]
        okeep, onoise = orc.denoise(np.concatenate([pts, np.ones((n, 1))], axis=1).astype(np.float32), thr, length)
        assert onoise == n - len(keep) and np.array_equal(okeep, keep.astype(np.int32))
        assert 0 < onoise < n


def test_convergence_criteria_and_fitness_against_numpy(orc, synth):
    """DefaultConvergenceCriteria replayed in numpy over the oracle's own iteration log -- iteration cap,
    transform epsilon, absolute and relative mse, in that order, prev_mse starting at DBL_MAX -- must stop where the oracle
    stopped and for the same reason; getFitnessScore (A9) = mean of the un-gated NN d2 of the aligned cloud within max_range."""
    import sys
    views, poses = synth.turntable_sequence(12, 1500)
    src, tgt = views[1], views[0]
    guess = (synth.perturbation() @ np.linalg.inv(poses[0]) @ poses[1]).astype(np.float32)
    cases = [dict(max_iterations=50, transformation_epsilon=1e-7, euclidean_fitness_epsilon=1e-9),
             dict(max_iterations=50, transformation_epsilon=0.0, euclidean_fitness_epsilon=1e-3),
             dict(max_iterations=4, transformation_epsilon=0.0, euclidean_fitness_epsilon=0.0),
             dict(max_iterations=50, transformation_epsilon=1e-4, euclidean_fitness_epsilon=0.0)]
    seen = set()
    for c in cases:
        o = orc.icp_align(src, tgt, orc.make_params(max_dist=4.0, reciprocal=True, **c), guess=guess)
        full = orc.icp_align(src, tgt, orc.make_params(max_iterations=c["max_iterations"], max_dist=4.0, reciprocal=True, fixed_iterations=True), guess=guess)
        prev, stop, why = sys.float_info.max, None, None
        for it, rec in enumerate(full["log"], start=1):
            T = rec["delta"]
            cos_angle = 0.5 * (np.trace(T[:3, :3]) - 1.0)
            t2 = float(T[:3, 3] @ T[:3, 3])
            if it >= c["max_iterations"]:
                stop, why = it, 1
            elif cos_angle >= 1.0 - c["transformation_epsilon"] and t2 <= c["transformation_epsilon"]:
                stop, why = it, 2
            elif abs(rec["mse"] - prev) < 1e-12:
                stop, why = it, 3
            elif abs(rec["mse"] - prev) / prev < c["euclidean_fitness_epsilon"]:
                stop, why = it, 4
            if stop:
                break
            prev = rec["mse"]
        assert (o["iterations"], o["reason"]) == (stop, why), c
        assert [r["n_corr"] for r in o["log"]] == [r["n_corr"] for r in full["log"][:stop]]
        seen.add(why)
    assert {1, 2, 4} <= seen
    D = d2_pinned(o["cloud"], tgt).min(axis=1).astype(np.float64)
    assert orc.fitness_score(o["cloud"], tgt) == pytest.approx(D.mean(), rel=1e-13)
    lim = float(np.median(D))
    assert orc.fitness_score(o["cloud"], tgt, max_range=lim) == pytest.approx(D[D <= lim].mean(), rel=1e-13)
