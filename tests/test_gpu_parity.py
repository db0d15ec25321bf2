"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): Morton keys / sort permutation / cell table / NN indices / d2 are
BIT-EXACT; per-iteration transforms agree within 1e-5 rad rotation and 1e-6 relative translation;
final RMSE within 1e-4 relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5       # rad
TRANS_REL_TOL = 1e-6
RMSE_REL_TOL = 1e-4


def rot_angle(A, B):
    """Angle of A*B^T from its skew part (sin(angle) = |vee(R - R^T)| / 2).  arccos(trace) is useless
    here: float32-rounded rotations are orthogonal only to ~1e-7, which arccos turns into ~3e-4 rad."""
    R = A[:3, :3].astype(np.float64) @ B[:3, :3].astype(np.float64).T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def assert_pose_close(A, B, scale=None):
    ang = rot_angle(A, B)
    assert ang <= ROT_TOL, "rotation differs by %g rad" % ang
    ta, tb = A[:3, 3].astype(np.float64), B[:3, 3].astype(np.float64)
    denom = scale if scale is not None else max(np.linalg.norm(tb), 1e-12)
    assert np.linalg.norm(ta - tb) / denom <= TRANS_REL_TOL, "translation differs: %r vs %r" % (ta, tb)


def random_cloud(rng, n, scale=50.0, center=(0.0, 0.0, 900.0)):
    p = np.empty((n, 4), dtype=np.float32)
    p[:, :3] = rng.normal(size=(n, 3)) * scale + np.asarray(center)
    p[:, 3] = 1.0
    return p


# ---- index: Morton keys, stable sort permutation, cell table ---------------------------------------
@pytest.mark.parametrize("n,bits", [(1, 3), (1000, 4), (50_000, 6), (200_000, 7)])
def test_index_bit_exact(ctx, orc, synth, n, bits):
    pts = synth.full_object(n, seed=123 + n)
    lo = pts[:, :3].min(axis=0)
    ext = float((pts[:, :3].max(axis=0) - lo).max()) or 1.0
    cell = ext * 1.0001 / (1 << bits)
    grid = dict(origin=lo, inv_cell=np.float32(1.0 / cell), cell=np.float32(cell), bits=bits)
    ctx.set_target(pts)
    ctx.index_build(0, grid)
    g, keys, perm, start = ctx.index_export(0, n)
    okeys = orc.morton_keys(pts, g["origin"], g["inv_cell"], g["bits"])
    operm = orc.stable_sort_perm(okeys)
    assert np.array_equal(perm, operm), "sort permutation differs"
    assert np.array_equal(keys, okeys[operm]), "sorted Morton keys differ"
    assert np.array_equal(start, orc.cell_table(okeys[operm], bits)), "cell table differs"


def test_index_nonfinite_points_sorted_last(ctx, orc, synth):
    pts = synth.full_object(5000, seed=9)
    pts[[3, 77, 4000], 0] = np.nan
    pts[[500], 2] = np.inf
    ctx.set_target(pts)
    ctx.index_build(0, None)
    g, keys, perm, start = ctx.index_export(0, len(pts))
    okeys = orc.morton_keys(pts, g["origin"], g["inv_cell"], g["bits"])
    operm = orc.stable_sort_perm(okeys)
    assert np.array_equal(perm, operm)
    assert set(perm[-4:].tolist()) == {3, 77, 500, 4000}
    assert start[-1] == len(pts) - 4


# ---- exact NN --------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(1, 10), (17, 100), (3000, 2000), (100_000, 50_000)])
def test_nn_bit_exact(ctx, orc, synth, m, n):
    tgt, q = synth.nn_sweep_case(m, n, seed=42 + m)
    ctx.set_target(tgt)
    idx, d2 = ctx.nn_query(q)
    oi, od = (orc.nn_brute if m * n <= 10_000_000 else orc.nn_kdtree)(tgt, q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("ppc", [1.0, 6.0, 40.0])
@pytest.mark.parametrize("m,n", [(1, 50), (300, 5000), (20_000, 200_000)])
def test_nn_dense_pass_bit_exact(ctx, orc, synth, m, n, ppc):
    """The warp-cooperative pass (many queries per target point) against the kd-tree oracle, for coarse and fine cells,
    queries near the surface, far from it and outside the grid, with non-finite points on both sides."""
    tgt, q = synth.nn_sweep_case(m, n)
    rng = np.random.default_rng(m + n)
    far = rng.choice(n, size=max(1, n // 50), replace=False)
    q[far, :3] += rng.normal(size=(len(far), 3)).astype(np.float32) * 60.0      # well off the surface / outside the grid
    q[rng.choice(n, size=3, replace=False), 1] = np.nan
    if m > 10:
        tgt[[1, 7], 0] = np.inf
    ctx.set_nn_options(ppc, 1e-9)      # always dense
    ctx.set_target(tgt)
    idx, d2 = ctx.nn_query(q)
    ctx.set_nn_options(6.0, 0.0)       # never dense: the per-thread row walk
    ctx.set_target(tgt)
    idx2, d22 = ctx.nn_query(q)
    ctx.set_nn_options(8.0, 8.0)
    oi, od = orc.nn_kdtree(tgt, q)
    assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(idx2, oi) and np.array_equal(d22.view(np.uint32), od.view(np.uint32))


def test_nn_ties_resolve_to_lowest_index(ctx, orc):
    rng = np.random.default_rng(5)
    base = random_cloud(rng, 500, scale=3.0)
    base[:, :3] = np.round(base[:, :3])            # lattice -> many exact distance ties
    tgt = np.concatenate([base, base[::-1], base])  # every point three times
    q = random_cloud(rng, 2000, scale=3.0)
    q[:, :3] = np.round(q[:, :3] * 2) / 2
    ctx.set_target(tgt)
    idx, d2 = ctx.nn_query(q)
    oi, od = orc.nn_brute(tgt, q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_nn_far_queries_and_outside_grid(ctx, orc):
    rng = np.random.default_rng(6)
    tgt = random_cloud(rng, 20_000, scale=20.0)
    q = random_cloud(rng, 3000, scale=400.0)        # most queries far outside the target's box
    ctx.set_target(tgt)
    idx, d2 = ctx.nn_query(q)
    oi, od = orc.nn_kdtree(tgt, q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_nn_degenerate_targets(ctx, orc):
    rng = np.random.default_rng(7)
    q = random_cloud(rng, 500, scale=5.0)
    for tgt in (
        np.tile(np.array([[1.0, 2.0, 903.0, 1.0]], dtype=np.float32), (300, 1)),                    # all identical
        np.stack([np.linspace(-5, 5, 400), np.zeros(400), np.full(400, 900.0), np.ones(400)], 1),    # a line
        np.concatenate([random_cloud(rng, 400, scale=5.0)[:, :2], np.full((400, 1), 900.0), np.ones((400, 1))], 1),  # a plane
    ):
        tgt = np.ascontiguousarray(tgt, dtype=np.float32)
        ctx.set_target(tgt)
        idx, d2 = ctx.nn_query(q)
        oi, od = orc.nn_brute(tgt, q)
        assert np.array_equal(idx, oi)
        assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_nn_nonfinite_queries_and_targets(ctx, orc):
    rng = np.random.default_rng(8)
    tgt = random_cloud(rng, 4000, scale=10.0)
    tgt[10, 1] = np.nan
    tgt[11, 0] = -np.inf
    q = random_cloud(rng, 1000, scale=10.0)
    q[5, 2] = np.nan
    ctx.set_target(tgt)
    idx, d2 = ctx.nn_query(q)
    oi, od = orc.nn_brute(tgt, q)
    assert np.array_equal(idx, oi)
    assert idx[5] == -1 and np.isinf(d2[5])
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("bits", [2, 5, 8])
@pytest.mark.parametrize("cell", [0.5, 3.0, 40.0])
def test_nn_exact_for_any_grid_resolution(ctx, orc, synth, bits, cell):
    tgt, q = synth.nn_sweep_case(30_000, 5000, seed=77)
    ctx.set_index_options(cell_edge=cell, max_bits=bits)
    try:
        ctx.set_target(tgt)
        idx, d2 = ctx.nn_query(q)
    finally:
        ctx.set_index_options(0.0, 8)
    oi, od = orc.nn_kdtree(tgt, q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.gpu
def test_nn_exact_when_points_and_queries_clamp_to_the_grid_boundary(ctx, orc):
    """An explicit grid that covers only the middle of the cloud: points outside it are filed in clamped
    boundary cells and queries outside it start from a clamped cell; the search must stay exact."""
    rng = np.random.default_rng(21)
    tgt = random_cloud(rng, 20_000, scale=30.0)
    q = random_cloud(rng, 4000, scale=60.0)
    ctx.set_target(tgt)
    for bits, cell in ((4, 2.0), (6, 0.5), (3, 10.0)):
        half = 0.5 * cell * (1 << bits)
        grid = dict(origin=np.array([0.0 - half, 0.0 - half, 900.0 - half], dtype=np.float32),
                    inv_cell=np.float32(1.0 / cell), cell=np.float32(cell), bits=bits)
        ctx.index_build(0, grid)
        idx, d2 = ctx.nn_query(q)
        oi, od = orc.nn_kdtree(tgt, q)
        assert np.array_equal(idx, oi)
        assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("max_dist", [1.0, 4.0, 25.0, 400.0])
def test_correspondences_far_apart_clouds(ctx, orc, max_dist):
    """Source mostly outside the gate of the target (exercises the gate prefilter and empty cells)."""
    rng = np.random.default_rng(22)
    tgt = random_cloud(rng, 15_000, scale=10.0)
    src = random_cloud(rng, 15_000, scale=10.0, center=(18.0, 0.0, 900.0))
    ctx.set_target(tgt)
    ctx.set_source(src)
    for reciprocal in (False, True):
        q, m, d = ctx.correspondences(len(src), max_dist, reciprocal)
        oq, om, od = orc.correspondences(src, tgt, max_dist, reciprocal)
        assert np.array_equal(q, oq) and np.array_equal(m, om)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32))


# ---- correspondences -------------------------------------------------------------------------------
@pytest.mark.parametrize("reciprocal", [False, True])
@pytest.mark.parametrize("max_dist", [0.5, 4.0, 1e9])
def test_correspondences_bit_exact(ctx, orc, synth, reciprocal, max_dist):
    views, _ = synth.turntable_sequence(12, 20_000)
    src, tgt = views[1], views[0]
    ctx.set_target(tgt)
    ctx.set_source(src)
    q, m, d = ctx.correspondences(len(src), max_dist, reciprocal)
    oq, om, od = orc.correspondences(src, tgt, max_dist, reciprocal)
    assert np.array_equal(q, oq) and np.array_equal(m, om)
    assert np.array_equal(d.view(np.uint32), od.view(np.uint32))


def test_correspondences_empty_and_tiny(ctx, orc):
    rng = np.random.default_rng(3)
    tgt = random_cloud(rng, 100, scale=1.0)
    src = random_cloud(rng, 2, scale=1.0) + np.float32([500, 0, 0, 0])
    ctx.set_target(tgt)
    ctx.set_source(src)
    q, m, d = ctx.correspondences(len(src), 1.0, True)
    assert len(q) == 0
    ctx.set_source(np.zeros((0, 4), dtype=np.float32))
    q, m, d = ctx.correspondences(0, 1.0, True)
    assert len(q) == 0


# ---- LUM edge statistics: moments of the correspondence pairs ------------------------------------------
@pytest.mark.parametrize("reciprocal", [True, False])
def test_pair_moments_match_numpy_sums_over_oracle_correspondences(ctx, orc, mvr, synth, reciprocal):
    src, tgt, guess, _ = _pair(synth, 30_000, n_views=12)
    ctx.set_target(tgt)
    ctx.set_source(src)
    got = ctx.pair_moments(4.0, reciprocal, guess).as_dict()
    a = orc.transform(src, guess)
    q, m, d2 = orc.correspondences(a, tgt, 4.0, reciprocal)
    want = mvr.PairMoments.from_pairs(a[q, :3], tgt[m, :3], origin=got["origin"]).as_dict()
    assert got["n"] == len(q) > 1000
    for k in ("sa", "sb", "sba", "saa", "sbb"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-11, atol=1e-7 * len(q))
    assert abs(got["d2"] - float(d2.astype(np.float64).sum())) <= 1e-12 * got["d2"]


def test_pair_moments_without_correspondences(ctx, mvr):
    rng = np.random.default_rng(4)
    ctx.set_target(random_cloud(rng, 500, scale=5.0))
    ctx.set_source(random_cloud(rng, 300, scale=5.0) + np.float32([1000, 0, 0, 0]))
    m = ctx.pair_moments(2.0, True, None)
    assert m.n == 0 and m.d2 == 0


@pytest.mark.parametrize("mode", ["WARP", "THREAD", "CELL", "SEEDED"])
def test_nn_every_kernel_bit_exact(mvr, orc, synth, mode):
    """The four NN kernels (one warp per query, per-thread row walk, cell-cooperative, cell-sorted and seeded) against the oracle: queries near the
    surface, far from the cloud (tens of cells of empty space), outside the grid, non-finite, and exact ties."""
    c = mvr.Context(0)
    c.set_nn_mode(getattr(mvr, "NN_" + mode))
    tgt, q = synth.nn_sweep_case(60_000, 30_000, seed=77)
    rng = np.random.default_rng(5)
    far = q[:2000].copy(); far[:, :3] += rng.normal(size=(2000, 3)).astype(np.float32) * 60.0     # far from the surface, some outside the grid
    dup = tgt[rng.integers(0, len(tgt), 3000)].copy()                                              # exactly on target points (d2 = 0)
    tgt2 = np.concatenate([tgt, tgt[:5000]])                                                       # duplicated target points: ties -> lowest index
    bad = q[:10].copy(); bad[::2, 0] = np.nan; bad[1::2, 2] = np.inf
    qq = np.concatenate([q, far, dup, bad])
    c.set_target(tgt2)
    idx, d2 = c.nn_query(qq)
    ok = np.isfinite(qq[:, :3]).all(axis=1)
    oi, od = orc.nn_kdtree(tgt2, qq[ok])
    assert np.array_equal(idx[ok], oi)
    assert np.array_equal(d2[ok].view(np.uint32), od.view(np.uint32))
    assert np.all(idx[~ok] == -1) and np.all(np.isinf(d2[~ok]))
    # a one-point target and an all-far batch
    c.set_target(tgt[:1])
    idx, d2 = c.nn_query(qq[ok][:5000])
    assert np.all(idx == 0)
    c.close()


# ---- ICP -------------------------------------------------------------------------------------------
def _pair(synth, n, n_views=24):
    tgt, Tt = synth.turntable_view(0, n_views, n)
    src, Ts = synth.turntable_view(1, n_views, n)
    guess = (synth.perturbation() @ Ts).astype(np.float32)
    return src, tgt, guess, Ts


@pytest.mark.parametrize("reciprocal", [True, False])
def test_icp_fixed_iterations_matches_oracle_per_iteration(ctx, orc, mvr, synth, reciprocal):
    src, tgt, guess, _ = _pair(synth, 20_000)
    iters = 12
    ctx.set_target(tgt)
    ctx.set_source(src)
    p = mvr.default_params(max_iterations=iters, max_dist=4.0, reciprocal=int(reciprocal), fixed_iterations=1)
    r = ctx.icp_align(p, guess=guess, n_source=len(src), want_cloud=True)
    op = orc.make_params(max_iterations=iters, max_dist=4.0, reciprocal=reciprocal, fixed_iterations=True)
    o = orc.icp_align(src, tgt, op, guess=guess)
    assert r["status"] == 0 and r["iterations"] == o["iterations"] == iters
    ext = float(np.abs(tgt[:, :3] - tgt[:, :3].mean(axis=0)).max())
    for a, b in zip(r["log"], o["log"]):
        assert a["n_corr"] == b["n_corr"], "iteration %d: correspondence count" % a["iteration"]
        assert abs(a["mse"] - b["mse"]) <= 1e-9 * max(b["mse"], 1e-30)
        # a delta's translation is tiny, so relative-to-scene-extent is the meaningful 1e-6 bar
        assert_pose_close(a["delta"], b["delta"], scale=ext)
    assert_pose_close(r["final"], o["final"])
    rm, orm = np.sqrt(r["mse"]), np.sqrt(o["mse"])
    assert abs(rm - orm) <= RMSE_REL_TOL * orm
    # output cloud = transform(input, final) in pinned float arithmetic
    assert np.array_equal(r["cloud"], orc.transform(src, r["final"]))


def test_icp_crowded_cell_is_reported_and_stays_exact(ctx, orc, mvr, synth):
    """Thousands of identical points (an invalid-return value repeated by a scanner) land in one grid cell: the index build no
    longer ranks that cell (it would be quadratic), says so, and the correspondences stay those of the oracle."""
    src, tgt, guess, _ = _pair(synth, 12_000)
    tgt = tgt.copy(); src = src.copy()
    tgt[:3000, :3] = tgt[5000, :3]        # 3000 copies of one target point
    src[:2500, :3] = src[7000, :3]
    ctx.set_target(tgt); ctx.set_source(src)
    p = mvr.default_params(max_iterations=4, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    r = ctx.icp_align(p, guess=guess, n_source=len(src))
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=4, max_dist=4.0, reciprocal=True, fixed_iterations=True), guess=guess)
    assert [a["n_corr"] for a in r["log"]] == [b["n_corr"] for b in o["log"]]
    assert ctx.debug_value(8) >= 2500 and "grid cell holds" in ctx.last_error()
    src2, tgt2, guess2, _ = _pair(synth, 12_000)
    ctx.set_target(tgt2); ctx.set_source(src2)
    ctx.icp_align(p, guess=guess2, n_source=len(src2))
    assert ctx.debug_value(8) == 0


def test_icp_gate_mask_changes_nothing(mvr, orc, synth):
    """The optional gate mask of the target (one bit per cell: anything within the gate?) only skips searches that cannot
    find a partner: every iteration's correspondences and the pose are bit-identical with and without it, for a pair with
    little overlap (views three steps apart) and for a gate of several cells."""
    tgt, Tt = synth.turntable_view(0, 12, 30_000)
    src, Ts = synth.turntable_view(3, 12, 30_000)
    guess = (synth.perturbation() @ np.linalg.inv(Tt) @ Ts).astype(np.float32)
    for max_dist, recip in ((4.0, 1), (4.0, 0), (9.0, 1)):
        p = mvr.default_params(max_iterations=8, max_dist=max_dist, reciprocal=recip, fixed_iterations=1)
        got = []
        for on in (False, True):
            c = mvr.Context(0)
            c.set_gate_mask(on)
            c.set_target(tgt); c.set_source(src)
            got.append(c.icp_align(p, guess=guess, n_source=len(src)))
            c.close()
        a, b = got
        assert [x["n_corr"] for x in a["log"]] == [x["n_corr"] for x in b["log"]] and np.array_equal(a["final"], b["final"]) and a["mse"] == b["mse"]
        o = orc.icp_align(src, tgt, orc.make_params(max_iterations=8, max_dist=max_dist, reciprocal=bool(recip), fixed_iterations=True), guess=guess)
        assert [x["n_corr"] for x in b["log"]] == [x["n_corr"] for x in o["log"]]


def test_set_clouds_device_equals_per_cloud_calls(mvr, synth):
    """mvr_set_clouds_device (every bounding box from one launch) leaves the contexts exactly as mvr_set_target_device /
    mvr_set_source_device do: the same aligns come out bit for bit, non-finite points and an empty cloud included; the
    iteration log and the aligned cloud, produced on demand, belong to the right align."""
    import torch
    clouds, ptrs = [], []
    for k in range(4):
        p, _ = synth.turntable_view(k, 12, 9_000 + 1_000 * k)
        if k == 1:
            p[5, 0] = np.nan; p[77, 2] = np.inf
        clouds.append(torch.from_numpy(p).cuda())
    empty = torch.zeros((0, 4), dtype=torch.float32, device="cuda")
    prm = mvr.default_params(max_iterations=6, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    guesses = [(synth.perturbation() @ np.linalg.inv(synth.view_pose(k, 12)) @ synth.view_pose(k + 1, 12)).astype(np.float32) for k in range(3)]
    a = [mvr.Context(0) for _ in range(3)]
    b = [mvr.Context(0) for _ in range(3)]
    for k in range(3):
        a[k].set_target_device(clouds[k].data_ptr(), len(clouds[k]))
        a[k].set_source_device(clouds[k + 1].data_ptr(), len(clouds[k + 1]))
    mvr.set_clouds_device(b + b, [mvr.TARGET] * 3 + [mvr.SOURCE] * 3, [clouds[k].data_ptr() for k in range(3)] + [clouds[k + 1].data_ptr() for k in range(3)],
                          [len(clouds[k]) for k in range(3)] + [len(clouds[k + 1]) for k in range(3)])
    ra = [a[k].icp_align(prm, guess=guesses[k], n_source=len(clouds[k + 1])) for k in range(3)]
    rb = mvr.icp_align_batch(b, prm, guesses)
    for k in range(3):
        assert ra[k]["status"] == rb[k]["status"] == 0
        assert np.array_equal(ra[k]["final"], rb[k]["final"]) and ra[k]["n_corr"] == rb[k]["n_corr"] and ra[k]["mse"] == rb[k]["mse"]
        la, lb = ra[k]["log"], b[k].iterations()
        assert len(la) == len(lb) == 6
        assert all(x["n_corr"] == y["n_corr"] and x["mse"] == y["mse"] and np.array_equal(x["delta"], y["delta"]) for x, y in zip(la, lb))
        assert b[k].fitness_score() == a[k].fitness_score()
    # an empty cloud through the batched call
    mvr.set_clouds_device([b[0]], [mvr.SOURCE], [empty.data_ptr() or 0], [0])
    assert mvr.icp_align_batch(b[:1], prm, guesses[:1])[0]["status"] == mvr.ERR_NO_INPUT
    for c in a + b:
        c.close()


def test_icp_align_batch_equals_separate_aligns(mvr, synth):
    """mvr_icp_align_batch: pairs of different sizes advancing in lock-step (one launch per iteration half for all of
    them) return exactly what one mvr_icp_align per pair returns; criteria stop each pair on its own."""
    sizes = [20_000, 7_000, 33_000, 0, 12_000]
    ctxs, guesses = [], []
    for k, n in enumerate(sizes):
        c = mvr.Context(0)
        tgt, _ = synth.turntable_view(k, 12, max(n, 1))
        src, Ts = synth.turntable_view(k + 1, 12, max(n, 1))
        Tt = synth.view_pose(k, 12)
        c.set_target(tgt)
        c.set_source(src[:n])
        guesses.append((synth.perturbation() @ np.linalg.inv(Tt) @ Ts).astype(np.float32))
        ctxs.append(c)
    for prm in (mvr.default_params(max_iterations=9, max_dist=4.0, reciprocal=1, fixed_iterations=1),
                mvr.default_params(max_iterations=40, max_dist=4.0, reciprocal=0, euclidean_fitness_epsilon=1e-3),
                mvr.default_params(max_iterations=25, max_dist=3.0, reciprocal=1, transformation_epsilon=1e-7)):
        got = mvr.icp_align_batch(ctxs, prm, guesses)
        for k, c in enumerate(ctxs):
            if sizes[k] == 0:
                assert got[k]["status"] == mvr.ERR_NO_INPUT
                continue
            one = c.icp_align(prm, guess=guesses[k], n_source=sizes[k])
            assert got[k]["status"] == one["status"] == 0
            assert got[k]["iterations"] == one["iterations"] and got[k]["n_corr"] == one["n_corr"] and got[k]["reason"] == one["reason"]
            assert np.array_equal(got[k]["final"], one["final"]) and got[k]["mse"] == one["mse"]
    iters = [g["iterations"] for k, g in enumerate(got) if sizes[k]]
    assert all(1 <= it <= 25 for it in iters)
    # the FIRST context of a batch drops out (empty source): the rest regroups and still equals the separate aligns
    order = [3, 0, 1, 2, 4]
    prm = mvr.default_params(max_iterations=7, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    got = mvr.icp_align_batch([ctxs[k] for k in order], prm, [guesses[k] for k in order])
    assert got[0]["status"] == mvr.ERR_NO_INPUT
    for j, k in enumerate(order[1:], start=1):
        one = ctxs[k].icp_align(prm, guess=guesses[k], n_source=sizes[k])
        assert got[j]["status"] == one["status"] == 0 and np.array_equal(got[j]["final"], one["final"]) and got[j]["n_corr"] == one["n_corr"]
    for c in ctxs:
        c.close()


def test_icp_lockstep_correspondences_bit_exact(ctx, orc, mvr, synth):
    """Given the oracle's own per-iteration cloud, the GPU correspondences are bit-identical."""
    src, tgt, guess, _ = _pair(synth, 15_000)
    cur = orc.transform(src, guess)
    ctx.set_target(tgt)
    for _ in range(4):
        ctx.set_source(cur)
        q, m, d = ctx.correspondences(len(cur), 4.0, True)
        oq, om, od = orc.correspondences(cur, tgt, 4.0, True)
        assert np.array_equal(q, oq) and np.array_equal(m, om)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32))
        T = orc.estimate_rigid_svd(cur, tgt, oq, om).astype(np.float32)
        cur = orc.transform(cur, T)


def test_icp_reference_configuration_stops_after_one_iteration(ctx, orc, mvr, synth):
    """registrationICP's settings (mvr/src/registrator.cpp:551-560): fitness epsilon 64 is forwarded
    as the RELATIVE mse threshold, so |mse - DBL_MAX| / DBL_MAX = 1 < 64 ends every align at once."""
    src, tgt, guess, _ = _pair(synth, 10_000)
    ctx.set_target(tgt)
    ctx.set_source(src)
    p = mvr.default_params(max_iterations=2**31 - 1, max_dist=4.0, reciprocal=1, transformation_epsilon=1e-6,
                           euclidean_fitness_epsilon=64.0)
    r = ctx.icp_align(p, guess=guess, n_source=len(src))
    op = orc.make_params(max_iterations=2**31 - 1, max_dist=4.0, reciprocal=True, transformation_epsilon=1e-6,
                         euclidean_fitness_epsilon=64.0)
    o = orc.icp_align(src, tgt, op, guess=guess)
    assert r["iterations"] == o["iterations"] == 1
    assert r["reason"] == o["reason"] == 4
    assert_pose_close(r["final"], o["final"])


def test_icp_criteria_transform_epsilon(ctx, orc, mvr, synth):
    src, tgt, guess, _ = _pair(synth, 10_000)
    ctx.set_target(tgt)
    ctx.set_source(src)
    kw = dict(max_iterations=60, max_dist=4.0, transformation_epsilon=1e-4)
    r = ctx.icp_align(mvr.default_params(reciprocal=1, **kw), guess=guess, n_source=len(src))
    o = orc.icp_align(src, tgt, orc.make_params(reciprocal=True, **kw), guess=guess)
    assert r["iterations"] == o["iterations"] and r["reason"] == o["reason"]
    assert_pose_close(r["final"], o["final"])


def test_icp_too_few_correspondences(ctx, mvr):
    rng = np.random.default_rng(1)
    tgt = random_cloud(rng, 1000, scale=1.0)
    src = random_cloud(rng, 1000, scale=1.0) + np.float32([1000, 0, 0, 0])
    ctx.set_target(tgt)
    ctx.set_source(src)
    r = ctx.icp_align(mvr.default_params(max_iterations=5, max_dist=1.0, reciprocal=1), n_source=len(src))
    assert r["status"] == mvr.ERR_TOO_FEW and r["reason"] == 5 and r["converged"] == 0 and r["iterations"] == 0
    assert np.array_equal(r["final"], np.eye(4, dtype=np.float32))


def test_icp_known_answer_zero_noise(ctx, mvr, synth):
    """Same points under a known rigid transform: ICP must recover it."""
    tgt = synth.full_object(30_000, seed=3)
    T = synth.rotation_about_axis(np.deg2rad(1.5), axis=(0.3, 1.0, -0.2))
    T[:3, 3] += [0.8, -0.5, 0.3]
    Ti = np.linalg.inv(T)
    src = np.ones_like(tgt)
    src[:, :3] = (tgt[:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    ctx.set_target(tgt)
    ctx.set_source(src)
    r = ctx.icp_align(mvr.default_params(max_iterations=50, max_dist=10.0, reciprocal=1, fixed_iterations=1), n_source=len(src))
    assert rot_angle(r["final"], T) < 2e-6
    assert np.linalg.norm(r["final"][:3, 3] - T[:3, 3]) < 2e-3   # float32 coordinates at |p| ~ 1000 mm
    assert r["mse"] < 1e-6


def test_fitness_score_matches_oracle(ctx, orc, mvr, synth):
    src, tgt, guess, _ = _pair(synth, 15_000)
    ctx.set_target(tgt)
    ctx.set_source(src)
    p = mvr.default_params(max_iterations=5, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    r = ctx.icp_align(p, guess=guess, n_source=len(src), want_cloud=True)
    f = ctx.fitness_score()
    of = orc.fitness_score(r["cloud"], tgt)
    assert abs(f - of) <= 1e-12 * of
    f2 = ctx.fitness_score(max_range=1.0)
    of2 = orc.fitness_score(r["cloud"], tgt, max_range=1.0)
    assert abs(f2 - of2) <= 1e-12 * of2


def test_point_to_plane_matches_oracle(ctx, orc, mvr, synth):
    src, tgt, guess, _ = _pair(synth, 20_000)
    ctx.set_target(tgt)
    ctx.set_source(src)
    nrm = ctx.estimate_normals(0, len(tgt), 16, viewpoint=(0, 0, 0))
    onrm = orc.estimate_normals(tgt, 16, viewpoint=(0, 0, 0))
    cosang = np.abs(np.sum(nrm[:, :3].astype(np.float64) * onrm[:, :3], axis=1))
    assert np.all(cosang > 1 - 1e-6)
    ctx.set_target_normals(onrm)   # same normals on both sides so the ICP comparison is about ICP only
    kw = dict(max_iterations=8, max_dist=4.0, fixed_iterations=1)
    r = ctx.icp_align(mvr.default_params(reciprocal=0, estimator=1, **kw), guess=guess, n_source=len(src))
    o = orc.icp_align(src, tgt, orc.make_params(reciprocal=False, estimator=1, fixed_iterations=True, max_iterations=8, max_dist=4.0),
                      guess=guess, tgt_normals=onrm)
    assert r["iterations"] == o["iterations"] == 8
    for a, b in zip(r["log"], o["log"]):
        assert a["n_corr"] == b["n_corr"]
    assert_pose_close(r["final"], o["final"])


def test_normals_neighbours_bit_exact(ctx, orc, synth):
    pts = synth.full_object(20_000, seed=11, noise=0.1)
    ctx.set_source(pts)
    nrm, nbr = ctx.estimate_normals(1, len(pts), 16, viewpoint=(0, 0, 0), want_neighbours=True)
    onrm, onbr = orc.estimate_normals(pts, 16, viewpoint=(0, 0, 0), want_neighbours=True)
    assert np.array_equal(nbr, onbr)
    cosang = np.sum(nrm[:, :3].astype(np.float64) * onrm[:, :3], axis=1)
    assert np.all(cosang > 1 - 1e-6)
    assert np.allclose(nrm[:, 3], onrm[:, 3], rtol=1e-4, atol=1e-7)
