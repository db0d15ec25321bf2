"""GPU tests at BASELINE.json's full sizes.

Two kinds: (1) the CUDA path against the CPU oracle on the same seeded inputs at the sizes BASELINE.json names -- config 1
(50k x 30 iterations, reciprocal and one-way), one 200k ring pair x 30 of config 2, 2M-point point-to-plane with the oracle's
own k = 16 normals (config 4), 1M-target exact NN with random and cell-sorted queries (config 5) -- with the north star's
bars written out below; (2) size-independent properties of the whole 24 x 200k ring (invariance under batching / sharding,
ring closure, idempotence of a converged align) and of a 10M-query NN pass."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5         # rad, per-iteration increment and final pose (north star)
TRANS_REL_TOL = 1e-6   # relative translation
RMSE_REL_TOL = 1e-4    # final RMSE


def assert_icp_matches_oracle(r, o, tgt, iters):
    """Per-iteration parity of one align against the oracle's log: equal correspondence counts, increments within
    1e-5 rad / 1e-6 of the scene extent, mean squared distance to 1e-9 relative; final pose within 1e-5 rad / 1e-6
    relative translation; final RMSE within 1e-4 relative."""
    assert r["status"] == 0 and r["iterations"] == o["iterations"] == iters
    ext = float(np.abs(tgt[:, :3] - tgt[:, :3].mean(axis=0)).max())
    assert len(r["log"]) == len(o["log"]) == iters
    for a, b in zip(r["log"], o["log"]):
        assert a["n_corr"] == b["n_corr"], "iteration %d: %d correspondences, oracle %d" % (a["iteration"], a["n_corr"], b["n_corr"])
        assert abs(a["mse"] - b["mse"]) <= 1e-9 * max(b["mse"], 1e-30), "iteration %d: mse" % a["iteration"]
        assert rot_angle(a["delta"], b["delta"]) <= ROT_TOL, "iteration %d: increment rotation" % a["iteration"]
        dt = np.linalg.norm(a["delta"][:3, 3].astype(np.float64) - b["delta"][:3, 3].astype(np.float64))
        assert dt <= TRANS_REL_TOL * ext, "iteration %d: increment translation off by %g" % (a["iteration"], dt)
    assert rot_angle(r["final"], o["final"]) <= ROT_TOL
    tb = o["final"][:3, 3].astype(np.float64)
    assert np.linalg.norm(r["final"][:3, 3].astype(np.float64) - tb) <= TRANS_REL_TOL * max(np.linalg.norm(tb), 1e-12)
    assert abs(np.sqrt(r["mse"]) - np.sqrt(o["mse"])) <= RMSE_REL_TOL * np.sqrt(o["mse"])


def rot_angle(A, B):
    R = np.asarray(A, dtype=np.float64)[:3, :3] @ np.asarray(B, dtype=np.float64)[:3, :3].T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


# ---- (1) oracle parity at the BASELINE sizes ---------------------------------------------------------------------
@pytest.mark.parametrize("reciprocal", [True, False])
def test_config1_50k_x30_matches_oracle_per_iteration(mvr, orc, synth, reciprocal):
    """BASELINE config 1: 2-view pairwise point-to-point ICP, 50k points per view, 30 fixed iterations, gate 4 mm
    (the loop of icp.align, mvr/src/registrator.cpp:551-576)."""
    n, iters = 50_000, 30
    tgt, _ = synth.turntable_view(0, 24, n)
    src, Ts = synth.turntable_view(1, 24, n)
    guess = (synth.perturbation() @ Ts).astype(np.float32)
    c = mvr.Context(0)
    c.set_target(tgt); c.set_source(src)
    r = c.icp_align(mvr.default_params(max_iterations=iters, max_dist=4.0, reciprocal=int(reciprocal), fixed_iterations=1),
                    guess=guess, n_source=n, want_cloud=True)
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=iters, max_dist=4.0, reciprocal=reciprocal, fixed_iterations=True), guess=guess)
    assert_icp_matches_oracle(r, o, tgt, iters)
    assert np.array_equal(r["cloud"], orc.transform(src, r["final"]))
    # the correspondences of the converged pair, index for index and bit for bit
    c.set_source(r["cloud"])
    q, m, d = c.correspondences(n, 4.0, reciprocal)
    oq, om, od = orc.correspondences(r["cloud"], tgt, 4.0, reciprocal)
    assert np.array_equal(q, oq) and np.array_equal(m, om) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
    c.close()


def test_config2_one_200k_ring_pair_x30_matches_oracle(mvr, orc, synth):
    """One ring pair of BASELINE config 2 (view 1 onto view 0 of the 24 x 200k sequence, the bench's guess), 30 reciprocal
    iterations: the static source index has to stay exact over 30 in-place float transforms at |p| ~ 1000 mm."""
    V, n, iters = 24, 200_000, 30
    tgt, Tt = synth.turntable_view(0, V, n)
    src, Ts = synth.turntable_view(1, V, n)
    guess = (np.linalg.inv(Tt) @ (Ts @ synth.perturbation())).astype(np.float32)   # view_init_poses of bench.py: odd views perturbed
    c = mvr.Context(0)
    c.set_target(tgt); c.set_source(src)
    r = c.icp_align(mvr.default_params(max_iterations=iters, max_dist=4.0, reciprocal=1, fixed_iterations=1), guess=guess, n_source=n)
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=iters, max_dist=4.0, reciprocal=True, fixed_iterations=True), guess=guess)
    assert_icp_matches_oracle(r, o, tgt, iters)
    assert r["nn_queries"] == o["nn_queries"]
    # and a long align: 120 iterations, still count for count
    r2 = c.icp_align(mvr.default_params(max_iterations=120, max_dist=4.0, reciprocal=1, fixed_iterations=1), guess=guess, n_source=n)
    o2 = orc.icp_align(src, tgt, orc.make_params(max_iterations=120, max_dist=4.0, reciprocal=True, fixed_iterations=True), guess=guess)
    assert [a["n_corr"] for a in r2["log"]] == [b["n_corr"] for b in o2["log"]]
    assert rot_angle(r2["final"], o2["final"]) <= ROT_TOL
    c.close()


def test_config4_point_to_plane_2M_matches_oracle(mvr, orc, synth):
    """BASELINE config 4 against the oracle: k = 16 PCA normals of a 2M-point target (neighbour lists bit-exact on a 50k
    sample, normals to 1e-6), then 3 point-to-plane iterations with the ORACLE's normals on both sides."""
    n = 2_000_000
    tgt, _ = synth.turntable_view(0, 24, n)
    src, Ts = synth.turntable_view(1, 24, n)
    guess = (synth.perturbation() @ Ts).astype(np.float32)
    c = mvr.Context(0)
    c.set_target(tgt); c.set_source(src)
    nrm, nbr = c.estimate_normals(mvr.TARGET, n, 16, viewpoint=(0.0, 0.0, 0.0), want_neighbours=True)
    onrm, onbr = orc.estimate_normals(tgt, 16, viewpoint=(0.0, 0.0, 0.0), want_neighbours=True)
    sample = np.random.default_rng(4).choice(n, 50_000, replace=False)
    assert np.array_equal(nbr[sample], onbr[sample])
    cosang = np.sum(nrm[:, :3].astype(np.float64) * onrm[:, :3], axis=1)
    assert np.all(cosang > 1 - 1e-6)
    c.set_target_normals(onrm)
    r = c.icp_align(mvr.default_params(max_iterations=3, max_dist=4.0, reciprocal=0, fixed_iterations=1, estimator=mvr.POINT_TO_PLANE),
                    guess=guess, n_source=n)
    o = orc.icp_align(src, tgt, orc.make_params(max_iterations=3, max_dist=4.0, reciprocal=False, estimator=1, fixed_iterations=True),
                      guess=guess, tgt_normals=onrm)
    assert_icp_matches_oracle(r, o, tgt, 3)
    c.close()


@pytest.mark.parametrize("order", ["random", "morton"])
def test_config5_1M_target_nn_matches_oracle(mvr, orc, synth, order):
    """BASELINE config 5 against the oracle's kd-tree: 1M-point target, 1M queries in random and in cell-sorted order,
    then 8M queries (the dense pass) on a 1M sample of them; indices and d2 bits equal."""
    m, nq = 1_000_000, 1_000_000
    tgt, q = synth.nn_sweep_case(m, nq, order=order)
    c = mvr.Context(0)
    c.set_target(tgt)
    idx, d2 = c.nn_query(q)
    oi, od = orc.nn_kdtree(tgt, q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    tgt8, q8 = synth.nn_sweep_case(m, 8 * nq, order=order)
    assert np.array_equal(tgt8, tgt)
    idx8, d8 = c.nn_query(q8)
    sel = np.random.default_rng(5).choice(8 * nq, nq, replace=False)
    oi8, od8 = orc.nn_kdtree(tgt, q8[sel])
    assert np.array_equal(idx8[sel], oi8)
    assert np.array_equal(d8[sel].view(np.uint32), od8.view(np.uint32))
    c.close()


# ---- (2) size-independent properties -------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ring24(synth):
    V, n = 24, 200_000
    views, poses = synth.turntable_sequence(V, n)
    E = synth.perturbation()
    init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
    return V, n, views, poses, init


def test_config2_ring_24x200k_properties(mvr, synth, ring24):
    """24 views x 200k points, 24 ring pairs x 30 reciprocal iterations + loop closure (the bench workload)."""
    V, n, views, poses, init = ring24
    icp = mvr.default_params(max_iterations=30, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr.RING_PAIRS, loop_closure=1, lum_iterations=16)
    reg = mvr.Registrator(0, 1)
    got, reps = reg.register_turntable(views, tp, init_poses=init)
    assert all(r["status"] == 0 and r["iterations"] == 30 for r in reps)
    assert all(80_000 < r["n_corr"] < 120_000 and r["mse"] < 0.3 for r in reps)
    # every pair's relative pose is close to the truth, and closing the ring does not make any view worse than 0.02 rad
    for p, r in enumerate(reps):
        truth = np.linalg.inv(poses[p]) @ poses[(p + 1) % V]
        assert rot_angle(r["pose"], truth) < 0.012
    assert max(rot_angle(got[v], np.linalg.inv(poses[0]) @ poses[v]) for v in range(V)) < 0.02
    # the product of the relative poses around the ring is nearly the identity (consistency of independent aligns)
    loop = np.eye(4)
    for r in reps:
        loop = loop @ r["pose"].astype(np.float64)
    assert rot_angle(loop, np.eye(4)) < 0.05
    # invariance: one pair per launch instead of all 24, and a shard of the ring, give bit-identical pair results
    reg.context(0).set_batch_group(1)
    _, reps1 = reg.register_turntable(views, tp, init_poses=init)
    reg.context(0).set_batch_group(24)
    for a, b in zip(reps, reps1):
        assert np.array_equal(a["pose"], b["pose"]) and a["n_corr"] == b["n_corr"] and a["mse"] == b["mse"]
    tps = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr.RING_PAIRS, loop_closure=1,
                               lum_iterations=16, pair_begin=9, pair_end=12)
    need = {9, 10, 11, 12}
    _, reps2 = reg.register_turntable([views[v] if v in need else None for v in range(V)], tps, init_poses=init)
    for p in range(9, 12):
        assert np.array_equal(reps[p]["pose"], reps2[p]["pose"]) and reps[p]["n_corr"] == reps2[p]["n_corr"]
    # one pair through the plain context API: the same bits again
    c = mvr.Context(0)
    c.set_target(views[3]); c.set_source(views[4])
    one = c.icp_align(icp, guess=(np.linalg.inv(init[3]) @ init[4]).astype(np.float32), n_source=n)
    assert np.array_equal(one["final"], reps[3]["pose"]) and one["n_corr"] == reps[3]["n_corr"]
    # idempotence: a long align from the found pose barely moves it
    long = mvr.default_params(max_iterations=200, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    a = c.icp_align(long, guess=one["final"], n_source=n)
    b = c.icp_align(mvr.default_params(max_iterations=5, max_dist=4.0, reciprocal=1, fixed_iterations=1), guess=a["final"], n_source=n)
    assert rot_angle(a["final"], b["final"]) < 2e-5 and np.linalg.norm(a["final"][:3, 3] - b["final"][:3, 3]) < 2e-2
    c.close()
    reg.close()


def test_config4_point_to_plane_2M_properties(mvr, synth):
    n = 2_000_000
    tgt, _ = synth.turntable_view(0, 24, n)
    src, Ts = synth.turntable_view(1, 24, n)
    guess = (synth.perturbation() @ Ts).astype(np.float32)
    c = mvr.Context(0)
    c.set_target(tgt); c.set_source(src)
    nrm = c.estimate_normals(mvr.TARGET, n, 16, viewpoint=(0.0, 0.0, 0.0))
    ln = np.linalg.norm(nrm[:, :3], axis=1)
    assert np.all(np.abs(ln - 1.0) < 1e-3)
    # normals of the bumpy sphere follow the radial direction up to the bumps' slope and the scan noise (0.2 mm noise on
    # a 0.2 mm point spacing leaves a k = 16 patch only moderately flat), and all of them face the viewpoint
    out = tgt[:, :3] - np.asarray(synth.PIVOT, dtype=np.float32)
    out /= np.linalg.norm(out, axis=1, keepdims=True)
    assert np.mean(np.abs(np.sum(out * nrm[:, :3], axis=1)) > 0.5) > 0.75
    assert np.all(np.sum(nrm[:, :3] * (-tgt[:, :3]), axis=1) >= 0)
    p = mvr.default_params(max_iterations=30, max_dist=4.0, reciprocal=0, fixed_iterations=1, estimator=mvr.POINT_TO_PLANE)
    r = c.icp_align(p, guess=guess, n_source=n)
    assert r["status"] == 0 and r["iterations"] == 30
    mse = [it["mse"] for it in r["log"]]
    assert mse[-1] < 0.5 * mse[0] and mse[-1] <= min(mse) * 1.05
    assert rot_angle(r["final"], Ts) < 0.003      # point-to-plane converges much closer than point-to-point in 30 iterations
    # the same align in a batch of one other size: bit-identical
    c2 = mvr.Context(0)
    c2.set_target(tgt[:50_000]); c2.set_source(src[:40_000])
    c2.estimate_normals(mvr.TARGET, 50_000, 16, viewpoint=(0.0, 0.0, 0.0))
    both = mvr.icp_align_batch([c2, c], p, [guess, guess])
    assert np.array_equal(both[1]["final"], r["final"]) and both[1]["n_corr"] == r["n_corr"]
    c.close(); c2.close()


def test_config5_nn_sweep_size_properties(mvr, synth):
    import torch
    m, nq = 1_000_000, 10_000_000
    tgt, q = synth.nn_sweep_case(m, nq, order="random")
    c = mvr.Context(0)
    c.set_target(tgt)
    tq = torch.from_numpy(q).cuda()
    ti = torch.empty(nq, dtype=torch.int32, device="cuda"); td = torch.empty(nq, dtype=torch.float32, device="cuda")
    c.nn_query_device(tq.data_ptr(), nq, ti.data_ptr(), td.data_ptr()); c.synchronize()      # dense pass (10 queries per target point)
    idx, d2 = ti.cpu().numpy(), td.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < m
    # the reported distance is the pinned float32 distance to the reported point
    t = tgt[idx, :3]
    dx, dy, dz = q[:, 0] - t[:, 0], q[:, 1] - t[:, 1], q[:, 2] - t[:, 2]
    assert np.array_equal(((dx * dx + dy * dy) + dz * dz).view(np.uint32), d2.view(np.uint32))
    # the other pass agrees bit for bit, and no random other target point is closer
    c.set_nn_options(8.0, 0.0)
    c.set_target(tgt)
    ti2 = torch.empty_like(ti); td2 = torch.empty_like(td)
    c.nn_query_device(tq.data_ptr(), nq, ti2.data_ptr(), td2.data_ptr()); c.synchronize()
    assert torch.equal(ti, ti2) and torch.equal(td.view(torch.int32), td2.view(torch.int32))
    rng = np.random.default_rng(1)
    o = tgt[rng.integers(0, m, nq), :3]
    ex, ey, ez = q[:, 0] - o[:, 0], q[:, 1] - o[:, 1], q[:, 2] - o[:, 2]
    assert np.all(((ex * ex + ey * ey) + ez * ez) >= d2)
    # target points query themselves: distance 0 at an index holding the same coordinates
    c.set_nn_options(8.0, 8.0)
    si, sd = c.nn_query(tgt[:200_000])
    assert np.all(sd == 0) and np.array_equal(tgt[si, :3], tgt[:200_000, :3])
    c.close()
