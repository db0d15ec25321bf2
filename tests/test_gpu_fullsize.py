"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle would take minutes here):
invariance under batching / sharding, ring closure, idempotence of a converged align, self-consistency of NN results."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rot_angle(A, B):
    R = np.asarray(A, dtype=np.float64)[:3, :3] @ np.asarray(B, dtype=np.float64)[:3, :3].T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


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
