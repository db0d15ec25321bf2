"""GPU parity of the C++ registration driver (host/registrator.cpp behind the C ABI) against the CPU oracle
replaying the reference's workflows (mvr/src/registrator.cpp:517-588, 746-842, 877-990) step by step."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5       # rad (north star)
TRANS_REL_TOL = 1e-6


def rot_angle(A, B):
    R = np.asarray(A, dtype=np.float64)[:3, :3] @ np.asarray(B, dtype=np.float64)[:3, :3].T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def assert_pose_close(A, B, scale):
    assert rot_angle(A, B) < ROT_TOL
    dt = np.linalg.norm(np.asarray(A, dtype=np.float64)[:3, 3] - np.asarray(B, dtype=np.float64)[:3, 3])
    assert dt <= TRANS_REL_TOL * scale, (dt, scale)


@pytest.fixture(scope="module")
def seq(synth):
    V, n = 6, 8000
    views, poses = synth.turntable_sequence(V, n)
    E = synth.perturbation()
    init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
    return V, n, views, poses, init


def test_apply_pose_bit_exact(ctx, orc, synth):
    """PointCloud::getTransformedPoints: double multiply narrowed to float, 48-byte rich points accepted."""
    rng = np.random.default_rng(2)
    rich = rng.normal(size=(5000, 12)).astype(np.float32) * 50 + 400   # PointXYZRGBNormal-sized records
    M = synth.rotation_about_axis(0.7, axis=(0.2, -1.0, 0.1))
    M[:3, 3] += [3.0, -2.0, 1.5]
    out = ctx.apply_pose(rich, M)
    ref = orc.apply_pose_double(rich, M)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    xyz = np.ascontiguousarray(rich[:, :4])
    assert np.array_equal(ctx.apply_pose(xyz, M)[:, :3], out[:, :3])


def test_turntable_rotation_matches_reference_formula(mvr, synth):
    """getRotationMatrix = T(pivot) R(axis, angle) T(-pivot); initRotation's angle generalised to V views."""
    for v, V in ((1, 12), (6, 12), (7, 12), (11, 12), (13, 24)):
        ang = mvr.turntable_view_angle(v, V)
        if V == 12:
            assert ang == pytest.approx(((-v) if v < 7 else (12 - v)) * np.pi / 6, abs=1e-15)
        T = mvr.turntable_rotation(synth.PIVOT, synth.AXIS, ang)
        assert np.abs(T - synth.rotation_about_axis(ang, axis=synth.AXIS, pivot=synth.PIVOT)).max() < 1e-12


def test_pairwise_align_equals_context_align(mvr, ctx, synth, seq):
    V, n, views, poses, init = seq
    reg = mvr.Registrator(0, 1)
    p = mvr.default_params(max_iterations=8, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    guess = (np.linalg.inv(init[0]) @ init[1]).astype(np.float32)
    r = reg.pairwise_align(views[1], views[0], p, guess=guess)
    ctx.set_target(views[0])
    ctx.set_source(views[1])
    c = ctx.icp_align(p, guess=guess, n_source=n)
    assert np.array_equal(r["final"], c["final"]) and r["n_corr"] == c["n_corr"] and r["mse"] == c["mse"]
    reg.close()


def test_ring_register_matches_oracle_and_is_stream_and_shard_invariant(mvr, orc, synth, seq):
    V, n, views, poses, init = seq
    icp = mvr.default_params(max_iterations=10, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=2, mode=mvr.RING_PAIRS, loop_closure=1)
    reg1 = mvr.Registrator(0, 1)
    abs1, rep1 = reg1.register_turntable(views, tp, init_poses=init)
    # oracle: every ring pair, two aligns, the second continuing from the first
    op = orc.make_params(max_iterations=10, max_dist=4.0, reciprocal=True, fixed_iterations=True)
    ext = float(np.abs(views[0][:, :3] - views[0][:, :3].mean(axis=0)).max())
    for p in range(V):
        t, s = p, (p + 1) % V
        g = (np.linalg.inv(init[t]) @ init[s]).astype(np.float32)
        o = orc.icp_align(views[s], views[t], op, guess=g)
        o = orc.icp_align(views[s], views[t], op, guess=o["final"])
        assert rep1[p]["source_view"] == s and rep1[p]["target_view"] == t and rep1[p]["status"] == 0
        assert rep1[p]["n_corr"] == o["n_corr"]
        assert_pose_close(rep1[p]["pose"], o["final"], scale=max(1.0, float(np.linalg.norm(o["final"][:3, 3]))))
    # the relaxed absolute poses stay near the ground truth (60-degree steps of 8000-point scans: ICP itself is
    # only good to ~1e-2 rad here) and, unlike the plain chain, close the ring
    for v in range(V):
        truth = np.linalg.inv(poses[0]) @ poses[v]
        assert rot_angle(abs1[v], truth) < 5e-2
    radius = 0.5 * float((views[0][:, :3].max(axis=0) - views[0][:, :3].min(axis=0)).max())
    chain = mvr.ring_close([r["pose"] for r in rep1], relax=False)
    gap_chain = rot_angle(chain[V - 1].astype(np.float64) @ rep1[V - 1]["pose"].astype(np.float64), np.eye(4))
    gap_relaxed = rot_angle(abs1[V - 1].astype(np.float64) @ rep1[V - 1]["pose"].astype(np.float64), abs1[0])
    assert gap_relaxed < 0.5 * gap_chain + 1e-6
    # several streams (host threads + GPU contexts) and sharded pair ranges give bit-identical reports
    reg3 = mvr.Registrator(0, 3)
    abs3, rep3 = reg3.register_turntable(views, tp, init_poses=init)
    for p in range(V):
        assert np.array_equal(rep1[p]["pose"], rep3[p]["pose"]) and rep1[p]["n_corr"] == rep3[p]["n_corr"]
    for v in range(V):
        assert np.array_equal(abs1[v], abs3[v])
    got = {}
    for lo, hi in ((0, 2), (2, 5), (5, 6)):
        tps = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=2, mode=mvr.RING_PAIRS, loop_closure=1,
                                   pair_begin=lo, pair_end=hi)
        need = {q % V for p in range(lo, hi) for q in (p, p + 1)}
        _, reps = reg3.register_turntable([views[v] if v in need else None for v in range(V)], tps, init_poses=init)
        for p in range(lo, hi):
            got[p] = reps[p]
    rel = [got[p]["pose"] for p in range(V)]
    for p in range(V):
        assert np.array_equal(rel[p], rep1[p]["pose"])
    closed = mvr.ring_close(rel, [got[p]["n_corr"] for p in range(V)], relax=True, iterations=16, centre=synth.PIVOT, rot_scale=radius)
    for v in range(V):
        assert np.array_equal(closed[v], abs1[v])
    reg1.close()
    reg3.close()


def _oracle_accumulate(orc, views, init, order, op, repeats):
    """The reference's model-growing loop (mvr/src/registrator.cpp:562-577 / 909-983) on the CPU oracle."""
    pose = [np.array(T, dtype=np.float64) for T in init]
    model = orc.apply_pose_double(views[0], pose[0])
    ncorr = {}
    for v in order:
        src = orc.apply_pose_double(views[v], pose[v])
        for _ in range(repeats):
            o = orc.icp_align(src, model, op)
            pose[v] = o["final"].astype(np.float64) @ pose[v]
            src = o["cloud"]
            ncorr[v] = o["n_corr"]
        model = np.concatenate([model, src])
    return pose, ncorr


def test_automatic_registration_matches_oracle(mvr, orc, synth, seq):
    V, n, views, poses, init = seq
    icp = mvr.default_params(max_iterations=6, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=2, mode=mvr.ACCUMULATE)
    reg = mvr.Registrator(0, 1)
    got, reps = reg.register_turntable(views[:4], tp, init_poses=init[:4])
    op = orc.make_params(max_iterations=6, max_dist=4.0, reciprocal=True, fixed_iterations=True)
    want, _ = _oracle_accumulate(orc, views[:4], init[:4], [1, 2, 3], op, repeats=2)
    for v in range(4):
        assert_pose_close(got[v], want[v], scale=max(1.0, float(np.linalg.norm(want[v][:3, 3]))))
    assert [r["source_view"] for r in reps] == [1, 2, 3] and all(r["fitness"] > 0 for r in reps)
    reg.close()


def test_automatic_registration_axis_follows_every_view_and_logs_every_repeat(mvr, orc, synth, seq):
    """automaticRegistration as kept from the reference (mvr/src/registrator.cpp:746-842, 877-990): a view gets its turntable
    pose when its turn comes, from the axis as refined (refineAxis, :836 / :986) over the views registered so far; the fitness
    score is logged after EVERY repeat (:923-925).  Replayed here with the oracle's ICP and the host's refine_axis."""
    V, n, views, poses, _ = seq
    V = 4
    icp = mvr.default_params(max_iterations=5, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=2, mode=mvr.ACCUMULATE)
    reg = mvr.Registrator(0, 1)
    got, reps = reg.register_turntable(views[:V], tp)          # no initial poses: PointCloud::initRotation at each view's turn
    log = reg.fitness_log()
    reg.close()
    assert [(v, r) for v, r, _ in log] == [(v, r) for v in range(1, V) for r in range(2)]
    op = orc.make_params(max_iterations=5, max_dist=4.0, reciprocal=True, fixed_iterations=True)
    pivot, axis = np.array(synth.PIVOT, dtype=np.float64), np.array(synth.AXIS, dtype=np.float64)
    pose = [np.eye(4) for _ in range(V)]
    model = orc.apply_pose_double(views[0], pose[0])
    scores = []
    for v in range(1, V):
        pose[v] = mvr.turntable_rotation(pivot, axis, mvr.turntable_view_angle(v, V))
        src = orc.apply_pose_double(views[v], pose[v])
        for r in range(2):
            o = orc.icp_align(src, model, op)
            pose[v] = o["final"].astype(np.float64) @ pose[v]
            src = o["cloud"]
            scores.append(orc.fitness_score(src, model))
        model = np.concatenate([model, src])
        pivot, axis = mvr.refine_axis([pose[k] for k in range(1, v + 1)], pivot, axis)
    for v in range(V):
        assert_pose_close(got[v], pose[v], scale=max(1.0, float(np.linalg.norm(pose[v][:3, 3]))))
    for (_, _, f), want in zip(log, scores):
        assert abs(f - want) <= 1e-6 * want   # the replay refines the axis from float32 poses (mvr_refine_axis), the driver from doubles


def test_registration_icp_order_and_reference_settings(mvr, orc, synth, seq):
    """registrationICP: order 1, V-1, 2, V-2, ..., middle; transEps 1e-6, fitEps 64 => one iteration per align."""
    V, n, views, poses, init = seq
    icp = mvr.default_params(max_iterations=2**31 - 1, max_dist=4.0)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr.ICP_ORDER)
    reg = mvr.Registrator(0, 1)
    got, reps = reg.register_turntable(views, tp, init_poses=init)
    assert [r["source_view"] for r in reps] == [1, 5, 2, 4, 3]
    assert all(r["iterations"] == 1 for r in reps)
    op = orc.make_params(max_iterations=2**31 - 1, max_dist=4.0, reciprocal=True, transformation_epsilon=1e-6, euclidean_fitness_epsilon=64.0)
    want, _ = _oracle_accumulate(orc, views, init, [1, 5, 2, 4, 3], op, repeats=1)
    for v in range(V):
        assert_pose_close(got[v], want[v], scale=max(1.0, float(np.linalg.norm(want[v][:3, 3]))))
    reg.close()


def test_registration_lum_reduces_ring_error(mvr, synth, seq):
    """registrationLUM (mvr/src/registrator.cpp:611-678): every outer loop re-derives the ring's reciprocal
    correspondences and relaxes all poses at once; the error against the true poses falls steadily."""
    V, n, views, poses, init = seq
    before = max(rot_angle(np.linalg.inv(init[0]) @ init[v], np.linalg.inv(poses[0]) @ poses[v]) for v in range(V))
    errs = []
    for loops in (1, 4, 12):
        icp = mvr.default_params(max_iterations=16 * loops, max_dist=4.0)
        tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, mode=mvr.LUM)
        reg = mvr.Registrator(0, 2)
        got, _ = reg.register_turntable(views, tp, init_poses=init)
        reg.close()
        errs.append(max(rot_angle(np.linalg.inv(got[0]) @ got[v], np.linalg.inv(poses[0]) @ poses[v]) for v in range(V)))
    assert before > 0.02 and errs[1] > errs[2] and errs[2] < 0.6 * before and errs[0] < 1.05 * before


def test_registration_lum_matches_the_oracle_loop(mvr, orc, synth, seq):
    """registrationLUM on the GPU (posing, reciprocal correspondences and their moments on the device, pcl::registration::LUM's
    sweeps on the moments) against the reference's loop restated on the CPU (oracle/lum_oracle.registration_lum: the oracle's
    correspondences, PCL's computeEdge over the point lists): poses after 1 and 3 outer loops within the north star's bars
    (1e-5 rad, 1e-6 relative translation); measured gap reported in the assertion message."""
    import lum_oracle
    V, n, views, poses, init = seq
    for loops in (1, 3):
        icp = mvr.default_params(max_iterations=16 * loops, max_dist=4.0)
        tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, mode=mvr.LUM)
        reg = mvr.Registrator(0, 1)
        got, _ = reg.register_turntable(views, tp, init_poses=init)
        reg.close()
        P = lum_oracle.registration_lum(views, init, 16 * loops, 4.0, orc.correspondences, orc.apply_pose_double)
        for v in range(V):
            ang = rot_angle(got[v], P[v])
            dt = np.linalg.norm(got[v][:3, 3].astype(np.float64) - P[v][:3, 3]) / max(np.linalg.norm(P[v][:3, 3]), 1e-12)
            # got is float32 (the ABI's pose type): 6e-8 relative on its own
            assert ang < 1e-5 and dt < 1e-6, "view %d after %d loop(s): %.3g rad, %.3g relative translation" % (v, loops, ang, dt)


def test_merge_registered_matches_reference_arithmetic(mvr, synth):
    """Registrator::saveRegisteredPoints (mvr/src/registrator.cpp:344-383): registered views posed (double math narrowed
    to float) and concatenated; normals go through the whole matrix like the reference's preMult, or are only rotated."""
    rng = np.random.default_rng(12)
    views, poses, reg = [], [], [1, 0, 1, 1]
    for v in range(4):
        n = [1000, 500, 0, 777][v]
        a = np.zeros(n, dtype=mvr.RICH_POINT)
        for k in ("x", "y", "z", "normal_x", "normal_y", "normal_z", "curvature"):
            a[k] = rng.normal(size=n).astype(np.float32) * (100.0 if k in "xyz" else 1.0)
        a["r"], a["g"], a["b"] = rng.integers(0, 256, n), rng.integers(0, 256, n), rng.integers(0, 256, n)
        views.append(a)
        poses.append(synth.view_pose(v, 12) @ synth.perturbation())
    ctx = mvr.Context(0)
    for full in (True, False):
        got = ctx.merge_registered(views, poses, registered=reg, full_matrix_normals=full)
        want = []
        for v in range(4):
            if not reg[v]:
                continue
            a = views[v].copy()
            P = np.stack([a["x"], a["y"], a["z"]], axis=1).astype(np.float64)
            N = np.stack([a["normal_x"], a["normal_y"], a["normal_z"]], axis=1).astype(np.float64)
            R, t = poses[v][:3, :3], poses[v][:3, 3]
            # the same left-to-right sum the kernel evaluates: ((m0 x + m4 y) + m8 z) + m12
            Pp = ((R[:, 0] * P[:, :1] + R[:, 1] * P[:, 1:2]) + R[:, 2] * P[:, 2:3]) + t
            Np = ((R[:, 0] * N[:, :1] + R[:, 1] * N[:, 1:2]) + R[:, 2] * N[:, 2:3]) + (t if full else 0.0 * t)
            a["x"], a["y"], a["z"] = Pp[:, 0].astype(np.float32), Pp[:, 1].astype(np.float32), Pp[:, 2].astype(np.float32)
            a["normal_x"], a["normal_y"], a["normal_z"] = Np[:, 0].astype(np.float32), Np[:, 1].astype(np.float32), Np[:, 2].astype(np.float32)
            want.append(a)
        want = np.concatenate(want)
        assert len(got) == len(want) == 1777
        for k in ("r", "g", "b", "curvature"):
            assert np.array_equal(got[k], want[k])
        for k in ("x", "y", "z", "normal_x", "normal_y", "normal_z"):
            np.testing.assert_allclose(got[k], want[k], rtol=0, atol=np.abs(want[k]).max() * 2.0 ** -23)   # within 1 float ulp (device FMA contraction is off)
    assert len(ctx.merge_registered(views, poses, registered=[0, 0, 0, 0])) == 0
    ctx.close()


def test_compute_error_matches_oracle(mvr, orc, synth, seq):
    """Registrator::computeError (mvr/src/registrator.cpp:466-515): per neighbouring pair the reciprocal correspondences of
    the POSED clouds; counts equal the oracle's, mean squared distances to float accumulation accuracy."""
    V, n, views, poses, init = seq
    reg = mvr.Registrator(0, 1)
    got = reg.compute_error(views, init, 4.0)
    assert len(got) == V
    for i, (cnt, msd) in enumerate(got):
        j = (i + 1) % V
        a = orc.apply_pose_double(views[i], init[i])
        b = orc.apply_pose_double(views[j], init[j])
        q, m, d2 = orc.correspondences(a, b, 4.0, True)
        assert cnt == len(q)
        assert abs(msd - float(d2.astype(np.float64).mean())) <= 1e-9 * msd
    reg.close()


def test_multi_gpu_entry_point_matches_the_single_gpu_driver(mvr, synth):
    """mvr_register_turntable_multi (one host thread per GPU, ncclAllGather of the pair records, host loop closure) returns,
    whatever the number of GPUs, bit for bit the pair poses of mvr_register_turntable on one GPU."""
    import torch
    import mvr_b200.ring as ring
    V, n = 8, 20_000
    views, poses = synth.turntable_sequence(V, n)
    E = synth.perturbation()
    init = [(poses[v] @ E) if v % 2 else poses[v].copy() for v in range(V)]
    icp = mvr.default_params(max_iterations=10, max_dist=4.0, reciprocal=1, fixed_iterations=1)
    tp = mvr.turntable_params(pivot=synth.PIVOT, axis=synth.AXIS, icp=icp, repeat_times=1, mode=mvr.RING_PAIRS, loop_closure=1, lum_iterations=16)
    reg = mvr.Registrator(0, 1)
    one_poses, one_reps = reg.register_turntable(views, tp, init_poses=init)
    reg.close()
    want = ring.pose_checksum(ring.pack_reports(one_reps, 0, V))
    counts = [1] + [g for g in (2, 4, 8) if g <= torch.cuda.device_count()]
    for g in counts:
        m = mvr.MultiRegistrator(range(g))
        got_poses, reps, recs, ms = m.register_turntable(views, tp, init_poses=init)
        assert ring.pose_checksum(recs) == want, "%d GPUs: pair records differ from the single-GPU run" % g
        for a, b in zip(reps, one_reps):
            assert np.array_equal(a["pose"], b["pose"]) and a["n_corr"] == b["n_corr"] and a["mse"] == b["mse"] and a["nn_queries"] == b["nn_queries"]
        for a, b in zip(got_poses, one_poses):
            assert np.array_equal(a, b)
        # resident views: the same again without the upload
        m.upload(views, init)
        got2, reps2, recs2, _ = m.register_turntable(views, tp, init_poses=init, use_resident=True)
        assert ring.pose_checksum(recs2) == want
        assert len(ms) == g and all(x > 0 for x in ms)
        m.close()


def test_denoise_matches_the_oracle(mvr, orc, synth):
    """PointCloud::denoise (mvr/src/point_cloud.cpp:423-466): kept indices, in the reference's output order, equal the oracle's
    (scipy neighbour pairs + connected components) -- a scan plus isolated specks, 16-byte and 48-byte records, thresholds."""
    rng = np.random.default_rng(21)
    scan, _ = synth.turntable_view(2, 12, 40_000)
    lo, hi = scan[:, :3].min(axis=0), scan[:, :3].max(axis=0)
    specks = []
    for k in range(60):   # small clusters floating off the surface: 1 .. 14 points each
        c = lo + rng.random(3) * (hi - lo) + np.array([0, 0, 60.0])
        specks.append(c + rng.normal(size=(1 + k % 14, 3)) * 0.4)
    specks = np.concatenate(specks).astype(np.float32)
    pts = np.ones((len(scan) + len(specks), 4), dtype=np.float32)
    pts[:len(scan), :3] = scan[:, :3]; pts[len(scan):, :3] = specks
    pts = pts[rng.permutation(len(pts))]
    c = mvr.Context(0)
    for thr, length in ((10, 2.5), (3, 2.5), (10, 1.0), (1, 0.5)):
        keep, noise = c.denoise(pts, thr, length)
        okeep, onoise = orc.denoise(pts, thr, length)
        assert noise == onoise and np.array_equal(keep, okeep), "segment_threshold %d, triangle_length %g" % (thr, length)
    keep, noise = c.denoise(pts, 10, 2.5)
    assert 0 < noise < 1000 and len(keep) + noise == len(pts)
    # PointXYZRGBNormal records, a non-finite point, the empty cloud
    rich = np.zeros(len(pts), dtype=mvr.RICH_POINT)
    rich["x"], rich["y"], rich["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    rich["x"][5] = np.nan
    k2, n2 = c.denoise(rich, 10, 2.5)
    p2 = pts.copy(); p2[5, 0] = np.nan
    ok2, on2 = orc.denoise(p2, 10, 2.5)
    assert n2 == on2 and np.array_equal(k2, ok2) and 5 not in k2
    k0, n0 = c.denoise(pts[:0], 10, 2.5)
    assert len(k0) == 0 and n0 == 0
    c.close()


def test_pcl_adapter_runs_icp_through_the_pcl_shaped_classes(mvr, tmp_path):
    """The adapter header's GpuICP / GpuCorrespondenceEstimation (include/mvr_pcl_adapter.hpp) drive a small align on the GPU."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_abi_and_host import _build_adapter_check
    exe = _build_adapter_check(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    # the reference's settings (fitness epsilon 64 = relative MSE threshold) end every align after one iteration
    assert out.returncode == 0 and out.stdout.startswith("gpu ok: iterations 1"), out.stdout + out.stderr
    assert "corr 2000" in out.stdout
    tx = float(out.stdout.split("tx")[1].split()[0])
    assert 0.2 < tx < 0.4      # the target is the source shifted by 0.3 along x
