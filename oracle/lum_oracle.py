"""lum_oracle.py -- CPU restatement of pcl::registration::LUM as the reference drives it (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu legs may import this; the product never does).

Reference call sites: mvr/src/registrator.cpp:627-663 (lum.addPointCloud, lum.setCorrespondences, lum.setMaxIterations(16),
lum.compute(), lum.getTransformation(i)).  The arithmetic lives in PCL (un-vendored, un-pinned: mvr/CMakeLists.txt:10); this
file restates pcl/registration/impl/lum.hpp of PCL 1.7.x FROM MEMORY -- parity unpinned: no PCL build
and no golden vectors exist here.  What is checked instead (tests/test_oracle.py): the restated edge model is the exact
linearisation it claims to be (finite differences), compute() removes a synthetic drift from a ring, and the product's
moment-based implementation reproduces this point-level one to rounding.

  vertex       : a cloud + a 6-vector pose (x, y, z, roll, pitch, yaw), vertex 0 fixed at 0; pcl::getTransformation(pose) =
                 Translation(x, y, z) * Rz(yaw) * Ry(pitch) * Rx(roll)
  computeEdge  : with both clouds compounded onto their current poses, per correspondence a_k = (s + t) / 2, d_k = s - t;
                 M'M (6 x 6) and M'Z (6) of the linearised pose-difference model  d_k ~ M_k D,
                     M_k = [ I | (0, -a_y, a_z; -a_z, a_x, 0; a_y, 0, -a_x) ],
                 D = (M'M)^-1 M'Z,  s^2 = sum |d_k - M_k D|^2,  cinv = M'M / s^2,  cinvd = M'Z / s^2
                 (an edge with < 3 pairs or s^2 < 1e-13 carries zero information)
  compute      : max_iterations times: all edges; dense G (6 (V - 1) square) and B from cinv / cinvd (forward edge +, backward
                 edge -); X = G^-1 B (PCL: colPivHouseholderQr); pose_v += -incidenceCorrection(pose_v)^-1 X_v; stop when the
                 summed update norm <= convergence_threshold * (V - 1) (PCL default 0: never)
PCL computes all of this in float32; here it is float64 (strictly more accurate, like the oracle's umeyama).
"""
import numpy as np


def get_transformation(pose):
    """pcl::getTransformation(x, y, z, roll, pitch, yaw) as a 4x4 (p' = T p)."""
    x, y, z, roll, pitch, yaw = [float(v) for v in pose]
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [x, y, z]
    return T


def incidence_correction(pose):
    """LUM::incidenceCorrection: maps an increment of (x, y, z, roll, pitch, yaw) to the linearised parameters D."""
    x, y, z = float(pose[0]), float(pose[1]), float(pose[2])
    cx, sx, cy, sy = np.cos(pose[3]), np.sin(pose[3]), np.cos(pose[4]), np.sin(pose[4])
    out = np.eye(6)
    out[0, 4] = y * sx - z * cx
    out[0, 5] = y * cx * cy + z * sx * cy
    out[1, 3] = z
    out[1, 4] = -x * sx
    out[1, 5] = -x * cx * cy + z * sy
    out[2, 3] = -y
    out[2, 4] = x * cx
    out[2, 5] = -x * sx * cy - y * sy
    out[3, 5] = sy
    out[4, 4] = sx
    out[4, 5] = cx * cy
    out[5, 4] = cx
    out[5, 5] = -sx * cy
    return out


def edge_from_pairs(a, d):
    """M'M, M'Z of computeEdge from the averaged points a (n x 3) and differences d (n x 3); returns (cinv, cinvd, D, ss)."""
    n = len(a)
    if n < 3:
        return np.zeros((6, 6)), np.zeros(6), np.zeros(6), 0.0
    ax, ay, az = a[:, 0], a[:, 1], a[:, 2]
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    MM = np.zeros((6, 6))
    MM[0, 4] = -ay.sum(); MM[0, 5] = az.sum()
    MM[1, 3] = -az.sum(); MM[1, 4] = ax.sum()
    MM[2, 3] = ay.sum(); MM[2, 5] = -ax.sum()
    MM[3, 4] = -(ax * az).sum(); MM[3, 5] = -(ax * ay).sum(); MM[4, 5] = -(ay * az).sum()
    MM[3, 3] = (ay * ay + az * az).sum(); MM[4, 4] = (ax * ax + ay * ay).sum(); MM[5, 5] = (ax * ax + az * az).sum()
    MM[0, 0] = MM[1, 1] = MM[2, 2] = float(n)
    MM = MM + np.triu(MM, 1).T
    MZ = np.array([dx.sum(), dy.sum(), dz.sum(), (ay * dz - az * dy).sum(), (ax * dy - ay * dx).sum(), (az * dx - ax * dz).sum()])
    D = np.linalg.solve(MM, MZ)
    ss = float(((dx - (D[0] + az * D[5] - ay * D[4])) ** 2 + (dy - (D[1] + ax * D[4] - az * D[3])) ** 2 +
                (dz - (D[2] + ay * D[3] - ax * D[5])) ** 2).sum())
    if ss < 0.0000000000001 or not np.isfinite(ss):
        return np.zeros((6, 6)), np.zeros(6), D, ss
    return MM / ss, MZ / ss, D, ss


def compute_edge(src_pts, tgt_pts, src_pose, tgt_pose, iq, im):
    """LUM::computeEdge for one edge: src_pts / tgt_pts are the vertex clouds (n x >=3), iq / im the correspondence indices."""
    Ts, Tt = get_transformation(src_pose), get_transformation(tgt_pose)
    s = src_pts[iq, :3].astype(np.float64) @ Ts[:3, :3].T + Ts[:3, 3]
    t = tgt_pts[im, :3].astype(np.float64) @ Tt[:3, :3].T + Tt[:3, 3]
    ok = np.isfinite(s).all(axis=1) & np.isfinite(t).all(axis=1)
    s, t = s[ok], t[ok]
    return edge_from_pairs(0.5 * (s + t), s - t)


def lum_compute(clouds, edges, max_iterations=5, convergence_threshold=0.0):
    """LUM::compute.  clouds: list of V point arrays; edges: list of (source, target, index_query, index_match).
    Returns (poses V x 6, transforms list of V 4x4 = lum.getTransformation(v))."""
    V = len(clouds)
    poses = np.zeros((V, 6))
    if V < 2:
        return poses, [get_transformation(p) for p in poses]
    n = 6 * (V - 1)
    for _ in range(max_iterations):
        info = {}
        for (s, t, iq, im) in edges:
            cinv, cinvd, _, _ = compute_edge(clouds[s], clouds[t], poses[s], poses[t], iq, im)
            info[(s, t)] = (cinv, cinvd)
        G = np.zeros((n, n))
        B = np.zeros(n)
        for vi in range(1, V):
            for vj in range(V):
                if (vi, vj) in info:
                    (cinv, cinvd), sign = info[(vi, vj)], 1.0
                elif (vj, vi) in info:
                    (cinv, cinvd), sign = info[(vj, vi)], -1.0
                else:
                    continue
                if vj > 0:
                    G[6 * (vi - 1):6 * vi, 6 * (vj - 1):6 * vj] = -cinv
                G[6 * (vi - 1):6 * vi, 6 * (vi - 1):6 * vi] += cinv
                B[6 * (vi - 1):6 * vi] += sign * cinvd
        X = np.linalg.lstsq(G, B, rcond=None)[0]   # PCL: G.colPivHouseholderQr().solve(B)
        total = 0.0
        for vi in range(1, V):
            diff = -np.linalg.solve(incidence_correction(poses[vi]), X[6 * (vi - 1):6 * vi])
            total += float(np.linalg.norm(diff))
            poses[vi] = poses[vi] + diff
        if total <= convergence_threshold * (V - 1):
            break
    return poses, [get_transformation(p) for p in poses]


def registration_lum(views, init_poses, max_iterations, max_distance, correspondences, transform_points):
    """The reference's registrationLUM loop (mvr/src/registrator.cpp:611-663) around lum_compute: max(1, max_iterations / 16)
    outer loops of { pose every view (getTransformedPoints: double math narrowed to float), reciprocal correspondences of the
    ring edges i -> (i + 1) % V, 16 LUM sweeps, pose_v <- T_v pose_v }.  `correspondences(src, tgt, max_dist, reciprocal)` and
    `transform_points(pts, pose)` come from the main oracle.  Returns the list of V poses (4x4 float64)."""
    V = len(views)
    P = [np.array(T, dtype=np.float64) for T in init_poses]
    for _ in range(max(1, max_iterations // 16)):
        clouds = [transform_points(views[v], P[v]) for v in range(V)]
        edges = []
        for i in range(V if V > 2 else 1):
            s, t = i, (i + 1) % V
            iq, im, _ = correspondences(clouds[s], clouds[t], max_distance, True)
            edges.append((s, t, iq, im))
        _, T = lum_compute(clouds, edges, max_iterations=16)
        P = [T[v] @ P[v] for v in range(V)]
    return P
