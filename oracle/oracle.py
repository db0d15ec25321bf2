"""ctypes binding of the CPU oracle (oracle/mvr_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (multi-view-registration_b200/) never does.  PARITY UNPINNED: see the header
of mvr_oracle.cpp -- the reference's arithmetic lives in an un-vendored, un-pinned PCL that cannot be
built here and the reference has no tests or golden vectors.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmvr_oracle.so")


class IcpParams(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int),
        ("max_correspondence_distance", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("euclidean_fitness_epsilon", C.c_double),
        ("use_reciprocal", C.c_int),
        ("estimator", C.c_int),
        ("fixed_iterations", C.c_int),
        ("min_correspondences", C.c_int),
    ]


class IcpReport(C.Structure):
    _fields_ = [
        ("iterations", C.c_int),
        ("converged", C.c_int),
        ("reason", C.c_int),
        ("n_correspondences", C.c_int),
        ("mse", C.c_double),
        ("nn_queries", C.c_ulonglong),
    ]


class IterRecord(C.Structure):
    _fields_ = [("iteration", C.c_int), ("n_corr", C.c_int), ("mse", C.c_double), ("delta", C.c_double * 16)]


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "mvr_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        up = C.POINTER(C.c_uint32)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_transform.argtypes = [fp, C.c_int, fp, fp]
        L.orc_apply_pose_double.argtypes = [fp, C.c_int, C.c_int, dp, fp]
        L.orc_nn_brute.argtypes = [fp, C.c_int, fp, C.c_int, ip, fp]
        L.orc_nn_kdtree.argtypes = [fp, C.c_int, fp, C.c_int, ip, fp]
        L.orc_kdtree_build.argtypes = [fp, C.c_int]
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_free.argtypes = [C.c_void_p]
        L.orc_kdtree_query.argtypes = [C.c_void_p, fp, C.c_int, ip, fp]
        L.orc_correspondences.argtypes = [fp, C.c_int, fp, C.c_int, C.c_double, C.c_int, ip, ip, fp]
        L.orc_correspondences.restype = C.c_int
        L.orc_estimate_rigid_svd.argtypes = [fp, fp, ip, ip, C.c_int, dp]
        L.orc_estimate_rigid_svd.restype = C.c_int
        L.orc_estimate_point_to_plane.argtypes = [fp, fp, fp, ip, ip, C.c_int, dp]
        L.orc_estimate_point_to_plane.restype = C.c_int
        L.orc_svd3.argtypes = [dp, dp, dp, dp]
        L.orc_icp_align.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.POINTER(IcpParams), fp, fp, fp,
                                    C.POINTER(IcpReport), C.POINTER(IterRecord), C.c_int, ip]
        L.orc_icp_align.restype = C.c_int
        L.orc_fitness_score.argtypes = [fp, C.c_int, fp, C.c_int, C.c_double]
        L.orc_fitness_score.restype = C.c_double
        L.orc_morton_keys.argtypes = [fp, C.c_int, fp, C.c_float, C.c_int, up]
        L.orc_stable_sort_perm.argtypes = [up, C.c_int, ip]
        L.orc_cell_table.argtypes = [up, C.c_int, C.c_int, up]
        L.orc_estimate_normals.argtypes = [fp, C.c_int, C.c_int, fp, fp, ip]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "points are n x 4 float32 (PCL PointXYZ layout)"
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def transform(pts, M):
    """M: 4x4 float32 (column-vector convention, numpy row-major) -> transformed n x 4."""
    pts = _pts(pts)
    Mc = np.ascontiguousarray(np.asarray(M, dtype=np.float32).T)  # column-major flat
    out = np.empty_like(pts)
    lib().orc_transform(_f(pts), len(pts), _f(Mc), _f(out))
    return out


def apply_pose_double(pts, M):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    Mc = np.ascontiguousarray(np.asarray(M, dtype=np.float64).T)
    out = np.empty((len(pts), 4), dtype=np.float32)
    lib().orc_apply_pose_double(_f(pts), len(pts), pts.shape[1], _d(Mc), _f(out))
    return out


def nn_brute(tgt, q):
    tgt, q = _pts(tgt), _pts(q)
    idx = np.empty(len(q), dtype=np.int32)
    d2 = np.empty(len(q), dtype=np.float32)
    lib().orc_nn_brute(_f(tgt), len(tgt), _f(q), len(q), _i(idx), _f(d2))
    return idx, d2


def nn_kdtree(tgt, q):
    tgt, q = _pts(tgt), _pts(q)
    idx = np.empty(len(q), dtype=np.int32)
    d2 = np.empty(len(q), dtype=np.float32)
    lib().orc_nn_kdtree(_f(tgt), len(tgt), _f(q), len(q), _i(idx), _f(d2))
    return idx, d2


class KdTree:
    def __init__(self, tgt):
        self._tgt = _pts(tgt)
        self._h = lib().orc_kdtree_build(_f(self._tgt), len(self._tgt))

    def query(self, q):
        q = _pts(q)
        idx = np.empty(len(q), dtype=np.int32)
        d2 = np.empty(len(q), dtype=np.float32)
        lib().orc_kdtree_query(self._h, _f(q), len(q), _i(idx), _f(d2))
        return idx, d2

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_kdtree_free(self._h)
            self._h = None


def correspondences(src, tgt, max_dist, reciprocal):
    src, tgt = _pts(src), _pts(tgt)
    n = len(src)
    q = np.empty(n, dtype=np.int32)
    m = np.empty(n, dtype=np.int32)
    d2 = np.empty(n, dtype=np.float32)
    c = lib().orc_correspondences(_f(src), n, _f(tgt), len(tgt), float(max_dist), int(bool(reciprocal)), _i(q), _i(m), _f(d2))
    return q[:c].copy(), m[:c].copy(), d2[:c].copy()


def estimate_rigid_svd(src, tgt, q, m):
    src, tgt = _pts(src), _pts(tgt)
    q = np.ascontiguousarray(q, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    T = np.empty(16, dtype=np.float64)
    rc = lib().orc_estimate_rigid_svd(_f(src), _f(tgt), _i(q), _i(m), len(q), _d(T))
    assert rc == 0
    return T.reshape(4, 4).T.copy()


def estimate_point_to_plane(src, tgt, tgt_normals, q, m):
    src, tgt, tgt_normals = _pts(src), _pts(tgt), _pts(tgt_normals)
    q = np.ascontiguousarray(q, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    T = np.empty(16, dtype=np.float64)
    rc = lib().orc_estimate_point_to_plane(_f(src), _f(tgt), _f(tgt_normals), _i(q), _i(m), len(q), _d(T))
    return rc, T.reshape(4, 4).T.copy()


def svd3(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    U = np.empty((3, 3)); s = np.empty(3); V = np.empty((3, 3))
    lib().orc_svd3(_d(A), _d(U), _d(s), _d(V))
    return U, s, V


def make_params(max_iterations=10, max_dist=None, transformation_epsilon=0.0, euclidean_fitness_epsilon=None,
                reciprocal=True, estimator=0, fixed_iterations=False, min_correspondences=3):
    import sys
    p = IcpParams()
    p.max_iterations = int(max_iterations)
    p.max_correspondence_distance = float(max_dist) if max_dist is not None else float(np.sqrt(sys.float_info.max))
    p.transformation_epsilon = float(transformation_epsilon)
    p.euclidean_fitness_epsilon = float(euclidean_fitness_epsilon) if euclidean_fitness_epsilon is not None else -sys.float_info.max
    p.use_reciprocal = int(bool(reciprocal))
    p.estimator = int(estimator)
    p.fixed_iterations = int(bool(fixed_iterations))
    p.min_correspondences = int(min_correspondences)
    return p


def icp_align(src, tgt, params, guess=None, tgt_normals=None, max_log=256):
    """Returns dict(final 4x4 float32, cloud n x 4, report, log list)."""
    src, tgt = _pts(src), _pts(tgt)
    g = np.eye(4, dtype=np.float32) if guess is None else np.asarray(guess, dtype=np.float32)
    gc = np.ascontiguousarray(g.T)
    fin = np.empty(16, dtype=np.float32)
    out = np.empty_like(src)
    rep = IcpReport()
    log = (IterRecord * max_log)()
    nlog = C.c_int(0)
    nptr = _f(_pts(tgt_normals)) if tgt_normals is not None else None
    rc = lib().orc_icp_align(_f(src), len(src), _f(tgt), len(tgt), nptr, C.byref(params), _f(gc), _f(fin), _f(out),
                             C.byref(rep), log, max_log, C.byref(nlog))
    recs = []
    for k in range(nlog.value):
        recs.append(dict(iteration=log[k].iteration, n_corr=log[k].n_corr, mse=log[k].mse,
                         delta=np.array(log[k].delta[:], dtype=np.float64).reshape(4, 4).T.copy()))
    return dict(status=rc, final=fin.reshape(4, 4).T.copy(), cloud=out,
                iterations=rep.iterations, converged=rep.converged, reason=rep.reason,
                n_corr=rep.n_correspondences, mse=rep.mse, nn_queries=int(rep.nn_queries), log=recs)


def fitness_score(cloud, tgt, max_range=None):
    import sys
    cloud, tgt = _pts(cloud), _pts(tgt)
    mr = sys.float_info.max if max_range is None else float(max_range)
    return lib().orc_fitness_score(_f(cloud), len(cloud), _f(tgt), len(tgt), mr)


def morton_keys(pts, origin, inv_cell, bits):
    pts = _pts(pts)
    o = np.ascontiguousarray(origin, dtype=np.float32)
    keys = np.empty(len(pts), dtype=np.uint32)
    lib().orc_morton_keys(_f(pts), len(pts), _f(o), C.c_float(np.float32(inv_cell)), int(bits), _u(keys))
    return keys


def stable_sort_perm(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    perm = np.empty(len(keys), dtype=np.int32)
    lib().orc_stable_sort_perm(_u(keys), len(keys), _i(perm))
    return perm


def cell_table(sorted_keys, bits):
    sorted_keys = np.ascontiguousarray(sorted_keys, dtype=np.uint32)
    start = np.empty((1 << (3 * bits)) + 1, dtype=np.uint32)
    lib().orc_cell_table(_u(sorted_keys), len(sorted_keys), int(bits), _u(start))
    return start


def estimate_normals(pts, k, viewpoint=(0.0, 0.0, 0.0), want_neighbours=False):
    pts = _pts(pts)
    vp = np.ascontiguousarray(viewpoint, dtype=np.float32)
    out = np.empty((len(pts), 4), dtype=np.float32)
    nbr = np.empty((len(pts), k), dtype=np.int32) if want_neighbours else None
    lib().orc_estimate_normals(_f(pts), len(pts), int(k), _f(vp), _f(out), _i(nbr) if nbr is not None else None)
    return (out, nbr) if want_neighbours else out


def denoise(pts, segment_threshold=10, triangle_length=2.5):
    """PointCloud::denoise (mvr/src/point_cloud.cpp:423-466) restated: connected components of the graph of point pairs at
    most triangle_length apart (the reference takes the edges from a CGAL Delaunay triangulation, :469-497; its short edges
    contain the short edges of the Euclidean minimum spanning tree, so the components are those of this radius graph),
    components smaller than segment_threshold dropped, output order = components by their smallest point index (the order
    boost::connected_components discovers them), points by index (:441-447).  Returns (kept indices, number of noise points).
    scipy does the neighbour search (cKDTree.query_pairs) and the labelling."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    p = np.ascontiguousarray(np.asarray(pts)[:, :3], dtype=np.float64)
    n = len(p)
    if n == 0:
        return np.zeros(0, dtype=np.int32), 0
    ok = np.isfinite(p).all(axis=1)
    idx = np.flatnonzero(ok)
    pairs = cKDTree(p[ok]).query_pairs(float(triangle_length), output_type="ndarray") if len(idx) else np.zeros((0, 2), dtype=np.int64)
    g = coo_matrix((np.ones(len(pairs), dtype=np.int8), (idx[pairs[:, 0]], idx[pairs[:, 1]])), shape=(n, n))
    _, label = connected_components(g, directed=False)
    size = np.bincount(label, minlength=label.max() + 1)
    first = np.full(label.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, label, np.arange(n))
    keep = np.flatnonzero(size[label] >= segment_threshold)
    order = np.lexsort((keep, first[label[keep]]))
    return keep[order].astype(np.int32), int(n - len(keep))
