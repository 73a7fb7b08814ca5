"""ctypes front-end of oracle/liboracle.so (strict.c).  TEST INFRASTRUCTURE ONLY.

All functions take / return numpy arrays (fp32 C-contiguous, int64 indices) shaped like
the reference's tensors.  See strict.c for the reference file:line each one follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only, a second or two)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "strict.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_chamfer.restype = C.c_double
        _lib.orc_index_points.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def square_distance(src, dst):
    src, ps = _f(src); dst, pd = _f(dst)
    B, N, _ = src.shape; M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().orc_square_distance(ps, pd, B, N, M, out.ctypes.data_as(_f32p))
    return out


FORM_KNN = 0      # expanded form, ref norm added first   (Utils/Layers.py:50-53)
FORM_QFIRST = 1   # expanded form, query norm added first (ball query / three-NN call sites)
FORM_DIRECT = 2   # direct (q-r)^2, FMA accumulation      (pytorch3d.knn_points restated)


def knn(ref, qry, k, form):
    """k nearest refs for each query -> (idx int64 [B,S,k], dist fp32 [B,S,k])."""
    ref, pr = _f(ref); qry, pq = _f(qry)
    B, N, _ = ref.shape; S = qry.shape[1]
    assert 1 <= k <= N
    idx = np.empty((B, S, k), np.int64); dist = np.empty((B, S, k), np.float32)
    lib().orc_knn(pr, pq, B, N, S, int(k), int(form), idx.ctypes.data_as(_i64p), dist.ctypes.data_as(_f32p))
    return idx, dist


def knn_point(nsample, xyz, new_xyz):
    return knn(xyz, new_xyz, nsample, FORM_KNN)[0]


def three_nn(unknown, known):
    idx, dist = knn(known, unknown, 3, FORM_QFIRST)
    return dist, idx


def knn_points(p1, p2, K):
    idx, dist = knn(p2, p1, K, FORM_DIRECT)
    return dist, idx


def radius_sq(radius):
    """fp32 rounding of the python-double radius**2 (Utils/Pointnet2Utils.py:103)."""
    return np.float32(float(radius) ** 2)


def query_ball_point(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz); new_xyz, pq = _f(new_xyz)
    B, N, _ = xyz.shape; S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    lib().orc_ball_query(C.c_float(radius_sq(radius)), int(nsample), px, pq, B, N, S, out.ctypes.data_as(_i64p))
    return out


def farthest_point_sample(xyz, npoint, start):
    xyz, px = _f(xyz); start, pst = _i(start)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int64)
    lib().orc_fps(px, B, N, int(npoint), pst, out.ctypes.data_as(_i64p))
    return out


def index_points(points, idx):
    points, pp = _f(points); idx, pi = _i(idx)
    B, N, Cc = points.shape
    R = int(np.prod(idx.shape[1:]))
    out = np.empty(idx.shape + (Cc,), np.float32)
    rc = lib().orc_index_points(pp, pi, B, N, Cc, C.c_int64(R), out.ctypes.data_as(_f32p))
    if rc != 0:
        raise IndexError("index out of range in index_points oracle")
    return out


def group_points(xyz, new_xyz, feat, idx, xyz_first=True):
    """The grouping tail the fused kernel replaces, restated step by step in numpy (test infrastructure):
      Utils/Layers.py:57-66      grouped = index_points(points, ind) - new_points.view(B,S,1,C)
                                 cat([grouped, index_points(features, ind)], -1).permute(0,3,2,1)   (xyz_first)
      Utils/Pointnet2Utils.py:243-253   cat([index_points(points, idx), grouped_xyz], -1).permute(0,3,2,1)
    xyz [B,N,3], new_xyz [B,S,3], feat [B,N,D] or None, idx [B,S,K] -> [B,3+D,K,S] float32."""
    xyz = np.ascontiguousarray(xyz, np.float32); new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    B, S = new_xyz.shape[0], new_xyz.shape[1]
    rel = (index_points(xyz, idx) - new_xyz.reshape(B, S, 1, 3)).astype(np.float32)         # one fp32 subtraction
    parts = [rel]
    if feat is not None and feat.shape[2] > 0:
        g = index_points(feat, idx)
        parts = [rel, g] if xyz_first else [g, rel]
    return np.ascontiguousarray(np.concatenate(parts, axis=-1).transpose(0, 3, 2, 1))


def three_weights(dist, variant):
    dist, pd = _f(dist)
    w = np.empty_like(dist)
    lib().orc_three_weights(pd, C.c_int64(dist.size // 3), int(variant), w.ctypes.data_as(_f32p))
    return w


def three_interpolate(feat, idx, weight):
    """feat [B,S,C], idx [B,N,3], weight [B,N,3] -> [B,N,C]."""
    feat, pf = _f(feat); idx, pi = _i(idx); weight, pw = _f(weight)
    B, S, Cc = feat.shape; N = idx.shape[1]
    out = np.empty((B, N, Cc), np.float32)
    lib().orc_three_interpolate(pf, pi, pw, B, S, N, Cc, out.ctypes.data_as(_f32p))
    return out


def nearest(a, b):
    """for every a_i the nearest b_j, direct form -> (min fp32 [B,N], arg int64 [B,N])."""
    a, pa = _f(a); b, pb = _f(b)
    B, N, _ = a.shape; M = b.shape[1]
    mn = np.empty((B, N), np.float32); arg = np.empty((B, N), np.int64)
    lib().orc_nearest(pa, pb, B, N, M, mn.ctypes.data_as(_f32p), arg.ctypes.data_as(_i64p))
    return mn, arg


def chamfer(x, y):
    """pytorch3d chamfer_distance defaults on [B,N,3],[B,M,3] -> (loss, dx, ix, dy, iy)."""
    x, px = _f(x); y, py = _f(y)
    B, N, _ = x.shape; M = y.shape[1]
    dx = np.empty((B, N), np.float32); ix = np.empty((B, N), np.int64)
    dy = np.empty((B, M), np.float32); iy = np.empty((B, M), np.int64)
    loss = lib().orc_chamfer(px, py, B, N, M, dx.ctypes.data_as(_f32p), ix.ctypes.data_as(_i64p),
                             dy.ctypes.data_as(_f32p), iy.ctypes.data_as(_i64p))
    return float(loss), dx, ix, dy, iy
