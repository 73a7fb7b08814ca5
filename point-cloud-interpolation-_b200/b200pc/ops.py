"""Tensor-level entry points: torch CUDA tensors in, torch CUDA tensors out, via the C ABI.

torch is used for device memory, streams and autograd bookkeeping only; every computation is a
call into libb200pc.so on the tensor's device and torch's current stream.  CPU tensors are
rejected (no CPU fallback).  Layouts and dtypes follow the reference: fp32 point-major
[B,N,3] / [B,N,C], int64 indices, fresh contiguous outputs.
"""
import ctypes as C
import os

import torch

from . import _lib

FORM_REF_NORM_FIRST = 0
FORM_QRY_NORM_FIRST = 1
FORM_DIRECT = 2

# index_points bounds checking costs a device->host sync per call; the reference raises
# IndexError on out-of-range indices, so tests switch this on.  Off: bad rows come back as zeros.
CHECK_BOUNDS = os.environ.get("B200PC_CHECK_BOUNDS", "0") == "1"

launch_count = 0     # C-ABI compute calls issued (bench.py reports kernels from the plan below)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _prep(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("b200pc: %s is on %s; this library has no CPU path (CUDA tensors only)" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _idx64(idx, dev):
    if idx.device != dev:
        idx = idx.to(dev)
    if idx.dtype != torch.int64:
        idx = idx.long()
    return idx.contiguous()


def _workspace(nbytes, dev):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=dev)


def _bump():
    global launch_count
    launch_count += 1


# ------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """Utils/Pointnet2Utils.py:20  [B,N,3],[B,M,3] -> [B,N,M], torch-CPU rounding order."""
    src = _prep(src, "src"); dst = _prep(dst, "dst")
    B, N, _ = src.shape; M = dst.shape[1]
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().b200pc_square_distance(_ptr(src), _ptr(dst), B, N, M, _ptr(out), _stream(src.device)))
    _bump()
    return out


def knn_search(ref, qry, k, form, want_dist=False):
    """k nearest refs per query.  ref [B,N,3], qry [B,S,3] -> idx [B,S,k] int64 (, dist [B,S,k])."""
    ref = _prep(ref, "ref"); qry = _prep(qry, "qry")
    B, N, _ = ref.shape; S = qry.shape[1]
    k = int(k)
    if k > N:   # torch.topk raises RuntimeError("selected index k out of range")
        raise RuntimeError("selected index k out of range (k=%d > %d reference points)" % (k, N))
    dev = ref.device
    idx = torch.empty(B, S, k, dtype=torch.int64, device=dev)
    dist = torch.empty(B, S, k, dtype=torch.float32, device=dev) if want_dist else None
    lib = _lib.load()
    nws = lib.b200pc_search_workspace_bytes(B, N, S, k)
    ws = _workspace(nws, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200pc_knn(_ptr(ref), _ptr(qry), B, N, S, k, int(form), _ptr(idx), _ptr(dist), _ptr(ws), nws,
                                  _stream(dev)))
    _bump()
    return (idx, dist) if want_dist else idx


def ball_query(radius, nsample, xyz, new_xyz):
    """Utils/Pointnet2Utils.py:88  -> [B,S,nsample] int64 (N where the ball is empty)."""
    xyz = _prep(xyz, "xyz"); new_xyz = _prep(new_xyz, "new_xyz")
    B, N, _ = xyz.shape; S = new_xyz.shape[1]
    dev = xyz.device
    # `sqrdists > radius ** 2`: python double squared, then compared against fp32 values
    r2 = torch.tensor(float(radius) ** 2, dtype=torch.float32).item()
    idx = torch.empty(B, S, int(nsample), dtype=torch.int64, device=dev)
    lib = _lib.load()
    nws = lib.b200pc_search_workspace_bytes(B, N, S, int(nsample))
    ws = _workspace(nws, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200pc_ball_query(_ptr(xyz), _ptr(new_xyz), B, N, S, C.c_float(r2), int(nsample), _ptr(idx),
                                         _ptr(ws), nws, _stream(dev)))
    _bump()
    return idx


def fps(xyz, npoint, start, want_xyz=False):
    """Utils/Pointnet2Utils.py:64  start [B] int64 (first centroid) -> [B,npoint] int64.
    want_xyz: also return the picks' coordinates [B,npoint,3] from the same C call (Sample.forward, Utils/Layers.py:23-27)."""
    xyz = _prep(xyz, "xyz")
    B, N, _ = xyz.shape
    dev = xyz.device
    start = _idx64(start, dev)
    idx = torch.empty(B, int(npoint), dtype=torch.int64, device=dev)
    new_xyz = torch.empty(B, int(npoint), 3, dtype=torch.float32, device=dev) if want_xyz else None
    with torch.cuda.device(dev):
        if want_xyz:
            _lib.check(_lib.load().b200pc_fps_sample(_ptr(xyz), B, N, int(npoint), _ptr(start), _ptr(idx), _ptr(new_xyz), _stream(dev)))
        else:
            _lib.check(_lib.load().b200pc_fps(_ptr(xyz), B, N, int(npoint), _ptr(start), _ptr(idx), C.c_void_p(0), 0,
                                              _stream(dev)))
    _bump()
    return (idx, new_xyz) if want_xyz else idx


# ------------------------------------------------------------------------------------------
def _gather_raw(points, idx_flat, out_shape):
    B, N, Cc = points.shape
    R = idx_flat.shape[1]
    dev = points.device
    out = torch.empty(B, R, Cc, dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev) if CHECK_BOUNDS else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200pc_gather(_ptr(points), _ptr(idx_flat), B, N, Cc, R, _ptr(out), _ptr(flag),
                                             _stream(dev)))
    _bump()
    if flag is not None and int(flag.item()) != 0:
        raise IndexError("index out of range in index_points (valid range is [-%d, %d))" % (N, N))
    return out.view(*out_shape)


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx_flat, out_shape):
        ctx.save_for_backward(idx_flat)
        ctx.pshape = points.shape
        return _gather_raw(points, idx_flat, out_shape)

    @staticmethod
    def backward(ctx, gout):
        (idx_flat,) = ctx.saved_tensors
        B, N, Cc = ctx.pshape
        gout = gout.contiguous().view(B, -1, Cc)
        gpts = torch.zeros(B, N, Cc, dtype=torch.float32, device=gout.device)
        with torch.cuda.device(gout.device):
            _lib.check(_lib.load().b200pc_gather_bwd(_ptr(gout), _ptr(idx_flat), B, N, Cc, idx_flat.shape[1], _ptr(gpts),
                                                     _stream(gout.device)))
        _bump()
        return gpts, None, None


def gather(points, idx):
    """Utils/Pointnet2Utils.py:44  points [B,N,C], idx [B,...] -> [B,...,C]  (differentiable in points)."""
    points = _prep(points, "points")
    idx = _idx64(idx, points.device)
    B = points.shape[0]
    out_shape = tuple(idx.shape) + (points.shape[2],)
    if points.shape[2] == 0 or idx.numel() == 0:      # nothing to move (e.g. clouds without extra channels)
        return torch.empty(out_shape, dtype=torch.float32, device=points.device)
    idx_flat = idx.reshape(B, -1)
    if points.requires_grad and torch.is_grad_enabled():
        return _GatherFn.apply(points, idx_flat, out_shape)
    return _gather_raw(points, idx_flat, out_shape)


# ------------------------------------------------------------------------------------------
def _group_raw(xyz, new_xyz, feat, idx, xyz_first):
    B, N, _ = xyz.shape
    S, K = idx.shape[1], idx.shape[2]
    D = 0 if feat is None else feat.shape[2]
    out = torch.empty(B, 3 + D, K, S, dtype=torch.float32, device=xyz.device)
    if CHECK_BOUNDS and idx.numel() and (int(idx.min()) < -N or int(idx.max()) >= N):
        raise IndexError("index out of range in group_points (valid range is [-%d, %d))" % (N, N))
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().b200pc_group_points(_ptr(xyz), _ptr(new_xyz), _ptr(feat if D else None), _ptr(idx), B, N, S, K,
                                                   D, int(bool(xyz_first)), _ptr(out), _stream(xyz.device)))
    _bump()
    return out


class _GroupFn(torch.autograd.Function):
    """differentiable in the features (what the reference's models train through); the coordinates get their
    gradient from the same three terms the unfused graph would produce, built with torch ops on the xyz channels."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, feat, idx, xyz_first):
        ctx.save_for_backward(idx)
        ctx.shapes = (xyz.shape, new_xyz.shape, None if feat is None else feat.shape, bool(xyz_first))
        return _group_raw(xyz, new_xyz, feat, idx, xyz_first)

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        xs, cs, fs, xyz_first = ctx.shapes
        B, N, _ = xs
        S, K = idx.shape[1], idx.shape[2]
        gout = gout.contiguous()
        gxyz = gnew = gfeat = None
        D = 0 if fs is None else fs[2]
        if fs is not None and ctx.needs_input_grad[2] and D:
            gfeat = torch.zeros(B, N, D, dtype=torch.float32, device=gout.device)
            with torch.cuda.device(gout.device):
                _lib.check(_lib.load().b200pc_group_points_bwd(_ptr(gout), _ptr(idx), B, N, S, K, D, int(xyz_first),
                                                               _ptr(gfeat), _stream(gout.device)))
            _bump()
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gx = gout[:, 0:3] if xyz_first else gout[:, D:D + 3]              # [B,3,K,S]
            if ctx.needs_input_grad[1]:
                gnew = -gx.sum(dim=2).transpose(1, 2).contiguous()            # [B,S,3]
            if ctx.needs_input_grad[0]:
                rows = gx.permute(0, 3, 2, 1).reshape(B, S * K, 3)
                flat = torch.where(idx < 0, idx + N, idx).reshape(B, S * K, 1).expand(-1, -1, 3)
                gxyz = torch.zeros(B, N, 3, dtype=torch.float32, device=gout.device).scatter_add_(1, flat, rows)
        return gxyz, gnew, gfeat, None, None


def group_points(xyz, new_xyz, feat, idx, xyz_first=True):
    """Utils/Layers.py:57-66 (xyz_first) / Utils/Pointnet2Utils.py:243-253 (features first): xyz [B,N,3],
    new_xyz [B,S,3], feat [B,N,D] or None, idx [B,S,K] -> the Conv2d input [B,3+D,K,S] in one kernel."""
    xyz = _prep(xyz, "xyz"); new_xyz = _prep(new_xyz, "new_xyz")
    if feat is not None:
        feat = _prep(feat, "feat")
        if feat.shape[2] == 0:
            feat = None
    idx = _idx64(idx, xyz.device)
    if idx.dim() != 3 or idx.shape[0] != xyz.shape[0] or idx.shape[1] != new_xyz.shape[1]:
        raise ValueError("group_points: idx must be [B,S,K] with S = new_xyz.shape[1], got %s" % (tuple(idx.shape),))
    if feat is not None and (feat.shape[0] != xyz.shape[0] or feat.shape[1] != xyz.shape[1]):
        raise ValueError("group_points: feat must be [B,N,D] with the same B, N as xyz")
    needs = torch.is_grad_enabled() and (xyz.requires_grad or new_xyz.requires_grad or (feat is not None and feat.requires_grad))
    if needs:
        return _GroupFn.apply(xyz, new_xyz, feat, idx, xyz_first)
    return _group_raw(xyz, new_xyz, feat, idx, xyz_first)


def three_nn(unknown, known, variant=0, want_weight=True):
    """three nearest `known` points for every `unknown` point.
    -> dist [B,N,3] (ascending raw expanded-form values), idx [B,N,3] int64, weight [B,N,3]."""
    unknown = _prep(unknown, "unknown"); known = _prep(known, "known")
    B, N, _ = unknown.shape; S = known.shape[1]
    dev = unknown.device
    dist = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    idx = torch.empty(B, N, 3, dtype=torch.int64, device=dev)
    weight = torch.empty(B, N, 3, dtype=torch.float32, device=dev) if want_weight else None
    lib = _lib.load()
    nws = lib.b200pc_search_workspace_bytes(B, S, N, 3)
    ws = _workspace(nws, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200pc_three_nn(_ptr(unknown), _ptr(known), B, N, S, int(variant), _ptr(dist), _ptr(idx),
                                       _ptr(weight), _ptr(ws), nws, _stream(dev)))
    _bump()
    return dist, idx, weight


def _interp_raw(feat, idx, weight):
    B, S, Cc = feat.shape; N = idx.shape[1]
    out = torch.empty(B, N, Cc, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(_lib.load().b200pc_three_interpolate(_ptr(feat), _ptr(idx), _ptr(weight), B, S, N, Cc, _ptr(out),
                                                        _stream(feat.device)))
    _bump()
    return out


class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, idx, weight):
        ctx.save_for_backward(feat, idx, weight)
        return _interp_raw(feat, idx, weight)

    @staticmethod
    def backward(ctx, gout):
        feat, idx, weight = ctx.saved_tensors
        B, S, Cc = feat.shape; N = idx.shape[1]
        gout = gout.contiguous()
        gfeat = torch.zeros_like(feat)
        gw = torch.empty_like(weight) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(gout.device):
            _lib.check(_lib.load().b200pc_three_interpolate_bwd(_ptr(gout), _ptr(feat), _ptr(idx), _ptr(weight), B, S, N,
                                                                Cc, _ptr(gfeat), _ptr(gw), _stream(gout.device)))
        _bump()
        return gfeat, None, gw


def three_interpolate(feat, idx, weight):
    """feat [B,S,C], idx [B,N,3], weight [B,N,3] -> [B,N,C] = (f0*w0 + f1*w1) + f2*w2."""
    feat = _prep(feat, "feat"); weight = _prep(weight, "weight")
    idx = _idx64(idx, feat.device)
    if (feat.requires_grad or weight.requires_grad) and torch.is_grad_enabled():
        return _InterpFn.apply(feat, idx, weight)
    return _interp_raw(feat, idx, weight)


# ------------------------------------------------------------------------------------------
class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        B, N, _ = x.shape; M = y.shape[1]
        dev = x.device
        dx = torch.empty(B, N, dtype=torch.float32, device=dev); ix = torch.empty(B, N, dtype=torch.int64, device=dev)
        dy = torch.empty(B, M, dtype=torch.float32, device=dev); iy = torch.empty(B, M, dtype=torch.int64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        lib = _lib.load()
        nws = max(lib.b200pc_search_workspace_bytes(B, M, N, 1), lib.b200pc_search_workspace_bytes(B, N, M, 1))
        ws = _workspace(nws, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.b200pc_chamfer_fwd(_ptr(x), _ptr(y), B, N, M, _ptr(dx), _ptr(ix), _ptr(dy), _ptr(iy),
                                              _ptr(loss), _ptr(ws), nws, _stream(dev)))
        _bump()
        ctx.save_for_backward(x, y, ix, iy)
        ctx.mark_non_differentiable(dx, ix, dy, iy)
        return loss.view(()), dx, ix, dy, iy

    @staticmethod
    def backward(ctx, gloss, *_):
        x, y, ix, iy = ctx.saved_tensors
        B, N, _ = x.shape; M = y.shape[1]
        gl = gloss.contiguous().float().view(1)
        gx = torch.empty_like(x); gy = torch.empty_like(y)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().b200pc_chamfer_bwd(_ptr(x), _ptr(y), _ptr(ix), _ptr(iy), _ptr(gl), B, N, M, _ptr(gx),
                                                      _ptr(gy), _stream(x.device)))
        _bump()
        return gx, gy


def chamfer(x, y):
    """x [B,N,3], y [B,M,3] -> (loss scalar, dx [B,N], ix [B,N], dy [B,M], iy [B,M]); loss is differentiable."""
    x = _prep(x, "x"); y = _prep(y, "y")
    return _ChamferFn.apply(x, y)


def fma_peak(iters=4096):
    """sustained packed-FP32 FMA rate of the current device in TFLOP/s (roofline denominator)."""
    tf = C.c_double(0.0); ms = C.c_double(0.0)
    _lib.check(_lib.load().b200pc_fma_peak(int(iters), C.byref(tf), C.byref(ms), _stream(torch.cuda.current_device())))
    return tf.value, ms.value
