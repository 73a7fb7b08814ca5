"""pytorch3d-shaped entry points backed by libb200pc.so.

The reference imports `knn_points`, `knn_gather` (Utils/Layers.py:10, PolyPCI/Models/Models_V1.py:12,
PointINet20230424/models/layers.py:16) and `chamfer_distance` (Utils/Utils.py:9) from pytorch3d,
which it neither vendors nor pins.  These functions reproduce the call signatures and the
documented semantics the reference relies on: squared L2 in direct form, K results ascending,
ties to the lower index, int64 indices; chamfer with point_reduction = batch_reduction = "mean".
PARITY UNPINNED: no golden vector for this arithmetic exists in the reference.
"""
from collections import namedtuple

import torch

from . import ops

_KNN = namedtuple("KNN", "dists idx knn")


def knn_gather(x, idx, lengths=None):
    """x [B,M,C], idx [B,N,K] -> [B,N,K,C]."""
    if lengths is not None:
        raise NotImplementedError("b200pc.knn_gather: ragged `lengths` are not used by the reference")
    return ops.gather(x, idx)


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    """for every point of p1 [B,P1,3] its K nearest points of p2 [B,P2,3].
    -> KNN(dists [B,P1,K] squared, idx [B,P1,K] int64, knn [B,P1,K,3] or None)."""
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("b200pc.knn_points: ragged `lengths` are not used by the reference")
    if norm != 2:
        raise NotImplementedError("b200pc.knn_points: only norm=2 (the reference's only use)")
    if p1.shape[-1] != 3 or p2.shape[-1] != 3:
        raise ValueError("b200pc.knn_points: points must be [B,P,3]")
    K = min(int(K), p2.shape[1])
    if K <= 0 or p1.shape[1] == 0:
        B, P1 = p1.shape[0], p1.shape[1]
        z = p1.new_zeros(B, P1, 0)
        return _KNN(z, z.long(), p1.new_zeros(B, P1, 0, 3) if return_nn else None)
    idx, dists = ops.knn_search(p2.detach(), p1.detach(), K, ops.FORM_DIRECT, want_dist=True)
    nn = None
    grad = torch.is_grad_enabled() and (p1.requires_grad or p2.requires_grad)
    if return_nn or grad:
        nn = ops.gather(p2, idx)
    if grad:  # differentiable distances recomputed from the gathered neighbours
        diff = p1.unsqueeze(2) - nn
        dists = (diff * diff).sum(-1)
    return _KNN(dists, idx, nn if return_nn else None)


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean", norm=2, **_ignored):
    """x [B,N,3], y [B,M,3] -> (loss, None), the only form the reference uses (Utils/Utils.py:47)."""
    if any(v is not None for v in (x_lengths, y_lengths, x_normals, y_normals, weights)):
        raise NotImplementedError("b200pc.chamfer_distance: lengths / normals / weights are not used by the reference")
    if batch_reduction != "mean" or point_reduction != "mean" or norm != 2:
        raise NotImplementedError("b200pc.chamfer_distance: only the pytorch3d defaults are implemented")
    loss = ops.chamfer(x, y)[0]
    return loss, None
