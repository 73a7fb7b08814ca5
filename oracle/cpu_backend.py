"""A CPU 'backend' for b200pc.pointinet built from the oracle's torch port -- TEST ONLY.
It lets the tests check the model glue (layouts, RNG draws, parameter names) against the real
upstream model without a GPU, and gives the GPU tests a same-weights CPU result to compare with."""
import types
from collections import namedtuple

import torch

from . import ref_torch

_KNN = namedtuple("KNN", "dists idx knn")


def _three_nn_weights(dense, sparse, variant=0):
    d, idx = ref_torch.dense_sqdist(dense, sparse).sort(dim=-1)
    d, idx = d[:, :, :3].clone(), idx[:, :, :3]
    if variant == 0:
        d[d < 1e-10] = 1e-10
        inv = 1.0 / d
    else:
        inv = 1.0 / (d + 1e-8)
    return d, idx, inv / inv.sum(dim=2, keepdim=True)


def _three_interpolate(feat, idx, w):
    B, N, _ = idx.shape
    return (ref_torch.gather_rows(feat, idx) * w.view(B, N, 3, 1)).sum(dim=2)


def _knn_points(p1, p2, K=1, return_nn=False, **_):
    if K <= 0:
        z = p1.new_zeros(p1.shape[0], p1.shape[1], 0)
        return _KNN(z, z.long(), p1.new_zeros(p1.shape[0], p1.shape[1], 0, 3))
    d, i = ref_torch.knn_points_dense(p1, p2, K)
    return _KNN(d, i, ref_torch.gather_rows(p2, i) if return_nn else None)


def make():
    return types.SimpleNamespace(
        farthest_point_sample=lambda xyz, n: ref_torch.fps(xyz, n),
        index_points=ref_torch.gather_rows,
        query_ball_point=ref_torch.ball,
        knn_point=ref_torch.knn_topk,
        three_nn_weights=_three_nn_weights,
        three_interpolate=_three_interpolate,
        knn_points=_knn_points,
        knn_gather=lambda x, idx, lengths=None: ref_torch.gather_rows(x, idx),
        randperm=lambda n, keep, device: torch.randperm(n)[:keep].to(device))
