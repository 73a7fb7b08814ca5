"""The reference's algorithm for the hot path, re-expressed with the ATen ops it runs on CPU.

TEST INFRASTRUCTURE ONLY.  This is the "port" that bench.py's cpu_baseline / --impl reference
legs time on the GPU box (where /root/reference does not exist): dense [B,N,M] fp32 distance
matrix from a K=3 matmul, then topk / full sort, and the python-loop FPS -- i.e. the same work
the reference asks of the host cores.  It is validated against the real reference in the build
container by tests/test_oracle.py (bit-identical outputs on tie-free rows) and against
oracle/strict.c.  It is never used to check the CUDA path (strict.c is), and never shipped.

Reference lines followed (relative to the reference root):
  dense_sqdist      Utils/Pointnet2Utils.py:20-41
  gather_rows       Utils/Pointnet2Utils.py:44-61
  fps               Utils/Pointnet2Utils.py:64-85
  ball              Utils/Pointnet2Utils.py:88-108
  knn_topk          Utils/Layers.py:50-53
  three_nn_interp   Utils/Layers.py:180-188 (variant 0), Utils/Pointnet2Utils.py:297-304 (variant 1)
  knn_points_dense  pytorch3d.ops.knn_points semantics (not vendored -> PARITY UNPINNED)
  chamfer_dense     Utils/Utils.py:39-48 -> pytorch3d chamfer_distance defaults (PARITY UNPINNED)
"""
import torch


def dense_sqdist(src, dst):
    """[B,N,3],[B,M,3] -> [B,N,M]:  -2 * src @ dst^T, then += |src|^2, then += |dst|^2."""
    out = torch.matmul(src, dst.transpose(1, 2)) * -2
    out += (src ** 2).sum(-1).unsqueeze(2)
    out += (dst ** 2).sum(-1).unsqueeze(1)
    return out


def gather_rows(points, idx):
    """points [B,N,C], idx [B,...] -> [B,...,C] by advanced indexing with a broadcast batch index."""
    B = points.shape[0]
    bshape = [B] + [1] * (idx.dim() - 1)
    batch = torch.arange(B, dtype=torch.long, device=points.device).view(bshape).expand_as(idx)
    return points[batch, idx, :]


def fps(xyz, npoint, start=None):
    """Iterative farthest point sampling.  start: optional [B] long (the reference draws
    torch.randint(0, N, (B,)) from the CPU generator; pass None to do the same)."""
    B, N, _ = xyz.shape
    picked = torch.zeros(B, npoint, dtype=torch.long)
    mind = torch.full((B, N), 1e10, dtype=xyz.dtype)
    far = torch.randint(0, N, (B,), dtype=torch.long) if start is None else start.clone()
    rows = torch.arange(B, dtype=torch.long)
    for i in range(npoint):
        picked[:, i] = far
        centre = xyz[rows, far, :].unsqueeze(1)
        d = ((xyz - centre) ** 2).sum(-1)
        closer = d < mind
        mind[closer] = d[closer]
        far = mind.max(-1)[1]
    return picked


def ball(radius, nsample, xyz, new_xyz):
    """Ball query by masking an index matrix and fully sorting each row."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    grid = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat(B, S, 1)
    d = dense_sqdist(new_xyz, xyz)
    grid[d > radius ** 2] = N
    grid = grid.sort(dim=-1)[0][:, :, :nsample]
    first = grid[:, :, :1].expand(-1, -1, nsample)
    empty = grid == N
    grid[empty] = first[empty]
    return grid


def knn_topk(nsample, xyz, new_xyz):
    """kNN as Group.forward does it: distances with the refs as `src`, topk along dim 1."""
    d = dense_sqdist(xyz, new_xyz)                       # [B,N,S]
    return d.topk(nsample, dim=1, largest=False)[1].transpose(1, 2).contiguous()


def three_nn_interp(dense_xyz, sparse_xyz, sparse_feat, variant=0):
    """three nearest sparse points by a full sort, inverse-distance weights, weighted sum.
    returns (out [B,N,C], dist [B,N,3], idx [B,N,3], weight [B,N,3])."""
    B, N, _ = dense_xyz.shape
    d, idx = dense_sqdist(dense_xyz, sparse_xyz).sort(dim=-1)
    d, idx = d[:, :, :3], idx[:, :, :3]
    if variant == 0:
        d = d.clone()
        d[d < 1e-10] = 1e-10
        inv = 1.0 / d
    else:
        inv = 1.0 / (d + 1e-8)
    w = inv / inv.sum(dim=2, keepdim=True)
    out = (gather_rows(sparse_feat, idx) * w.view(B, N, 3, 1)).sum(dim=2)
    return out, d, idx, w


def knn_points_dense(p1, p2, K):
    """pytorch3d.knn_points stand-in on CPU: direct-form squared distances, K smallest ascending.
    (Unpinned third-party semantics; used for timing the fusion-kNN / Chamfer rows only.)"""
    diff = p1.unsqueeze(2) - p2.unsqueeze(1)              # [B,P1,P2,3]
    d = (diff * diff).sum(-1)
    dk, ik = d.topk(K, dim=2, largest=False)
    return dk, ik


def chamfer_dense(x, y):
    """chamfer_distance(x, y) defaults: squared L2, mean over points, mean over batch."""
    dx, _ = knn_points_dense(x, y, 1)
    dy, _ = knn_points_dense(y, x, 1)
    return (dx[..., 0].mean(1) + dy[..., 0].mean(1)).mean()
