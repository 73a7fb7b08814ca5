"""The reference's geometric primitives under their own names, signatures and tensor layouts.

Mirrors the module-level functions of the reference's Utils/Pointnet2Utils.py (== PolyPCI/Utils/
Pointnet2Utils.py, ~= PointINet20230424/models/pointnet2_utils.py) so that Utils/Layers.py,
Models/*.py, train*.py and test.py can import them unmodified (see dropin.py), plus the three
names north_star asks for that exist in the reference only as inline code:

  knn_point(nsample, xyz, new_xyz)        <- Group.forward kNN branch, Utils/Layers.py:50-53
  three_nn(unknown, known)                <- Utils/Layers.py:180-182, Utils/Pointnet2Utils.py:297-299
  three_interpolate(feats, idx, weight)   <- Utils/Layers.py:187-188, Utils/Pointnet2Utils.py:304

Every function runs a hand-written sm_100a kernel from libb200pc.so; none has a CPU path.
"""
import torch

from . import ops


def square_distance(src, dst):
    """Utils/Pointnet2Utils.py:20-41.  src [B,N,C=3], dst [B,M,3] -> [B,N,M].
    Bit-identical to the reference's torch-CPU result: ((-2*src.dst) + |src|^2) + |dst|^2."""
    return ops.square_distance(src, dst)


def index_points(points, idx):
    """Utils/Pointnet2Utils.py:44-61.  points [B,N,C], idx [B,S] or [B,S,K] (any integer dtype)
    -> [B,S,C] / [B,S,K,C].  Negative indices wrap; out-of-range raises IndexError when
    b200pc.ops.CHECK_BOUNDS is on (off by default: it costs a host sync)."""
    return ops.gather(points, idx)


def farthest_point_sample(xyz, npoint):
    """Utils/Pointnet2Utils.py:64-85.  xyz [B,N,3] -> [B,npoint] int64.
    Like the reference (:76) the first centroid is drawn with torch.randint(0, N, (B,)) from
    the CPU default generator, so torch.manual_seed() reproduces the reference's samples."""
    B, N, _ = xyz.shape
    start = torch.randint(0, N, (B,), dtype=torch.long)
    return ops.fps(xyz, npoint, start.to(xyz.device, non_blocking=True))


def sample_points(xyz, npoint, start=None):
    """Sample.forward (Utils/Layers.py:23-27) on point-major input: farthest_point_sample + index_points(points, ind) in
    one C call.  xyz [B,N,3] -> (idx [B,npoint] int64, new_xyz [B,npoint,3]).  The start index is drawn exactly like
    farthest_point_sample does unless `start` [B] is given."""
    if start is None:
        start = torch.randint(0, xyz.shape[1], (xyz.shape[0],), dtype=torch.long).to(xyz.device, non_blocking=True)
    return ops.fps(xyz, npoint, start, want_xyz=True)


def farthest_point_sample_from(xyz, npoint, start):
    """same, with caller-provided first centroids `start` [B] (no RNG draw)."""
    return ops.fps(xyz, npoint, start)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """Utils/Pointnet2Utils.py:88-108.  xyz [B,N,3] refs, new_xyz [B,S,3] queries ->
    [B,S,nsample] int64: lowest-index neighbours within the radius, padded with the first one;
    a query whose ball is empty gets N in every slot (the reference's sentinel, never clamped)."""
    return ops.ball_query(radius, nsample, xyz, new_xyz)


def knn_point(nsample, xyz, new_xyz):
    """kNN as Group.forward computes it (Utils/Layers.py:50-53): xyz [B,N,3] refs, new_xyz
    [B,S,3] queries -> [B,S,nsample] int64 ascending by (distance, index)."""
    return ops.knn_search(xyz, new_xyz, nsample, ops.FORM_REF_NORM_FIRST)


def three_nn(unknown, known):
    """unknown [B,N,3], known [B,S,3] -> (dist [B,N,3], idx [B,N,3] int64), the three nearest
    known points in ascending order; dist are the reference's expanded-form squared distances."""
    dist, idx, _ = ops.three_nn(unknown, known, variant=0, want_weight=False)
    return dist, idx


def three_nn_weights(unknown, known, variant=0):
    """three_nn plus the inverse-distance weights.  variant 0 = FeaturePropagation
    (Utils/Layers.py:183-186), variant 1 = PointNetFeaturePropagation (Utils/Pointnet2Utils.py:301-303).
    When a coordinate tensor requires grad the distances and weights carry gradient to it, as the reference's
    `1.0 / dists` does (ISAPCInet trains through them, Models/New_Models0.py:164-172)."""
    if torch.is_grad_enabled() and (unknown.requires_grad or known.requires_grad):
        return ops.three_nn_autograd(unknown, known, variant=variant)
    return ops.three_nn(unknown, known, variant=variant, want_weight=True)


def feature_propagation(unknown, known, feats, variant=0):
    """The interpolation of FeaturePropagation.forward (Utils/Layers.py:180-188, variant 0) /
    PointNetFeaturePropagation.forward (Utils/Pointnet2Utils.py:297-304, variant 1) in one call:
    unknown [B,N,3], known [B,S,3], feats [B,S,C] -> [B,N,C]; differentiable in feats, and in the coordinates
    when they require grad."""
    return ops.feature_propagation(unknown, known, feats, variant)


def fusion_group(points1, points2, k, features2=None):
    """PointsFusion.knn_group (Utils/Layers.py:207-226, PointINet20230424/models/layers.py:346-368) and
    knn_group_withI (Utils/Layers.py:384-402) on point-major inputs: points1 [B,S,3] queries, points2 [B,N,3]
    refs, features2 [B,N,Cf] or None -> (new_features [B,4,S,k], nn [B,3,S,k], grouped features [B,Cf,S,k],
    idx [B,S,k]) from one C call (direct-form search + one feature kernel)."""
    return ops.fusion_group(points1, points2, k, features2)


def three_interpolate(feats, idx, weight):
    """feats [B,S,C], idx [B,N,3], weight [B,N,3] -> [B,N,C]; differentiable in feats and weight."""
    return ops.three_interpolate(feats, idx, weight)


def group_points(xyz, new_xyz, features, idx, xyz_first=True):
    """Fused tail of Group.forward (Utils/Layers.py:57-66) / the SA-MSG grouping (Utils/Pointnet2Utils.py:243-253):
    index_points(xyz, idx) - new_xyz.view(B,S,1,3), index_points(features, idx), cat, permute(0,3,2,1).contiguous()
    -> [B,3+D,K,S] in one kernel.  Point-major inputs (xyz [B,N,3], new_xyz [B,S,3], features [B,N,D] or None)."""
    return ops.group_points(xyz, new_xyz, features, idx, xyz_first)
