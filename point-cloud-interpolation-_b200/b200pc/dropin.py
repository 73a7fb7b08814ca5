"""Make the UNMODIFIED reference code run on libb200pc.so.

The reference has no plugin API: its layers bind the primitives by name at import time
(`from Utils.Pointnet2Utils import farthest_point_sample, index_points, square_distance,
query_ball_point ...`, Utils/Layers.py:8-9; `from pytorch3d.ops import knn_points, knn_gather`,
Utils/Layers.py:10; `from pytorch3d.loss import chamfer_distance`, Utils/Utils.py:9).  `install()`
therefore works on sys.modules:

  1. registers `pytorch3d`, `pytorch3d.ops`, `pytorch3d.loss` modules backed by pytorch3d_shim
     (only if the real pytorch3d is not importable, unless force=True);
  2. imports the reference's primitive module(s) and rebinds the four primitives in them and in
     every already-imported module that had bound the originals by name;
  3. optionally (patch_layers=True) swaps the two layer methods whose hot path is INLINE code in
     the reference -- `Group.forward` (kNN = square_distance + topk, Utils/Layers.py:50-53) and
     `FeaturePropagation.forward` (three-NN = square_distance + full sort, Utils/Layers.py:180-188)
     and `PointNetFeaturePropagation.forward` (Utils/Pointnet2Utils.py:297-304) -- for versions that
     call knn_point / three_nn / three_interpolate, so no [B,N,M] matrix is ever materialised.

Call it once, before or after importing the reference's modules.
"""
import importlib
import sys
import types

import torch
import torch.nn.functional as F

from . import pointnet2_utils as P
from . import pytorch3d_shim as shim

_PRIMS = ("square_distance", "index_points", "farthest_point_sample", "query_ball_point")


def _ensure_stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_pytorch3d(force=False):
    if not force:
        try:
            importlib.import_module("pytorch3d.ops")
            return False
        except Exception:
            pass
    pkg = _ensure_stub("pytorch3d")
    if not hasattr(pkg, "__path__"):
        pkg.__path__ = []          # mark as a package so `import pytorch3d.ops` resolves through sys.modules
    pkg.ops = _ensure_stub("pytorch3d.ops", knn_points=shim.knn_points, knn_gather=shim.knn_gather)
    pkg.loss = _ensure_stub("pytorch3d.loss", chamfer_distance=shim.chamfer_distance)
    return True


# ---- replacement layer methods (same tensor contracts as the reference's) ---------------------
def _group_forward(self, points, new_points, features):
    """Group.forward, Utils/Layers.py:42-66: [B,3,N],[B,3,S],[B,D,N] -> [B,3+D,nsample,S]."""
    pts = points.permute(0, 2, 1).contiguous()
    qry = new_points.permute(0, 2, 1).contiguous()
    feat = features.permute(0, 2, 1).contiguous()
    B, S, Cx = qry.shape
    if self.knn:
        ind = P.knn_point(self.num_samples, pts, qry)
    else:
        ind = P.query_ball_point(self.radius, self.num_samples, pts, qry)
    if Cx == 3:                                      # the fused kernel: gather, centre, cat and layout change at once
        return P.group_points(pts, qry, feat, ind)
    rel = P.index_points(pts, ind) - qry.view(B, S, 1, Cx)
    out = torch.cat([rel, P.index_points(feat, ind)], dim=-1)
    return out.permute(0, 3, 2, 1).contiguous()


def _fp_forward(self, points1, points2, features1, features2):
    """FeaturePropagation.forward, Utils/Layers.py:174-192: [B,3,S],[B,3,N],[B,D1,S],[B,D2,N]."""
    sparse = points1.permute(0, 2, 1).contiguous()
    dense = points2.permute(0, 2, 1).contiguous()
    feat = features1.permute(0, 2, 1).contiguous()
    _, ind, w = P.three_nn_weights(dense, sparse, variant=0)
    new = P.three_interpolate(feat, ind, w).permute(0, 2, 1).contiguous()
    new = torch.cat([new, features2], dim=1)
    return self.conv(new.unsqueeze(3)).squeeze(3)


def _pnfp_forward(self, xyz1, xyz2, points1, points2):
    """PointNetFeaturePropagation.forward, Utils/Pointnet2Utils.py:279-313."""
    dense = xyz1.permute(0, 2, 1)
    sparse = xyz2.permute(0, 2, 1)
    feat = points2.permute(0, 2, 1)
    B, N, _ = dense.shape
    S = sparse.shape[1]
    if S == 1:
        interp = feat.repeat(1, N, 1)
    else:
        _, ind, w = P.three_nn_weights(dense, sparse, variant=1)
        interp = P.three_interpolate(feat, ind, w)
    if points1 is not None:
        new = torch.cat([points1.permute(0, 2, 1), interp], dim=-1)
    else:
        new = interp
    new = new.permute(0, 2, 1)
    norms = getattr(self, "mlp_gns", None) or getattr(self, "mlp_bns", None)
    for i, conv in enumerate(self.mlp_convs):
        new = F.relu(norms[i](conv(new)))
    return new


def _patch_module(mod, originals):
    """rebind primitives in `mod` when it holds the reference's originals (or is their home)."""
    for name in _PRIMS:
        cur = getattr(mod, name, None)
        if cur is not None and (cur is originals.get(name) or getattr(cur, "__module__", "") == mod.__name__):
            setattr(mod, name, getattr(P, name))


def install(reference_modules=("Utils.Pointnet2Utils", "models.pointnet2_utils"), patch_layers=True,
            force_pytorch3d=False):
    """Activate the drop-in.  `reference_modules`: names of the reference's primitive modules to
    rebind if importable (the fork's, PolyPCI's and the upstream copy share these names)."""
    install_pytorch3d(force=force_pytorch3d)
    if "lib2to3.pgen2.token" not in sys.modules:
        try:
            importlib.import_module("lib2to3.pgen2.token")
        except Exception:       # Utils/Pointnet2Utils.py:1 imports an unused symbol from it
            _ensure_stub("lib2to3"); _ensure_stub("lib2to3.pgen2"); _ensure_stub("lib2to3.pgen2.token", NAME=1)
    patched = []
    for modname in reference_modules:
        try:
            mod = importlib.import_module(modname)
        except Exception:
            continue
        originals = {n: getattr(mod, n, None) for n in _PRIMS}
        originals = {n: f for n, f in originals.items() if f is not None and f is not getattr(P, n)}
        for name in originals:
            setattr(mod, name, getattr(P, name))
        for other in list(sys.modules.values()):
            if other is None or other is mod or not hasattr(other, "__dict__"):
                continue
            for name, fn in originals.items():
                if other.__dict__.get(name) is fn:
                    setattr(other, name, getattr(P, name))
        if patch_layers and hasattr(mod, "PointNetFeaturePropagation"):
            mod.PointNetFeaturePropagation.forward = _pnfp_forward
        patched.append(modname)
    if patch_layers:
        for lname in ("Utils.Layers", "models.layers"):
            lay = sys.modules.get(lname)
            if lay is None:
                try:
                    lay = importlib.import_module(lname)
                except Exception:
                    continue
            if hasattr(lay, "Group"):
                lay.Group.forward = _group_forward
            if hasattr(lay, "FeaturePropagation"):
                lay.FeaturePropagation.forward = _fp_forward
            patched.append(lname)
    return patched
