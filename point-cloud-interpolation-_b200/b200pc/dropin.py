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
  3. optionally (patch_layers=True) swaps the layer methods whose hot path is INLINE code in
     the reference -- `Group.forward` (kNN = square_distance + topk, Utils/Layers.py:50-53),
     `FeaturePropagation.forward` (three-NN = square_distance + full sort, Utils/Layers.py:180-188),
     `PointNetFeaturePropagation.forward` (Utils/Pointnet2Utils.py:297-304),
     `PointNetSetAbstractionMsg.forward` (grouping :240-253), `PointsFusion(.2).knn_group`
     (Utils/Layers.py:207-226, upstream layers.py:346-368) and `knn_group_withI` (:384-402) -- for versions
     that call knn_point / feature_propagation / group_points / fusion_group, so no [B,N,M] matrix is ever
     materialised and each grouping is one C call.  Every replacement keeps the reference's tensor contract
     and falls back to the reference's own differentiable torch composition (over the shim) when a coordinate
     tensor requires grad.

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
_MISSING = object()
_UNDO = []          # (object, attribute, previous value or _MISSING), in the order install() changed them


def _set(obj, name, value):
    _UNDO.append((obj, name, obj.__dict__.get(name, _MISSING) if hasattr(obj, "__dict__") else getattr(obj, name, _MISSING)))
    setattr(obj, name, value)


def uninstall():
    """undo every rebinding install() made (reference modules, classes, pytorch3d stand-ins), newest first"""
    while _UNDO:
        obj, name, old = _UNDO.pop()
        if old is _MISSING:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, old)
    for name in [n for n, m in sys.modules.items() if getattr(m, "__b200pc_dropin_stub__", False)]:
        del sys.modules[name]


def _ensure_stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__b200pc_dropin_stub__ = True
        sys.modules[name] = m
    for k, v in attrs.items():
        _set(m, k, v)
    return m


_P3D_NAMES = ("knn_points", "knn_gather", "chamfer_distance")


def _imports_pytorch3d(mod):
    """does this module's source bind names from pytorch3d (`from pytorch3d.ops import knn_points, knn_gather`)?"""
    path = getattr(mod, "__file__", None)
    if not path or not path.endswith(".py"):
        return False
    try:
        with open(path, "r", errors="replace") as fh:
            return "from pytorch3d" in fh.read()
    except OSError:
        return False


def _is_stub(m):
    return getattr(m, "__b200pc_dropin_stub__", False) or getattr(m, "__b200pc_stub__", False)


def _real_pytorch3d():
    m = sys.modules.get("pytorch3d.ops")
    if m is not None:
        return not _is_stub(m)
    try:
        importlib.import_module("pytorch3d.ops")
        return True
    except Exception:
        return False


def install_pytorch3d(force=False):
    """`pytorch3d.ops.knn_points / knn_gather` and `pytorch3d.loss.chamfer_distance` backed by the shim, unless a real
    pytorch3d is importable (force=True replaces that too).  Modules that already bound the old functions by name
    (`from pytorch3d.ops import knn_points`) are rebound as well."""
    if _real_pytorch3d() and not force:
        return False
    pkg = _ensure_stub("pytorch3d")
    if not hasattr(pkg, "__path__"):
        _set(pkg, "__path__", [])      # mark as a package so `import pytorch3d.ops` resolves through sys.modules
    _set(pkg, "ops", _ensure_stub("pytorch3d.ops", knn_points=shim.knn_points, knn_gather=shim.knn_gather))
    _set(pkg, "loss", _ensure_stub("pytorch3d.loss", chamfer_distance=shim.chamfer_distance))
    for other in list(sys.modules.values()):
        if other is None or not hasattr(other, "__dict__") or getattr(other, "__name__", "").startswith(("pytorch3d", "b200pc")):
            continue
        if not any(n in other.__dict__ for n in _P3D_NAMES) or not _imports_pytorch3d(other):
            continue
        for n in _P3D_NAMES:
            if n in other.__dict__ and other.__dict__[n] is not getattr(shim, n):
                _set(other, n, getattr(shim, n))
    return True


# ---- replacement layer methods (same tensor contracts as the reference's) ---------------------
def _wants_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


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
    """FeaturePropagation.forward, Utils/Layers.py:174-192: [B,3,S],[B,3,N],[B,D1,S],[B,D2,N].  The weights carry
    gradient to the coordinates when those require grad (ops.three_nn_autograd), like the reference's 1.0 / dists."""
    sparse = points1.permute(0, 2, 1).contiguous()
    dense = points2.permute(0, 2, 1).contiguous()
    feat = features1.permute(0, 2, 1).contiguous()
    new = P.feature_propagation(dense, sparse, feat, variant=0).permute(0, 2, 1).contiguous()
    new = torch.cat([new, features2], dim=1)
    return self.conv(new.unsqueeze(3)).squeeze(3)


def _norms(self, *names):
    for n in names:
        v = getattr(self, n, None)
        if v is not None:
            return v
    raise AttributeError("none of %s on %s" % (names, type(self).__name__))


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
        interp = P.feature_propagation(dense, sparse, feat, variant=1)
    if points1 is not None:
        new = torch.cat([points1.permute(0, 2, 1), interp], dim=-1)
    else:
        new = interp
    new = new.permute(0, 2, 1)
    norms = _norms(self, "mlp_gns", "mlp_bns")
    for i, conv in enumerate(self.mlp_convs):
        new = F.relu(norms[i](conv(new)))
    return new


def _samsg_forward(self, xyz, points):
    """PointNetSetAbstractionMsg.forward, Utils/Pointnet2Utils.py:226-264: FPS + gather in one call, then per radius a
    ball query and the fused grouping kernel in its features-first layout (xyz_first=False) -- the reference's
    index_points x2, in-place centring, cat and permute (:243-253) never materialise."""
    pts_xyz = xyz.permute(0, 2, 1)
    feat = points.permute(0, 2, 1) if points is not None else None
    B, N, _ = pts_xyz.shape
    S = self.npoint
    _, new_xyz = P.sample_points(pts_xyz, S)                       # draws torch.randint like farthest_point_sample (:76)
    norms = _norms(self, "gn_blocks", "bn_blocks")
    outs = []
    for i, radius in enumerate(self.radius_list):
        K = self.nsample_list[i]
        group_idx = P.query_ball_point(radius, K, pts_xyz, new_xyz)
        grouped = P.group_points(pts_xyz, new_xyz, feat, group_idx, xyz_first=False)     # [B, D+3, K, S]
        for j, conv in enumerate(self.conv_blocks[i]):
            grouped = F.relu(norms[i][j](conv(grouped)))
        outs.append(torch.max(grouped, 2)[0])
    return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)


def _make_fusion_knn_group(original, with_features):
    """PointsFusion.knn_group: fork signature (points1, points2, k) -> 2 tensors (Utils/Layers.py:207-226); upstream
    signature (points1, points2, features2, k) -> 3 tensors (PointINet20230424/models/layers.py:346-368)."""
    if with_features:
        def knn_group(self, points1, points2, features2, k):
            if k < 1 or _wants_grad(points1, points2, features2):
                return original(self, points1, points2, features2, k)      # torch composition over the shim: differentiable
            resi, nn, gf, _ = P.fusion_group(points1.permute(0, 2, 1), points2.permute(0, 2, 1), k, features2.permute(0, 2, 1))
            return resi, nn, gf
    else:
        def knn_group(self, points1, points2, k):
            if k < 1 or _wants_grad(points1, points2):
                return original(self, points1, points2, k)
            resi, nn, _, _ = P.fusion_group(points1.permute(0, 2, 1), points2.permute(0, 2, 1), k)
            return resi, nn
    knn_group.__doc__ = original.__doc__
    knn_group._b200pc_original = original
    return knn_group


def _make_knn_group_withI(original):
    def knn_group_withI(points1, points2, intensity2, k):
        """knn_group_withI, Utils/Layers.py:384-402."""
        if k < 1 or _wants_grad(points1, points2, intensity2):
            return original(points1, points2, intensity2, k)
        resi, nn, gf, _ = P.fusion_group(points1.permute(0, 2, 1), points2.permute(0, 2, 1), k, intensity2.permute(0, 2, 1))
        return resi, nn, gf
    knn_group_withI._b200pc_original = original
    return knn_group_withI


def _patch_module(mod, originals):
    """rebind primitives in `mod` when it holds the reference's originals (or is their home)."""
    for name in _PRIMS:
        cur = getattr(mod, name, None)
        if cur is not None and (cur is originals.get(name) or getattr(cur, "__module__", "") == mod.__name__):
            _set(mod, name, getattr(P, name))


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
            _set(mod, name, getattr(P, name))
        for other in list(sys.modules.values()):
            if other is None or other is mod or not hasattr(other, "__dict__"):
                continue
            for name, fn in originals.items():
                if other.__dict__.get(name) is fn:
                    _set(other, name, getattr(P, name))
        if patch_layers and hasattr(mod, "PointNetFeaturePropagation"):
            _set(mod.PointNetFeaturePropagation, "forward", _pnfp_forward)
        if patch_layers and hasattr(mod, "PointNetSetAbstractionMsg"):
            _set(mod.PointNetSetAbstractionMsg, "forward", _samsg_forward)
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
                _set(lay.Group, "forward", _group_forward)
            if hasattr(lay, "FeaturePropagation"):
                _set(lay.FeaturePropagation, "forward", _fp_forward)
            for cname in ("PointsFusion", "PointsFusion2"):
                cls = getattr(lay, cname, None)
                if cls is not None and not hasattr(cls.knn_group, "_b200pc_original"):
                    nargs = cls.knn_group.__code__.co_argcount           # self, points1, points2, [features2,] k
                    _set(cls, "knn_group", _make_fusion_knn_group(cls.knn_group, with_features=nargs == 5))
            fn = getattr(lay, "knn_group_withI", None)
            if fn is not None and not hasattr(fn, "_b200pc_original"):
                new_fn = _make_knn_group_withI(fn)
                for other in list(sys.modules.values()):                   # modules that imported it by name
                    if other is not None and hasattr(other, "__dict__") and other.__dict__.get("knn_group_withI") is fn:
                        _set(other, "knn_group_withI", new_fn)
            patched.append(lname)
    return patched


def import_reference(root, module, extra_stubs=("emd", "open3d", "wandb")):
    """Import one of the reference's own modules from a checkout at `root` (e.g. ".../PointINet20230424" + "models.models",
    or the repository root + "Models.New_Models0") with the drop-in active.  The reference does not import on a stock
    Python 3.12 image: Utils/Pointnet2Utils.py:1 needs lib2to3, Utils/Utils.py:10 the `emd` extension, the datasets and
    visualisers open3d / wandb -- all unused by the hot path; empty placeholder modules are registered for those that are
    not installed.  Models/*.py:11 set CUDA_LAUNCH_BLOCKING at import; the caller's value is restored."""
    import os
    for name in extra_stubs:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _ensure_stub(name)
    install_pytorch3d()
    if "lib2to3.pgen2.token" not in sys.modules:
        try:
            importlib.import_module("lib2to3.pgen2.token")
        except Exception:
            _ensure_stub("lib2to3"); _ensure_stub("lib2to3.pgen2"); _ensure_stub("lib2to3.pgen2.token", NAME=1)
    if root not in sys.path:
        sys.path.insert(0, root)
    blocking = os.environ.get("CUDA_LAUNCH_BLOCKING")
    try:
        mod = importlib.import_module(module)
    finally:
        if blocking is None:
            os.environ.pop("CUDA_LAUNCH_BLOCKING", None)
        else:
            os.environ["CUDA_LAUNCH_BLOCKING"] = blocking
    install()
    return mod
