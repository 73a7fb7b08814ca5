"""Import the REAL reference from /root/reference (build container only).  TEST INFRASTRUCTURE.

The reference does not import on this image as-is:
  * Utils/Pointnet2Utils.py:1 imports lib2to3 (absent in Python 3.12) for an unused symbol;
  * Utils/Layers.py:10 / Utils/Utils.py:9-10 import pytorch3d and emd (not installed).
We register empty in-memory modules for those names, then import the reference's own files
UNMODIFIED.  Nothing is copied into this repo.  /root/reference does not exist on the GPU box:
`available()` is False there and every caller must skip.

`pytorch3d_provider` lets a caller decide what backs knn_points / knn_gather /
chamfer_distance when importing Utils.Layers (e.g. the b200pc shim for the drop-in test, or
the torch stand-in of oracle/ref_torch.py for CPU runs).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("B200PC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "Utils", "Pointnet2Utils.py"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__dict__["__b200pc_stub__"] = True
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs(pytorch3d_provider=None):
    """lib2to3 / emd / open3d / wandb placeholders + a pytorch3d module backed by the provider."""
    _stub("lib2to3"); _stub("lib2to3.pgen2"); _stub("lib2to3.pgen2.token", NAME=1)
    _stub("emd")
    for n in ("open3d", "wandb"):
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    if pytorch3d_provider is None:
        from . import ref_torch

        def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1,
                       return_nn=False, return_sorted=True):
            d, i = ref_torch.knn_points_dense(p1, p2, K)
            nn = ref_torch.gather_rows(p2, i) if return_nn else None
            return d, i, nn

        def knn_gather(x, idx, lengths=None):
            return ref_torch.gather_rows(x, idx)

        def chamfer_distance(x, y, **kw):
            return ref_torch.chamfer_dense(x, y), None

        pytorch3d_provider = types.SimpleNamespace(knn_points=knn_points, knn_gather=knn_gather,
                                                   chamfer_distance=chamfer_distance)
    _stub("pytorch3d")
    _stub("pytorch3d.ops", knn_points=pytorch3d_provider.knn_points, knn_gather=pytorch3d_provider.knn_gather)
    _stub("pytorch3d.loss", chamfer_distance=pytorch3d_provider.chamfer_distance)


def _import_from(root, modname):
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module(modname)


def pointnet2_utils():
    """the reference's Utils/Pointnet2Utils.py module object."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs()
    return _import_from(REF_ROOT, "Utils.Pointnet2Utils")


def layers(pytorch3d_provider=None):
    """the reference's Utils/Layers.py module object."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs(pytorch3d_provider)
    return _import_from(REF_ROOT, "Utils.Layers")


def upstream_pointinet(pytorch3d_provider=None):
    """PointINet20230424/models/models.py (the coherent upstream PointINet).  Its packages are
    named `models.*`, so PointINet20230424/ itself goes on sys.path."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs(pytorch3d_provider)
    return _import_from(os.path.join(REF_ROOT, "PointINet20230424"), "models.models")
