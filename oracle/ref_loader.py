"""Import the REAL reference: from /root/reference in the build container, from its byte-for-byte staged
copy oracle/_ref/ (oracle/make_ref.py; git-ignored, travels with gpurun) on the GPU box.  TEST INFRASTRUCTURE.

The reference does not import on this image as-is:
  * Utils/Pointnet2Utils.py:1 imports lib2to3 (absent in Python 3.12) for an unused symbol;
  * Utils/Layers.py:10 / Utils/Utils.py:9-10 import pytorch3d and emd (not installed).
We register empty in-memory modules for those names, then import the reference's own files
UNMODIFIED.  Nothing is copied into the repo's history.  Where neither location exists `available()` is
False and every caller must skip.

`pytorch3d_provider` lets a caller decide what backs knn_points / knn_gather /
chamfer_distance when importing Utils.Layers (e.g. the b200pc shim for the drop-in test, or
the torch stand-in of oracle/ref_torch.py for CPU runs).
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")


def _find_root():
    env = os.environ.get("B200PC_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", STAGED_ROOT]:
        if cand and os.path.isfile(os.path.join(cand, "Utils", "Pointnet2Utils.py")):
            return cand
    return env or "/root/reference"


REF_ROOT = _find_root()


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
    pkg = _stub("pytorch3d")
    pkg.ops = _stub("pytorch3d.ops", knn_points=pytorch3d_provider.knn_points, knn_gather=pytorch3d_provider.knn_gather)
    pkg.loss = _stub("pytorch3d.loss", chamfer_distance=pytorch3d_provider.chamfer_distance)


def _import_from(root, modname):
    if root not in sys.path:
        sys.path.insert(0, root)
    blocking = os.environ.get("CUDA_LAUNCH_BLOCKING")
    try:
        return importlib.import_module(modname)
    finally:
        # Models/*.py:11 set CUDA_LAUNCH_BLOCKING=1 at import; a test / bench process must not inherit that
        if blocking is None:
            os.environ.pop("CUDA_LAUNCH_BLOCKING", None)
        else:
            os.environ["CUDA_LAUNCH_BLOCKING"] = blocking


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


def fork_models(pytorch3d_provider=None, name="Models.New_Models0"):
    """the fork's model zoo (Models/New_Models0.py: FlowNet3D + ISAPCInet, the one train.py:13 uses).  Imports
    Dataset.InterpolationData and Utils.Visualize, hence the open3d stub."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs(pytorch3d_provider)
    return _import_from(REF_ROOT, name)


def utils_losses(pytorch3d_provider=None):
    """Utils/Utils.py (chamfer_loss, :39-48)."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs(pytorch3d_provider)
    return _import_from(REF_ROOT, "Utils.Utils")


def demo_bins():
    """paths of the shipped sweeps: ([KITTI f32x4 ...], [nuScenes f32x5 ...])"""
    import glob
    kitti = sorted(glob.glob(os.path.join(REF_ROOT, "PointINet20230424/data/demo_data/original/*.bin")))
    nusc = sorted(glob.glob(os.path.join(REF_ROOT, "Demos/20230508test/demo_data/Inputs/*.bin")))
    return kitti, nusc


def strict_provider():
    """pytorch3d stand-in for the CPU arm of the drop-in tests: knn_points from oracle/strict.c (threaded C, direct
    form, (distance, index) order -- the semantics this repo pins for the unvendored pytorch3d), knn_gather by
    advanced indexing, chamfer_distance from strict.c.  Forward only (no autograd through the distances)."""
    import numpy as np
    import torch
    from . import ref_torch, strict

    def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
        K = min(int(K), p2.shape[1])
        if K <= 0:
            z = p1.new_zeros(p1.shape[0], p1.shape[1], 0)
            return z, z.long(), (p1.new_zeros(p1.shape[0], p1.shape[1], 0, 3) if return_nn else None)
        d, i = strict.knn_points(p1.detach().cpu().numpy(), p2.detach().cpu().numpy(), K)
        d = torch.from_numpy(d); i = torch.from_numpy(i)
        return d, i, (ref_torch.gather_rows(p2, i) if return_nn else None)

    def knn_gather(x, idx, lengths=None):
        return ref_torch.gather_rows(x, idx)

    def chamfer_distance(x, y, **kw):
        return torch.tensor(strict.chamfer(x.detach().cpu().numpy(), y.detach().cpu().numpy())[0], dtype=torch.float32), None

    return types.SimpleNamespace(knn_points=knn_points, knn_gather=knn_gather, chamfer_distance=chamfer_distance)


def rebind_pytorch3d(provider):
    """point the already-imported reference modules at another pytorch3d provider (they bind the functions by name)"""
    install_stubs(provider)
    for mod in list(sys.modules.values()):
        if mod is None or not hasattr(mod, "__dict__") or getattr(mod, "__name__", "").startswith("pytorch3d"):
            continue
        file = getattr(mod, "__file__", None) or ""
        if not file.startswith(REF_ROOT):
            continue
        for n in ("knn_points", "knn_gather", "chamfer_distance"):
            if n in mod.__dict__:
                setattr(mod, n, getattr(provider, n))


def polypci_models(pytorch3d_provider=None):
    """PolyPCI/Models/Models_V1.py (FlowNet3D + PolyPCI: rebuild :102-114, fitting_and_predict :116-124, forward :126-222).
    The file imports `Dataset.Dataset.NuscenesDataset` (PolyPCI's own loader, open3d-based) and `Utils.*` (byte-identical to
    the top-level copies): PolyPCI/ goes on sys.path behind the repository root, and a placeholder stands in for the
    dataset module when the top-level `Dataset` package (which has no Dataset.py) was imported first."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    install_stubs(pytorch3d_provider)
    if "Dataset.Dataset" not in sys.modules:
        try:
            if REF_ROOT not in sys.path:
                sys.path.insert(0, REF_ROOT)
            importlib.import_module("Dataset.Dataset")
        except Exception:
            _stub("Dataset"); _stub("Dataset.Dataset", NuscenesDataset=object)
    poly_root = os.path.join(REF_ROOT, "PolyPCI")
    if poly_root not in sys.path:
        sys.path.append(poly_root)
    return _import_from(REF_ROOT, "Models.Models_V1")
