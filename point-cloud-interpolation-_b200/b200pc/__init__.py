"""b200pc -- B200-native (sm_100a) geometric hot path for point-cloud frame interpolation.

Host-side mirror of the reference's operator interface over the C ABI of libb200pc.so
(include/b200pc.h).  Import cost is nil; the shared object is loaded on first use and there is
no CPU fallback (a missing library or a CPU tensor raises).

  b200pc.pointnet2_utils   square_distance, index_points, farthest_point_sample,
                           query_ball_point, knn_point, three_nn, three_interpolate, group_points
  b200pc.pytorch3d_shim    knn_points, knn_gather, chamfer_distance
  b200pc.dropin.install()  run the unmodified reference models on these kernels
  b200pc.polypci           PolyPCI's per-point polynomial fit + evaluation as one device kernel
  b200pc.io                raw .bin sweeps -> fixed-size clouds by farthest point sampling on the device
  b200pc.synth             deterministic HDL-64-shaped synthetic sweeps (tests / bench)
"""
from . import _lib  # noqa: F401  (no dlopen at import)

__all__ = ["pointnet2_utils", "pytorch3d_shim", "dropin", "ops", "synth", "dist", "hostio", "io", "pointinet", "polypci"]
__version__ = "0.1.0"


def __getattr__(name):
    if name in __all__:
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
