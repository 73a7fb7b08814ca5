"""b200pc.dropin.install(): the reference's own modules end up bound to the CUDA entry points.
Mechanics only (no GPU here, no reference on the GPU box): build container only."""
import sys

import pytest
import torch

from oracle import ref_loader

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")


@needs_ref
def test_install_rebinds_reference_modules():
    import subprocess, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
from b200pc import dropin, pointnet2_utils as P, pytorch3d_shim as S
patched = dropin.install()
import Utils.Pointnet2Utils as RU, Utils.Layers as RL
assert "Utils.Pointnet2Utils" in patched and "Utils.Layers" in patched, patched
for n in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point"):
    assert getattr(RU, n) is getattr(P, n), n
    assert getattr(RL, n) is getattr(P, n), n
assert RL.knn_points is S.knn_points and RL.knn_gather is S.knn_gather
assert RL.Group.forward is dropin._group_forward and RL.FeaturePropagation.forward is dropin._fp_forward
assert RU.PointNetFeaturePropagation.forward is dropin._pnfp_forward
import pytorch3d.loss
assert pytorch3d.loss.chamfer_distance is S.chamfer_distance
# the reference's layer classes still construct, and calling them on CPU fails loudly (no fallback)
g = RL.Group(None, 4, knn=True)
try:
    g(torch.zeros(1, 3, 8), torch.zeros(1, 3, 2), torch.zeros(1, 1, 8))
    raise SystemExit("expected a RuntimeError: CPU tensors must be rejected")
except RuntimeError as e:
    assert "no CPU path" in str(e)
print("ok")
''' % (root, os.path.join(root, "point-cloud-interpolation-_b200"), ref_loader.REF_ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
