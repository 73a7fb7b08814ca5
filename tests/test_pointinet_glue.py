"""The PointINet caller (b200pc.pointinet) against the REAL upstream model of the reference
(PointINet20230424/models/models.py), same weights, same seed, on CPU with the torch stand-ins for
the geometric ops: checks layouts, parameter names and the order of CPU-RNG draws.  Build container
only (skipped where /root/reference is absent)."""
import numpy as np
import pytest
import torch

from b200pc import pointinet, synth
from oracle import ref_loader

from oracle import cpu_backend

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")


def _inputs(n=2048, extra=1):
    a, b = synth.frame_pair(5, n)
    g = torch.Generator().manual_seed(1)
    p1 = torch.cat([torch.from_numpy(a).t(), torch.rand(extra, n, generator=g)], 0).unsqueeze(0)
    p2 = torch.cat([torch.from_numpy(b).t(), torch.rand(extra, n, generator=g)], 0).unsqueeze(0)
    z = torch.zeros(1, 3, n)
    return p1.contiguous(), p2.contiguous(), z, z.clone()


def test_state_dict_keys_match_reference_naming():
    net = pointinet.PointINet(backend=cpu_backend.make())
    keys = set(net.state_dict().keys())
    for k in ("flow.set_conv1.conv.0.weight", "flow.flow_embedding.conv.3.bias", "flow.set_upconv1.conv2.0.weight",
              "flow.fp.conv.1.running_mean", "flow.classifier.3.weight", "fusion.conv.6.weight"):
        assert k in keys
    assert all(not p.requires_grad for p in net.flow.parameters())       # freeze=1, models.py:83-85
    assert all(p.requires_grad for p in net.fusion.parameters())


@needs_ref
@pytest.mark.parametrize("t", [0.5, 0.3])
def test_forward_equals_upstream_model_on_cpu(t):
    up = ref_loader.upstream_pointinet()
    torch.manual_seed(0)
    ref_net = up.PointINet(freeze=1).eval()
    mine = pointinet.PointINet(backend=cpu_backend.make()).eval()
    missing = mine.load_state_dict(ref_net.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    p1, p2, f1, f2 = _inputs()
    tt = torch.tensor([t], dtype=torch.float32)
    torch.manual_seed(3000)
    with torch.no_grad():
        want = ref_net(p1, p2, f1, f2, tt)
    torch.manual_seed(3000)
    with torch.no_grad():
        got = mine(p1, p2, f1, f2, tt)
    assert got.shape == want.shape == (1, 4, 2048)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@needs_ref
def test_flownet3d_equals_upstream_on_cpu():
    up = ref_loader.upstream_pointinet()
    torch.manual_seed(1)
    ref_net = up.FlowNet3D().eval()
    mine = pointinet.FlowNet3D(cpu_backend.make()).eval()
    mine.load_state_dict(ref_net.state_dict(), strict=True)
    p1, p2, f1, f2 = _inputs(extra=0)
    torch.manual_seed(7)
    with torch.no_grad():
        want = ref_net(p1, p2, f1, f2)
    torch.manual_seed(7)
    with torch.no_grad():
        got = mine(p1, p2, f1, f2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


def test_batched_points_fusion_equals_per_item_loop_cpu():
    # with one time stamp for the batch, the per-item loop of PointsFusion (upstream layers.py:389-411)
    # collapses into two batched searches; same RNG draws in the same order, so every item's output is unchanged
    torch.manual_seed(0)
    net = pointinet.PointINet(backend=cpu_backend.make()).eval()
    p1, p2, f1, f2 = _inputs()
    rep = lambda x, s: torch.cat([x, x.roll(s, dims=2) * 1.01], 0).contiguous()          # two different items
    P1, P2, F1, F2 = rep(p1, 7), rep(p2, 11), rep(f1, 0), rep(f2, 0)
    tt = torch.tensor([0.3, 0.3], dtype=torch.float32)
    outs = []
    for batched in (True, False):
        net.fusion.batched = batched
        torch.manual_seed(3000)
        with torch.no_grad():
            outs.append(net(P1, P2, F1, F2, tt))
    assert outs[0].shape == (2, 4, 2048)
    torch.testing.assert_close(outs[0], outs[1], rtol=1e-6, atol=1e-6)
    # different time stamps inside one batch keep the loop
    net.fusion.batched = True
    torch.manual_seed(3000)
    with torch.no_grad():
        mixed_t = net(P1, P2, F1, F2, torch.tensor([0.3, 0.6]))
    assert mixed_t.shape == (2, 4, 2048) and torch.isfinite(mixed_t).all()
