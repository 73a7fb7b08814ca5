"""The UNMODIFIED reference models on the B200 through b200pc.dropin.install().

The reference's own Python files (staged byte for byte under oracle/_ref/ by oracle/make_ref.py, or /root/reference in
the build container) are imported as they are.  Each test runs the real module twice with the same weights and the same
CPU-RNG seed: once on the CPU with the reference's own primitives (Utils/Pointnet2Utils.py, models/pointnet2_utils.py;
pytorch3d calls served by oracle/strict.c) -- the checker -- and once on cuda:0 after `dropin.install()`, where every
geometric primitive is a b200pc kernel.  Neighbour selection depends on coordinates only, so in FlowNet3D and the
feature abstractor the two arms pick identical neighbours and the outputs differ by conv rounding alone (TF32 is off).
"""
import sys

import numpy as np
import pytest
import torch

from b200pc import dropin, synth
from oracle import ref_loader

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference not staged (python oracle/make_ref.py)")


@pytest.fixture(autouse=True)
def _exact_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture()
def installed():
    """reference modules imported with the CPU providers first; dropin installed for the body; undone afterwards"""
    ref_loader.rebind_pytorch3d(ref_loader.strict_provider())
    state = {"on": False}

    def switch(on):
        if on and not state["on"]:
            dropin.install(force_pytorch3d=True)
        elif not on and state["on"]:
            dropin.uninstall()
            ref_loader.rebind_pytorch3d(ref_loader.strict_provider())
        state["on"] = on
    yield switch
    switch(False)


def _pair(seed, n, extra=0):
    a, b = synth.frame_pair(seed, n)
    g = torch.Generator().manual_seed(seed)
    mk = lambda f: torch.cat([torch.from_numpy(f).t()] + ([torch.rand(extra, n, generator=g)] if extra else []), 0).unsqueeze(0).contiguous()
    return mk(a), mk(b)


def _close(got, want, rtol, atol, frac=1.0):
    got = got.detach().cpu().float(); want = want.detach().cpu().float()
    assert got.shape == want.shape
    ok = (got - want).abs() <= atol + rtol * want.abs()
    assert ok.float().mean().item() >= frac, "only %.4f%% of the elements agree (max abs diff %.3g)" % (
        100 * ok.float().mean().item(), (got - want).abs().max().item())


@needs_ref
def test_upstream_flownet3d_forward(cuda_dev, installed):
    """PointINet20230424/models/models.py:9-77 (FlowNet3D): SetConv x4 (FPS, ball query, grouping), FlowEmbedding and
    SetUpConv x3 (kNN grouping), FeaturePropagation (three-NN interpolation)."""
    up = ref_loader.upstream_pointinet(ref_loader.strict_provider())
    torch.manual_seed(0)
    net = up.FlowNet3D().eval()
    p1, p2 = _pair(11, 4096)
    f = torch.zeros(1, 3, 4096)
    torch.manual_seed(3000)
    with torch.no_grad():
        want = net(p1, p2, f, f)
    installed(True)
    assert sys.modules["models.layers"].Group.forward is dropin._group_forward
    net = net.to(cuda_dev)
    torch.manual_seed(3000)
    with torch.no_grad():
        got = net(p1.to(cuda_dev), p2.to(cuda_dev), f.to(cuda_dev), f.to(cuda_dev))
    _close(got, want, rtol=1e-3, atol=1e-4)


@needs_ref
def test_upstream_pointinet_forward(cuda_dev, installed):
    """PointINet20230424/models/models.py:79-125: two FlowNet3D passes, warp, PointsFusion (layers.py:335-430) whose
    knn_group runs as b200pc_fusion_group.  The fusion searches run on warped points that differ between the arms by
    conv rounding, so a few near-tied neighbours may swap: 99 % of the output coordinates within 1e-3."""
    up = ref_loader.upstream_pointinet(ref_loader.strict_provider())
    torch.manual_seed(0)
    net = up.PointINet(freeze=1).eval()
    p1, p2 = _pair(12, 4096, extra=1)
    f = torch.zeros(1, 3, 4096)
    t = torch.tensor([0.5])
    torch.manual_seed(3001)
    with torch.no_grad():
        want = net(p1, p2, f, f, t)
    installed(True)
    assert hasattr(sys.modules["models.layers"].PointsFusion.knn_group, "_b200pc_original")
    net = net.to(cuda_dev)
    torch.manual_seed(3001)
    with torch.no_grad():
        got = net(p1.to(cuda_dev), p2.to(cuda_dev), f.to(cuda_dev), f.to(cuda_dev), t.to(cuda_dev))
    assert got.shape == (1, 4, 4096)
    _close(got, want, rtol=1e-3, atol=1e-3, frac=0.99)


@needs_ref
def test_fork_feature_abstractor_and_transformer_64000_points(cuda_dev, installed):
    """Utils/Layers.py:498-530 (Pointnet2FeatureAbstract: SA-MSG x4 = FPS 64000->1024, ball queries r=0.1..1.6 through
    group_points(xyz_first=False); PointNetFeaturePropagation x4 incl. 64000<-1024 three-NN) and :405-445
    (TransformerLayer: self-kNN K=16 over 64000 points + knn_gather of 64-channel features), the ISAPCInet shapes
    (Models/New_Models0.py:170-182: 2*field*N = 64000 points)."""
    lay = ref_loader.layers(ref_loader.strict_provider())
    n = 64000
    a, _ = synth.frame_pair(13, 65536)
    flow = torch.from_numpy(a[:n]).t().unsqueeze(0).contiguous() * 0.05          # flow-sized coordinates, metres
    torch.manual_seed(1)
    ffab = lay.Pointnet2FeatureAbstract(64).eval()
    tr = lay.TransformerLayer(64, 64, 16).eval()
    torch.manual_seed(3002)
    with torch.no_grad():
        want_f = ffab(flow)
        want_t, want_attn = tr(flow, want_f)
    installed(True)
    ffab = ffab.to(cuda_dev); tr = tr.to(cuda_dev)
    torch.manual_seed(3002)
    with torch.no_grad():
        got_f = ffab(flow.to(cuda_dev))
        got_t, got_attn = tr(flow.to(cuda_dev), want_f.to(cuda_dev))            # same features in: isolates the layer
    # 64 channels x 64000 points through ~20 conv / GroupNorm / ReLU layers: cuDNN-vs-MKL rounding, a handful of outliers
    _close(got_f, want_f, rtol=1e-3, atol=1e-4, frac=0.9999)
    _close(got_t, want_t, rtol=1e-3, atol=1e-4, frac=0.9999)
    _close(got_attn, want_attn, rtol=1e-3, atol=1e-5, frac=0.9999)


@needs_ref
def test_fork_points_fusion_and_knn_group_withI(cuda_dev, installed):
    """Utils/Layers.py:195-283 (the fork's FPS-based PointsFusion) and :384-402 (knn_group_withI) on identical inputs."""
    lay = ref_loader.layers(ref_loader.strict_provider())
    p1, p2 = _pair(14, 4000)
    torch.manual_seed(2)
    fusion = lay.PointsFusion([64, 64, 128]).eval()
    t = torch.tensor([0.4])
    inten = torch.rand(1, 1, 4000)
    torch.manual_seed(3003)
    with torch.no_grad():
        want = fusion(p1, p2, 4000, 32, t)
        want_i = lay.knn_group_withI(p1, p2, inten, 8)
    installed(True)
    fusion = fusion.to(cuda_dev)
    torch.manual_seed(3003)
    with torch.no_grad():
        got = fusion(p1.to(cuda_dev), p2.to(cuda_dev), 4000, 32, t.to(cuda_dev))
        got_i = lay.knn_group_withI(p1.to(cuda_dev), p2.to(cuda_dev), inten.to(cuda_dev), 8)
    _close(got, want, rtol=1e-4, atol=1e-4)
    for g, w in zip(got_i, want_i):
        _close(g, w, rtol=1e-5, atol=1e-6)


@needs_ref
def test_feature_propagation_gradient_reaches_the_coordinates(cuda_dev, installed):
    """d loss / d xyz through PointNetFeaturePropagation (Utils/Pointnet2Utils.py:297-304) and FeaturePropagation
    (Utils/Layers.py:180-188): in the reference `weight = 1/dists` is differentiable w.r.t. both clouds, and ISAPCInet
    trains through it (Models/New_Models0.py:164-172).  Checked against the reference's own autograd on the CPU."""
    lay = ref_loader.layers(ref_loader.strict_provider())
    pu = ref_loader.pointnet2_utils()
    g = torch.Generator().manual_seed(5)
    dense = torch.randn(2, 3, 600, generator=g); sparse = dense[:, :, ::6].clone() + 0.01 * torch.randn(2, 3, 100, generator=g)
    f_sparse = torch.randn(2, 16, 100, generator=g); f_dense = torch.randn(2, 8, 600, generator=g)
    torch.manual_seed(3)
    pnfp = pu.PointNetFeaturePropagation(16 + 8, [32]).train()
    fp = lay.FeaturePropagation(16, 8, [32]).train()

    def run(dev):
        d = dense.clone().to(dev).requires_grad_(True); s = sparse.clone().to(dev).requires_grad_(True)
        fs = f_sparse.clone().to(dev).requires_grad_(True)
        out1 = pnfp.to(dev)(d, s, f_dense.to(dev), fs)
        out2 = fp.to(dev)(s, d, fs, f_dense.to(dev))
        loss = (out1 ** 2).mean() + (out2 ** 2).mean()
        loss.backward()
        return loss.detach().cpu(), d.grad.cpu(), s.grad.cpu(), fs.grad.cpu()

    want = run("cpu")
    installed(True)
    got = run(cuda_dev)
    assert want[1].abs().max() > 0 and want[2].abs().max() > 0            # the reference does send gradient to both clouds
    for gt, wt in zip(got, want):
        _close(gt, wt, rtol=2e-3, atol=1e-5 * float(wt.abs().max()) + 1e-8)


@needs_ref
def test_fork_isapcinet_forward_and_backward(cuda_dev, installed):
    """Models/New_Models0.py:90-195 (ISAPCInet, field=1) end to end on the drop-in: 4 FlowNet3D passes, T-nets, the
    feature abstractor on 2*field*N points, two transformer layers, warp, FPS-based PointsFusion, chamfer_loss
    (Utils/Utils.py:39-48) and a backward pass through all of it.  Smoke + finiteness on the GPU arm (the CPU arm of the
    pieces is covered above); gradients must reach the flow network's first layer."""
    installed(True)
    models = ref_loader.fork_models()
    losses = ref_loader.utils_losses()
    torch.manual_seed(4)
    net = models.ISAPCInet(field=1).train().to(cuda_dev)
    n = 2048
    frames = [torch.from_numpy(synth.frame_pair(20 + i, n)[0]).t().unsqueeze(0).contiguous().to(cuda_dev) for i in range(4)]
    ini = torch.zeros(1, 3, n, device=cuda_dev)
    t = torch.tensor([0.5], device=cuda_dev)
    torch.manual_seed(3004)
    out = net([frames[0]], [frames[1], frames[2]], [frames[3]], t, ini)
    assert out.shape == (1, 3, n) and torch.isfinite(out).all()
    loss = losses.chamfer_loss(out, frames[1])
    loss.backward()
    g = net.flow.set_conv1.conv[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0


@needs_ref
def test_polypci_forward(cuda_dev, installed):
    """PolyPCI/Models/Models_V1.py:92-222 (PolyPCI, field=2): four FlowNet3D passes, `rebuild` = knn_points(K=1, return_nn)
    (:102-114) after every warp, then the reference's own host-side polynomial fit (:116-124, :191-219 -- left as it is: the
    forward hard-codes the .cpu() / numpy / .cuda() round trip).  CPU arm = the same module with the reference's primitives;
    `rebuild` picks nearest neighbours of warped points that differ between the arms by conv rounding, so a few picks may
    swap: 99 % of the output coordinates within 1e-3."""
    import contextlib, io
    poly = ref_loader.polypci_models(ref_loader.strict_provider())
    torch.manual_seed(6)
    net = poly.PolyPCI(field=2, degree=2).eval()
    n = 2048
    frames = [torch.from_numpy(synth.frame_pair(60 + i, n)[0]).t().unsqueeze(0).contiguous() for i in range(5)]
    ini = torch.zeros(1, 3, n); t = torch.tensor([0.5]); T = [[0.0, -1.0, 1.0, -2.0, 2.0]]
    torch.manual_seed(3005)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):      # the reference prints every shape
        want = net([frames[1], frames[3]], frames[0], [frames[2], frames[4]], t, T, ini)
    installed(True)
    net = net.to(cuda_dev)
    fr = [f.to(cuda_dev) for f in frames]
    torch.manual_seed(3005)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        got = net([fr[1], fr[3]], fr[0], [fr[2], fr[4]], t.to(cuda_dev), T, ini.to(cuda_dev))
    assert got.shape == (1, 3, n) and got.is_cuda
    _close(got, want, rtol=1e-3, atol=1e-3, frac=0.99)
    # the device-side fit (b200pc.polypci) on the frames the reference stacked would give the same frame:
    # checked directly against the reference's fitting_and_predict in tests/test_gpu_ops.py
