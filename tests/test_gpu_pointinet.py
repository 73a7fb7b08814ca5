"""End-to-end PointINet forward on the CUDA kernels vs the same model (same weights, same CPU-RNG
seed) on the CPU torch port of the reference's primitives."""
import numpy as np
import pytest
import torch

from b200pc import pointinet, synth
from oracle import cpu_backend

pytestmark = pytest.mark.gpu


def _inputs(n, extra=1):
    a, b = synth.frame_pair(5, n)
    g = torch.Generator().manual_seed(1)
    p1 = torch.cat([torch.from_numpy(a).t(), torch.rand(extra, n, generator=g)], 0).unsqueeze(0).contiguous()
    p2 = torch.cat([torch.from_numpy(b).t(), torch.rand(extra, n, generator=g)], 0).unsqueeze(0).contiguous()
    z = torch.zeros(1, 3, n)
    return p1, p2, z, z.clone()


def test_pointinet_forward_gpu_matches_cpu_port(cuda_dev):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cpu_net = pointinet.PointINet(backend=cpu_backend.make()).eval()
    gpu_net = pointinet.PointINet().eval()
    gpu_net.load_state_dict(cpu_net.state_dict())
    gpu_net.to(cuda_dev)
    p1, p2, f1, f2 = _inputs(4096)
    t = torch.tensor([0.5])
    torch.manual_seed(3000)
    with torch.no_grad():
        want = cpu_net(p1, p2, f1, f2, t)
    torch.manual_seed(3000)
    with torch.no_grad():
        got = gpu_net(p1.to(cuda_dev), p2.to(cuda_dev), f1.to(cuda_dev), f2.to(cuda_dev), t.to(cuda_dev)).cpu()
    assert got.shape == want.shape == (1, 4, 4096)
    # the geometric decisions are exact; the MLPs differ by fp32 summation order (cuDNN vs MKL), which can
    # flip a near-tie in the fusion kNN for a handful of points
    err = (got - want).abs().amax(dim=1).reshape(-1)
    assert (err < 1e-3).float().mean() > 0.995, "only %.4f of the points agree" % (err < 1e-3).float().mean()
    assert torch.isfinite(got).all()


def test_flownet3d_gpu_matches_cpu_port(cuda_dev):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    cpu_net = pointinet.FlowNet3D(cpu_backend.make()).eval()
    gpu_net = pointinet.FlowNet3D().eval()
    gpu_net.load_state_dict(cpu_net.state_dict())
    gpu_net.to(cuda_dev)
    p1, p2, f1, f2 = _inputs(2048, extra=0)
    torch.manual_seed(7)
    with torch.no_grad():
        want = cpu_net(p1, p2, f1, f2)
    torch.manual_seed(7)
    with torch.no_grad():
        got = gpu_net(p1.to(cuda_dev), p2.to(cuda_dev), f1.to(cuda_dev), f2.to(cuda_dev)).cpu()
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4)


def test_graphed_pointinet_matches_eager(cuda_dev):
    """CUDA-graph replay with the RNG tape == eager forward under the same torch.manual_seed (no BN folding),
    and with BN folded within 1e-4."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    eager = pointinet.PointINet().eval().to(cuda_dev)
    sd = eager.state_dict()
    ins = [x.to(cuda_dev) for x in _inputs(4096)]
    t = torch.tensor([0.5], device=cuda_dev)
    for fold, tol in ((False, 1e-6), (True, 2e-4)):
        g = pointinet.GraphedPointINet(state_dict=sd, batch=1, npoints=4096, extra=1, t=0.5, device=cuda_dev, fold_bn=fold)
        g.capture(*ins)
        for seed in (11, 12):
            torch.manual_seed(seed)
            with torch.no_grad():
                want = eager(*ins, t)
            torch.manual_seed(seed)
            got = g(*ins).clone()
            err = (got - want).abs().amax(dim=1).reshape(-1)
            assert (err < max(tol, 1e-6) * 50).float().mean() > 0.995, (fold, seed, float(err.max()))


def test_flownet3d_training_step_gradients_match_cpu_port(cuda_dev):
    """BASELINE config 4 in miniature (train_sceneflow.py:132-185 in the reference): FlowNet3D forward in
    train mode, loss = chamfer(p1 + flow, p2), backward.  Gradients through index_points / three_interpolate /
    chamfer_distance (custom ops) must match autograd through the torch port on CPU."""
    from b200pc import pytorch3d_shim as S3
    from oracle import ref_torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    cpu_net = pointinet.FlowNet3D(cpu_backend.make()).train()
    gpu_net = pointinet.FlowNet3D().train()
    gpu_net.load_state_dict(cpu_net.state_dict())
    gpu_net.to(cuda_dev)
    a, b = synth.batch_pairs(8, 2, 2048)
    p1 = torch.from_numpy(a).transpose(1, 2).contiguous(); p2 = torch.from_numpy(b).transpose(1, 2).contiguous()
    f = torch.zeros(2, 3, 2048)
    torch.manual_seed(9)
    flow_c = cpu_net(p1, p2, f, f)
    loss_c = ref_torch.chamfer_dense((p1 + flow_c).transpose(1, 2), p2.transpose(1, 2))
    loss_c.backward()
    torch.manual_seed(9)
    d = lambda x: x.to(cuda_dev)
    flow_g = gpu_net(d(p1), d(p2), d(f), d(f))
    loss_g, _ = S3.chamfer_distance((d(p1) + flow_g).permute(0, 2, 1), d(p2).permute(0, 2, 1))
    loss_g.backward()
    assert abs(loss_g.item() - loss_c.item()) <= 1e-3 * abs(loss_c.item())
    checked = 0
    all_c, all_g = [], []
    for (n, pc), (_, pg) in zip(cpu_net.named_parameters(), gpu_net.named_parameters()):
        if pc.grad is None:
            continue
        # a conv bias directly in front of a train-mode BatchNorm has an exactly-zero true gradient (the batch
        # mean removes it): both sides only hold cancellation noise there, nothing to compare
        import re
        if re.search(r"conv\d?\.(0|3|6)\.bias$", n) or n == "classifier.0.bias":
            continue
        gc, gg = pc.grad.reshape(-1), pg.grad.cpu().reshape(-1)
        # ReLU / max-over-neighbours make the gradient piecewise: fp32 summation-order differences between cuDNN and
        # MKL can flip an arg-max in the deep layers (train-mode BatchNorm over 256 samples), which moves a few entries
        # by a visible amount.  So the per-tensor criterion is directional (cosine), the global one is tight.
        cos = torch.dot(gc, gg) / (gc.norm() * gg.norm() + 1e-30)
        assert cos.item() > 0.97, (n, cos.item())
        all_c.append(gc); all_g.append(gg)
        checked += 1
    all_c, all_g = torch.cat(all_c), torch.cat(all_g)
    assert (torch.dot(all_c, all_g) / (all_c.norm() * all_g.norm())).item() > 0.999
    assert abs(all_c.norm().item() - all_g.norm().item()) <= 2e-2 * all_c.norm().item()
    assert checked > 40


def test_batched_points_fusion_equals_per_item_loop_gpu(cuda_dev):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = pointinet.PointINet().eval().to(cuda_dev)
    p1, p2, f1, f2 = _inputs(4096)
    rep = lambda x, s: torch.cat([x, x.roll(s, dims=2) * 1.01, x.roll(2 * s + 1, dims=2) * 0.99], 0).contiguous().to(cuda_dev)
    P1, P2, F1, F2 = rep(p1, 7), rep(p2, 11), rep(f1, 0), rep(f2, 0)
    tt = torch.tensor([0.5, 0.5, 0.5], device=cuda_dev)
    outs = []
    for batched in (True, False):
        net.fusion.batched = batched
        torch.manual_seed(3000)
        with torch.no_grad():
            outs.append(net(P1, P2, F1, F2, tt))
    assert outs[0].shape == (3, 4, 4096)
    # identical neighbour sets; the MLP runs on one [3,...] tensor instead of three [1,...] ones (cuDNN summation order)
    err = (outs[0] - outs[1]).abs().amax(dim=1).reshape(-1)
    assert (err < 1e-4).float().mean() > 0.999, "only %.4f of the points agree" % (err < 1e-4).float().mean()
