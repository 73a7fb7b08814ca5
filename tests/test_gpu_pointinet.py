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
