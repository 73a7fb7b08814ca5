"""fused grouping vs the unfused op sequence on a few shapes. usage: python tools/group_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import pointnet2_utils as P
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
for B, N, S, K, D in ((16, 16384, 4096, 16, 64), (16, 16384, 4096, 16, 128), (1, 16384, 1024, 16, 3), (1, 256, 256, 64, 128), (4, 8192, 2048, 16, 64), (16, 16384, 4096, 32, 13)):
    xyz = torch.randn(B, N, 3, device=dev); new = xyz[:, :S].contiguous(); feat = torch.randn(B, N, D, device=dev)
    idx = torch.randint(0, N, (B, S, K), device=dev)
    def unfused():
        rel = P.index_points(xyz, idx) - new.view(B, S, 1, 3)
        return torch.cat([rel, P.index_points(feat, idx)], dim=-1).permute(0, 3, 2, 1).contiguous()
    assert torch.equal(P.group_points(xyz, new, feat, idx), unfused())
    f = t(lambda: P.group_points(xyz, new, feat, idx)); u = t(unfused)
    gb = B * S * K * (8 + 4 * (3 + D)) + B * N * 4 * (3 + D) + B * S * 12
    print("B=%d N=%d S=%d K=%d D=%d: fused %.4f ms (%.0f GB/s algorithmic), unfused %.4f ms" % (B, N, S, K, D, f, gb / f / 1e6, u), flush=True)
