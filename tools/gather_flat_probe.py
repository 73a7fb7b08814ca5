"""A/B of the warp-per-row and the flat (thread-per-vector) gather / interpolate kernels at the C3 sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import pointnet2_utils as P
import b200pc.ops as _b200pc_ops; _b200pc_ops.TUNING_AUTORELOAD = True   # this probe flips B200PC_* knobs between calls (the library caches them)
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
B, N, S = 16, 16384, 4096
for C in (64, 128, 256):
    feats = torch.randn(B, N, C, device=dev); idx = torch.randint(0, N, (B, S), device=dev)
    sfeat = torch.randn(B, S, C, device=dev); i3 = torch.randint(0, S, (B, N, 3), device=dev); w3 = torch.rand(B, N, 3, device=dev)
    res = {}
    for flat in ("0", "1"):   # forced: warp-per-row, flat
        for k in ("B200PC_GATHER_FLAT", "B200PC_INTERP_FLAT"):
            if flat == "1": os.environ[k] = "1"
            else: os.environ[k] = "0"
        g = P.index_points(feats, idx); it = P.three_interpolate(sfeat, i3, w3)
        res[flat] = (g, it, t(lambda: P.index_points(feats, idx)), t(lambda: P.three_interpolate(sfeat, i3, w3)))
    assert torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])
    gb = B * S * (C * 8 + 8); ib = B * N * C * 4 + B * S * C * 4 + B * N * 36
    print("C=%d  index_points rows %.4f ms (%.0f GB/s) flat %.4f ms (%.0f GB/s) | interpolate rows %.4f ms (%.0f GB/s) flat %.4f ms (%.0f GB/s)" % (
        C, res["0"][2], gb / res["0"][2] / 1e6, res["1"][2], gb / res["1"][2] / 1e6, res["0"][3], ib / res["0"][3] / 1e6, res["1"][3], ib / res["1"][3] / 1e6), flush=True)
