"""time the HBM-bound kernels at C3 with an L2 flush before each call"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
B = 16
a, b = synth.batch_pairs(0, 8, 16384)
import numpy as np
xyz = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
start = torch.arange(B, device=dev) * 7
fidx = ops.fps(xyz, 4096, start)
for C in (128, 64, 256):
    feats = torch.randn(B, 16384, C, device=dev)
    ms = t(lambda: P.index_points(feats, fidx)); by = B * 4096 * (C * 4 * 2 + 8)
    print("index_points C=%3d  %.1f us  %.0f GB/s  %.1f%% of 6537.6" % (C, ms * 1e3, by / ms / 1e6, by / ms / 1e6 / 65.376))
    known = P.index_points(xyz, fidx); sfeat = P.index_points(feats, fidx)
    _, i3, w3 = P.three_nn_weights(xyz, known)
    ms = t(lambda: P.three_interpolate(sfeat, i3, w3)); by = B * 16384 * C * 4 + B * 4096 * C * 4 + B * 16384 * 36
    print("three_interp C=%3d  %.1f us  %.0f GB/s  %.1f%% of 6537.6" % (C, ms * 1e3, by / ms / 1e6, by / ms / 1e6 / 65.376))
big = torch.randn(B, 16384, 128, device=dev); kidx = torch.randint(0, 16384, (B, 16384, 16), device=dev)
ms = t(lambda: P.index_points(big, kidx), n=5); by = B * 16384 * 16 * (128 * 4 + 8) + B * 16384 * 128 * 4
print("grouped gather [16,16384,128]x[16,16384,16]  %.1f us  %.0f GB/s (out+idx+src)" % (ms * 1e3, by / ms / 1e6))
