import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/point-cloud-interpolation-_b200")
import torch, numpy as np
from b200pc import ops, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
xyz16 = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, N, npt in ((1, 16384, 1024), (16, 16384, 4096), (1, 1024, 256), (1, 256, 64), (2, 16384, 1024), (1, 8192, 2048), (8, 4096, 1024)):
    x = xyz16[:B, :N].contiguous(); st = torch.zeros(B, dtype=torch.long, device=dev)
    ms = t(lambda: ops.fps(x, npt, st))
    ms2 = t(lambda: ops.fps(x, npt, st, want_xyz=True))
    print("FPS B=%2d N=%5d -> %4d : %.3f ms  %.3f us/round   (+coordinates: %.3f ms)" % (B, N, npt, ms, ms * 1e3 / npt, ms2), flush=True)
