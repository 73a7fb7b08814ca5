"""Where does the cell-ordered variant of the top-k search start to pay?  blind / thresholds (B200PC_GRID=2) / sorted (=3) / default
on a ladder of shapes.  usage: python tools/grid_crossover.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import ops, synth
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
SHAPES = [(1, 16384, 16384, 16), (2, 16384, 16384, 16), (4, 16384, 16384, 16), (8, 16384, 16384, 8), (8, 16384, 16384, 32),
          (8, 8192, 8192, 16), (32, 8192, 8192, 16), (32, 8192, 8192, 32), (1, 65536, 65536, 16), (4, 65536, 65536, 16), (16, 16384, 4096, 16),
          (8, 16384, 16384, 3), (8, 16384, 16384, 4)]
for B, N, S, k in SHAPES:
    fr = [synth.frame_pair(300 + i, max(N, S)) for i in range(min(B, 4))]
    ref = torch.from_numpy(np.stack([fr[i % len(fr)][0][:N] for i in range(B)])).to(dev)
    qry = torch.from_numpy(np.stack([fr[i % len(fr)][1][:S] for i in range(B)])).to(dev)
    r = {}
    for name, g in (("blind", "0"), ("thresholds", "2"), ("sorted", "3"), ("default", None)):
        if g is None: os.environ.pop("B200PC_GRID", None)
        else: os.environ["B200PC_GRID"] = g
        ops.reload_tuning()
        r[name] = t(lambda: ops.knn_search(ref, qry, k, 0))
    print("B=%2d N=%5d S=%5d k=%2d (2^%.0f pairs)  " % (B, N, S, k, np.log2(float(B) * N * S)) + "  ".join("%s %.3f" % kv for kv in r.items()), flush=True)
