"""Forced CTA shapes / ref splits on the B=1 shapes of a PointINet forward. usage: python tools/sweep_b1.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
import b200pc.ops as _b200pc_ops; _b200pc_ops.TUNING_AUTORELOAD = True   # this probe flips B200PC_* knobs between calls (the library caches them)
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 1, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
r4, q4 = ref[:, :4096].contiguous(), qry[:, :4096].contiguous()
r1, q1 = ref[:, :1024].contiguous(), qry[:, :1024].contiguous()
cases = [
 ("fusion direct k16 16384x16384", lambda: ops.knn_search(ref, qry, 16, 2, want_dist=True)),
 ("knn_point k16 4096q x 4096", lambda: P.knn_point(16, r4, q4)),
 ("knn_point k16 1024q x 4096", lambda: P.knn_point(16, r4, q1)),
 ("ball ns16 1024q x 16384", lambda: P.query_ball_point(0.5, 16, ref, q1)),
 ("three_nn 16384 <- 4096", lambda: P.three_nn(ref, r4)),
]
shapes = ["0,0,0", "1,14,8", "1,14,4", "1,8,4", "1,8,2", "1,4,2", "1,4,1", "1,7,2", "1,7,4", "2,7,4", "2,4,2", "1,2,1", "1,14,2"]
for name, fn in cases:
    for sh in shapes:
        q, w, sp = sh.split(",")
        for k, v in (("B200PC_FORCE_Q", q), ("B200PC_FORCE_WARPS", w), ("B200PC_FORCE_SPLIT", sp)):
            if v == "0": os.environ.pop(k, None)
            else: os.environ[k] = v
        print("%-32s Q=%s W=%-2s split=%s  %.3f ms" % (name, q, w, sp, t(fn)), flush=True)
