"""time searches at the PointINet / FlowNet3D shapes (B=1) to check the planner"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
cases = [
 ("fusion knn_points k16 B1 16384x16384", lambda: ops.knn_search(ref[:1], qry[:1], 16, 2, want_dist=True), 16384*16384),
 ("ball r.5 ns16 B1 1024q x 16384", lambda: P.query_ball_point(0.5, 16, ref[:1], qry[:1, :1024].contiguous()), 1024*16384),
 ("ball r1 ns16 B1 256q x 1024", lambda: P.query_ball_point(1.0, 16, ref[:1, :1024].contiguous(), qry[:1, :256].contiguous()), 256*1024),
 ("three_nn B1 16384 <- 1024", lambda: P.three_nn(ref[:1], qry[:1, :1024].contiguous()), 1024*16384),
 ("knn k64 256x256", lambda: P.knn_point(64, ref[:1, :256].contiguous(), qry[:1, :256].contiguous()), 65536),
 ("knn k16 C2 B8", lambda: P.knn_point(16, ref, qry), 8*16384*16384),
 ("ball C2 B8", lambda: P.query_ball_point(1.0, 32, ref, qry), 8*16384*16384),
 ("chamfer 4x8192", lambda: ops.chamfer(ref[:4, :8192].contiguous(), qry[:4, :8192].contiguous()), 2*4*8192*8192),
 ("knn k1 C5-like B1 65536x65536", None, 0),
]
for name, fn, pairs in cases:
    if fn is None: continue
    ms = t(fn)
    print("%-40s %.3f ms  %.1f TFLOP/s" % (name, ms, pairs * 8 / ms / 1e9), flush=True)
import numpy as np
big = torch.from_numpy(np.concatenate([a[i] for i in range(4)], 0)[None]).to(dev)   # 65536 pts
ms = t(lambda: ops.knn_search(big, big, 1, 2, want_dist=True), n=3)
print("%-40s %.3f ms  %.1f TFLOP/s" % ("knn k1 B1 65536x65536 (C5 rebuild)", ms, 65536*65536*8/ms/1e9))
