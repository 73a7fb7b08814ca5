"""Time search variants on C2/C3 shapes for each forced CTA shape. usage: python tools/sweep.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
import b200pc.ops as _b200pc_ops; _b200pc_ops.TUNING_AUTORELOAD = True   # this probe flips B200PC_* knobs between calls (the library caches them)
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
known = ref[:, ::4].contiguous()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
cases = {
  "knn16": (lambda: P.knn_point(16, ref, qry), 8 * 16384 * 16384),
  "knn_direct16": (lambda: ops.knn_search(ref, qry, 16, 2), 8 * 16384 * 16384),
  "ball32": (lambda: P.query_ball_point(1.0, 32, ref, qry), 8 * 16384 * 16384),
  "three_nn": (lambda: P.three_nn(ref, known), 8 * 16384 * 4096),
  "nn1": (lambda: ops.knn_search(ref, qry, 1, 2), 8 * 16384 * 16384),
}
configs = [c.split("x") for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0x0", "2x7", "2x4", "1x8", "1x14", "1x4"])]
for name, (fn, pairs) in cases.items():
    for q, w in configs:
        os.environ["B200PC_FORCE_Q"] = q; os.environ["B200PC_FORCE_WARPS"] = w
        ms = t(fn)
        print("%-14s Q=%s W=%-2s  %.3f ms  %.1f TFLOP/s (%.1f%% of 74.4)" % (name, q, w, ms, pairs * 8 / ms / 1e9, pairs * 8 / ms / 1e9 / 74.4 * 100), flush=True)
