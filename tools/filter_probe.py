"""A/B of the two prefilters (B200PC_FILTER=0 broadcast refs, 1 lanes hold refs) with and without the drain"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
import b200pc.ops as _b200pc_ops; _b200pc_ops.TUNING_AUTORELOAD = True   # this probe flips B200PC_* knobs between calls (the library caches them)
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
shapes = [s.split(",") for s in (sys.argv[1] if len(sys.argv) > 1 else "2,7 1,14 1,8 2,4 1,16 2,8").split()]
for filt in ("0", "1"):
    os.environ["B200PC_FILTER"] = filt
    for nodrain in (0, 1):
        if nodrain: os.environ["B200PC_DEBUG_NODRAIN"] = "1"
        else: os.environ.pop("B200PC_DEBUG_NODRAIN", None)
        for q, w in shapes:
            os.environ["B200PC_FORCE_Q"] = q; os.environ["B200PC_FORCE_WARPS"] = w
            ms = t(lambda: P.knn_point(16, ref, qry))
            print("filter=%s nodrain=%d Q=%s W=%-2s knn16 form0 %.3f ms (%.1f%%)" % (filt, nodrain, q, w, ms, 8*16384*16384*8/ms/1e9/74.1*100), flush=True)
