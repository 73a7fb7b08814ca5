import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/point-cloud-interpolation-_b200")
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
for nodrain in (0, 1):
    if nodrain: os.environ["B200PC_DEBUG_NODRAIN"] = "1"
    for q, w in (("2", "7"), ("1", "14")):
        os.environ["B200PC_FORCE_Q"] = q; os.environ["B200PC_FORCE_WARPS"] = w
        for name, fn in (("knn16 form0", lambda: P.knn_point(16, ref, qry)), ("knn16 direct", lambda: ops.knn_search(ref, qry, 16, 2)), ("three_nn-ish form1 k3 16k", lambda: ops.knn_search(ref, qry, 3, 1))):
            ms = t(fn); print("nodrain=%d Q=%s W=%s %-26s %.3f ms (%.1f%%)" % (nodrain, q, w, name, ms, 8*16384*16384*8/ms/1e9/74.4*100), flush=True)
