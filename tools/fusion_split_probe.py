import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/point-cloud-interpolation-_b200")
import torch
from b200pc import ops, synth
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
for q, w, sp in (("0","0","0"), ("1","2","1"), ("1","4","1"), ("1","4","2"), ("1","4","4"), ("1","8","4"), ("2","4","4"), ("1","8","8"), ("2","2","2"), ("1","2","2"), ("1","3","3"), ("2","4","8")):
    for k_, v in (("B200PC_FORCE_Q", q), ("B200PC_FORCE_WARPS", w), ("B200PC_FORCE_SPLIT", sp)):
        if v == "0": os.environ.pop(k_, None)
        else: os.environ[k_] = v
    ms = t(lambda: ops.knn_search(ref, qry, 16, 2, want_dist=True))
    print("Q=%s W=%s split=%s : %.3f ms" % (q, w, sp, ms), flush=True)
