"""C2 knn_point under a few tuning settings (arguments: NAME=VALUE,... groups separated by spaces); prints ms per call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
base = None
for grp in sys.argv[1:] or ["B200PC_GRID=0"]:
    kv = dict(x.split("=") for x in grp.split(","))
    for k_, v in kv.items(): os.environ[k_] = v
    ops.reload_tuning()
    out = P.knn_point(16, ref, qry)
    if base is None: base = out
    print("%-50s %.3f ms  identical=%s" % (grp, t(lambda: P.knn_point(16, ref, qry)), torch.equal(out, base)), flush=True)
    for k_ in kv: os.environ.pop(k_)
