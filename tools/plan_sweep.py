"""Forced {warps per CTA, ref split} against the planner's choice for one search shape (B200PC_DEBUG_PLAN=1 shows the plans).
usage: python tools/plan_sweep.py B N S k [form]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import ops, synth
B, N, S, k = (int(x) for x in sys.argv[1:5]); form = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
fr = [synth.frame_pair(700 + i, max(N, S)) for i in range(min(B, 2))]
ref = torch.from_numpy(np.stack([fr[i % len(fr)][0][:N] for i in range(B)])).to(dev)
qry = torch.from_numpy(np.stack([fr[i % len(fr)][1][:S] for i in range(B)])).to(dev)
def run(**env):
    for kk in ("B200PC_FORCE_SPLIT", "B200PC_FORCE_WARPS"): os.environ.pop(kk, None)
    for kk, v in env.items(): os.environ["B200PC_" + kk] = str(v)
    ops.reload_tuning()
    return t(lambda: ops.knn_search(ref, qry, k, form))
print("planner: %.3f ms" % run(), flush=True)
for split in (1, 2, 3, 4, 8, 16):
    print("split %2d: " % split + "  ".join("w=%d %.3f" % (w, run(FORCE_SPLIT=split, FORCE_WARPS=w)) for w in (2, 4, 7, 10, 14)), flush=True)
