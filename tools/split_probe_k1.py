"""k = 1 searches of the C5 shapes: the planner's choice against forced ref splits (B200PC_DEBUG_PLAN=1 prints the plans)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/point-cloud-interpolation-_b200")
import numpy as np, torch
from b200pc import ops, synth
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
fr = synth.frame_pair(200, 65536)
for B, S in ((1, 65536), (4, 65536), (1, 32768), (4, 8192)):
    r = torch.from_numpy(np.stack([fr[0]] * B)).to(dev); q = torch.from_numpy(np.stack([fr[1][:S]] * B)).to(dev)
    base = None
    envs = ({},) if os.environ.get("ONLY_DEFAULT") else ({}, {"B200PC_FORCE_SPLIT": "2"}, {"B200PC_FORCE_SPLIT": "2", "B200PC_FORCE_WARPS": "14"}, {"B200PC_FORCE_SPLIT": "4"}, {"B200PC_FORCE_WARPS": "14"})
    for env in envs:
        for k_ in ("B200PC_FORCE_SPLIT", "B200PC_FORCE_WARPS"): os.environ.pop(k_, None)
        os.environ.update(env); ops.reload_tuning()
        out = ops.knn_search(r, q, 1, 2)
        if base is None: base = out
        print("B=%d S=%d %-50s %.3f ms same=%s" % (B, S, env, t(lambda: ops.knn_search(r, q, 1, 2)), torch.equal(out, base)), flush=True)
