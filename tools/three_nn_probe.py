"""three_nn at C3 (16 x 16384 queries against 4096 FPS picks, k=3) under the grid variants: thresholds (default) 0.320 ms, blind 0.341,
cell order 0.342, cell order + box-walking thresholds 0.36-0.39 -- ncu: the lock-step re-test of hit chunks is 14 % of the instructions at 9
active lanes, but neither ordering nor tighter thresholds pay for their extra kernels on a 0.3 ms search."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/point-cloud-interpolation-_b200")
import numpy as np, torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
a, b = synth.batch_pairs(0, 8, 16384)
x16 = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
fidx = ops.fps(x16, 4096, torch.arange(16, device=dev) * 7)
known = P.index_points(x16, fidx)
base = None
for env in ({}, {"B200PC_GRID": "0"}, {"B200PC_GRID": "3"}, {"B200PC_GRID": "3", "B200PC_SEED": "2"}, {"B200PC_GRID": "3", "B200PC_SEED": "5"}, {"B200PC_GRID": "3", "B200PC_SEED": "5", "B200PC_NATURAL_ORDER": "1"}):
    for k_ in ("B200PC_GRID", "B200PC_SEED", "B200PC_NATURAL_ORDER"): os.environ.pop(k_, None)
    os.environ.update(env); ops.reload_tuning()
    out = P.three_nn(x16, known)
    if base is None: base = out
    print("%-70s %.3f ms same=%s" % (env, t(lambda: P.three_nn(x16, known)), all(torch.equal(u, v) for u, v in zip(out, base))), flush=True)
