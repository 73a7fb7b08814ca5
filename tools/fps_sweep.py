"""FPS 16384 points: two-level vs flat exchange for forced cluster sizes"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
import torch, numpy as np
from b200pc import ops, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
xyz16 = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (1, 4, 16):
    x = xyz16[:B].contiguous(); st = torch.arange(B, dtype=torch.long, device=dev) * 3
    for C in (2, 4, 8, 16):
        row = []
        for flat in ("0", "1"):
            os.environ["B200PC_FPS_FLAT"] = flat; os.environ["B200PC_FPS_CLUSTER"] = str(C); ops.reload_tuning()
            row.append(t(lambda: ops.fps(x, 1024, st)) * 1e3 / 1024)
        print("B=%2d C=%2d  two-level %.3f us/round   flat %.3f us/round" % (B, C, row[0], row[1]), flush=True)
