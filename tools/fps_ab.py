"""A/B of the two FPS kernels (B200PC_FPS_FLAT=0: two-level arg-max; 1: flat exchange of warp keys): time per round and
identical picks.  python tools/fps_ab.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
import torch, numpy as np
from b200pc import ops, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
xyz16 = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
big = torch.from_numpy(synth.frame_pair(3, 65536)[0][None]).to(dev)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, N, npt in ((1, 16384, 1024), (4, 16384, 1024), (16, 16384, 4096), (1, 1024, 256), (1, 256, 64), (1, 8192, 2048), (8, 4096, 1024), (1, 65536, 1024), (1, 40000, 512)):
    x = (big[:, :N] if N > 16384 else xyz16[:B, :N]).contiguous(); st = torch.arange(B, dtype=torch.long, device=dev) * 3
    res = {}; outs = {}
    for flat in ("0", "1"):
        os.environ["B200PC_FPS_FLAT"] = flat; ops.reload_tuning()
        outs[flat] = ops.fps(x, npt, st)
        res[flat] = t(lambda: ops.fps(x, npt, st))
    print("FPS B=%2d N=%5d -> %4d : two-level %.3f ms (%.3f us/round)   flat %.3f ms (%.3f us/round)   identical=%s" % (
        B, N, npt, res["0"], res["0"] * 1e3 / npt, res["1"], res["1"] * 1e3 / npt, torch.equal(outs["0"], outs["1"])), flush=True)
