"""three_interpolate at C3 under the tuning knobs (rows in flight, flat kernel)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
a, b = synth.batch_pairs(0, 8, 16384)
xyz = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
fidx = ops.fps(xyz, 4096, torch.arange(16, device=dev) * 7)
known = P.index_points(xyz, fidx)
_, i3, w3 = P.three_nn_weights(xyz, known)
def train(fn, n=10):
    def build(w):
        g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            flush_buf.zero_(); fn()
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    flush_buf.zero_()
                    if w: k = fn()
        torch.cuda.current_stream().wait_stream(s); return g
    def run(g):
        for _ in range(2): g.replay()
        torch.cuda.synchronize(); best = 1e9
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        return best
    return (run(build(True)) - run(build(False))) / n * 1e3
for C in (128, 256, 64):
    sf = torch.randn(16, 4096, C, device=dev)
    nbytes = 16 * 16384 * C * 4 + 16 * 4096 * C * 4 + 16 * 16384 * 36
    for env in ({}, {"B200PC_INTERP_ROWS": "1"}, {"B200PC_INTERP_ROWS": "4"}, {"B200PC_INTERP_FLAT": "1"}, {"B200PC_INTERP_FLAT": "0"}):
        for k in ("B200PC_INTERP_ROWS", "B200PC_INTERP_FLAT"): os.environ.pop(k, None)
        os.environ.update(env); ops.reload_tuning()
        us = train(lambda: P.three_interpolate(sf, i3, w3))
        print("C=%d %-28s %.1f us  %.2f of HBM" % (C, env or "default", us, nbytes / us / 1e3 / 6537.6), flush=True)
