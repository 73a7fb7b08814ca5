"""A/B of the occupancy-grid variants of the top-k searches on the bench shapes: blind (B200PC_GRID=0), starting thresholds
only (2), refs and queries visited in cell order as well (3; B200PC_INTERLEAVE=1: warps dealt out over the CTAs instead of contiguous blocks) and
the default choice; checks that all return identical indices.  python tools/grid_probe.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from b200pc import ops, pointnet2_utils as P, synth
ops.TUNING_AUTORELOAD = True
dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def t(fn, n=10):
    for _ in range(3):
        flush_buf.zero_(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush_buf.zero_(); flush_buf.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n


def ab(name, fn, pairs):
    r = {}
    outs = {}
    modes = (("blind", "0", None, None), ("thresholds", "2", None, None), ("sorted", "3", None, None), ("sorted/interleaved", "3", None, "1"), ("default", None, None, None))
    for m, g, sd, dr in modes:
        for kk, v in (("B200PC_GRID", g), ("B200PC_SEED", sd), ("B200PC_INTERLEAVE", dr)):
            if v is None: os.environ.pop(kk, None)
            else: os.environ[kk] = v
        ops.reload_tuning()
        outs[m] = fn()
        r[m] = t(fn)
    def eq(x, y):
        return all(torch.equal(u, v) for u, v in zip(x, y)) if isinstance(x, tuple) else torch.equal(x, y)
    same = all(eq(outs["blind"], outs[m[0]]) for m in modes[1:])
    print("%-46s " % name + "  ".join("%s %.3f (%.1f%%)" % (m[0], r[m[0]], pairs * 8 / r[m[0]] / 1e9 / 74.1 * 100) for m in modes)
          + "  identical=%s" % same, flush=True)
    for kk in ("B200PC_GRID", "B200PC_SEED", "B200PC_INTERLEAVE"): os.environ.pop(kk, None)
    ops.reload_tuning()


a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
ab("knn_point k=16 8x16384x16384 (C2)", lambda: P.knn_point(16, ref, qry), 8 * 16384 * 16384)
ab("knn_points direct k=16, same", lambda: ops.knn_search(ref, qry, 16, ops.FORM_DIRECT, want_dist=True), 8 * 16384 * 16384)
ab("nearest neighbour k=1, same", lambda: ops.knn_search(ref, qry, 1, ops.FORM_DIRECT, want_dist=True), 8 * 16384 * 16384)
ab("k=64, same", lambda: P.knn_point(64, ref, qry), 8 * 16384 * 16384)
ab("fusion kNN B=1 16384x16384 k=16", lambda: ops.knn_search(ref[:1], qry[:1], 16, ops.FORM_DIRECT), 16384 * 16384)
x16 = torch.from_numpy(np.concatenate([a, b], 0)).to(dev)
fidx = ops.fps(x16, 4096, torch.arange(16, device=dev) * 7)
known = P.index_points(x16, fidx)
ab("three_nn 16x16384x4096 (C3)", lambda: P.three_nn(x16, known), 16 * 16384 * 4096)
fr = synth.frame_pair(200, 65536)
r65 = torch.from_numpy(fr[0][None]).to(dev); q65 = torch.from_numpy(fr[1][None]).to(dev)
ab("k=1 65536x65536 (C5, one frame)", lambda: ops.knn_search(r65, q65, 1, ops.FORM_DIRECT), 65536 * 65536)
srt = torch.from_numpy(np.ascontiguousarray(a[:, np.argsort(a[0, :, 0])])).to(dev)
ab("C2 with refs sorted along x", lambda: P.knn_point(16, srt, qry), 8 * 16384 * 16384)
far = qry + 500.0
ab("C2 with queries 500 m outside the ref box", lambda: P.knn_point(16, ref, far), 8 * 16384 * 16384)
