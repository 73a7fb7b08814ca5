"""A/B of group_points: register path (group.cu, B200PC_BULK=0) vs the asynchronous-copy path (rowmove.cu, B200PC_BULK=1)
on the C3 shapes; index_points / three_interpolate are timed too (their asynchronous variants were measured with an earlier
version of this probe and removed: both columns now show the register kernels).  L2 flushed before every timed call."""
import json
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
HBM = 6537.6
if os.path.exists("MEASURED_PEAKS.json"):
    HBM = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", HBM))


def t(fn, n=10):
    for _ in range(3):
        flush_buf.zero_(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush_buf.zero_(); flush_buf.zero_(); flush_buf.zero_()      # ~150 us of GPU work: the CPU gets ahead of the stream
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e-3


def ab(name, fn, nbytes):
    res = {}
    outs = {}
    for mode in ("0", "1"):
        os.environ["B200PC_BULK"] = mode; ops.reload_tuning()
        outs[mode] = fn()
        s = t(fn)
        res[mode] = s
    same = torch.equal(outs["0"], outs["1"])
    print("%-44s regs %.1f us (%.0f GB/s, %.2f)  bulk %.1f us (%.0f GB/s, %.2f)  identical=%s" % (
        name, res["0"] * 1e6, nbytes / res["0"] / 1e9, nbytes / res["0"] / 1e9 / HBM,
        res["1"] * 1e6, nbytes / res["1"] / 1e9, nbytes / res["1"] / 1e9 / HBM, same), flush=True)
    os.environ.pop("B200PC_BULK")


B3, N = 16, 16384
a, b = synth.batch_pairs(0, 8, N)
xyz = torch.from_numpy(np.concatenate([a, b], 0)[:B3]).to(dev)
start = torch.arange(B3, device=dev, dtype=torch.long) * 7
fidx = ops.fps(xyz, 4096, start)
for C in (128, 256, 64, 32):
    feats = torch.randn(B3, N, C, device=dev)
    ab("index_points [16,16384,%d] by [16,4096]" % C, lambda: P.index_points(feats, fidx), B3 * 4096 * (C * 4 * 2 + 8))
feats = torch.randn(B3, N, 128, device=dev)
known = P.index_points(xyz, fidx)
gidx = P.knn_point(16, xyz, known)
ab("knn_gather [16,16384,128] by [16,4096,16]", lambda: P.index_points(feats, gidx), B3 * 4096 * 16 * (128 * 4 + 8) + B3 * N * 128 * 4)
for C in (128, 256):
    sfeat = torch.randn(B3, 4096, C, device=dev)
    _, i3, w3 = P.three_nn_weights(xyz, known)
    ab("three_interpolate 16x16384<-4096 C=%d" % C, lambda: P.three_interpolate(sfeat, i3, w3),
       B3 * N * C * 4 + B3 * 4096 * C * 4 + B3 * N * 36)
for D in (64, 128):
    gfeat = torch.randn(B3, N, D, device=dev)
    ab("group_points B=16 S=4096 K=16 D=%d" % D, lambda: P.group_points(xyz, known, gfeat, gidx),
       B3 * 4096 * 16 * (8 + 4 * (3 + D)) + B3 * N * 4 * (3 + D) + B3 * 4096 * 12)
ab("group_points SA-MSG layout D=64", lambda: P.group_points(xyz, known, gfeat[:, :, :64].contiguous(), gidx, xyz_first=False),
   B3 * 4096 * 16 * (8 + 4 * 67) + B3 * N * 4 * 67 + B3 * 4096 * 12)
# ragged / hostile indices: negative wrap and out-of-range rows on both paths
bad = fidx.clone(); bad[:, ::7] = -3; bad[:, 5::11] = N + 5
ab("index_points with wrapped / out-of-range rows", lambda: P.index_points(feats, bad), B3 * 4096 * (128 * 4 * 2 + 8))
gbad = gidx.clone(); gbad[:, ::5, 3] = N; gbad[:, 1::9, 0] = -1
ab("group_points with empty-ball sentinel rows", lambda: P.group_points(xyz, known, gfeat, gbad), 1)
