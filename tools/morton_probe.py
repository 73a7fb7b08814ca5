"""does a spatially coherent QUERY order help the lock-step drain?  time knn_point on the C2 batch with the queries in
their random order, Morton-sorted, and Morton-sorted refs as well. usage: python tools/morton_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import pointnet2_utils as P, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
def morton(p, cell=0.25):
    q = np.clip(((p + 128.0) / cell).astype(np.int64), 0, (1 << 10) - 1)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(np.clip(q[..., 2], 0, 1023)) << 2)
def sort_rows(x, cell):
    out = np.empty_like(x)
    for i in range(x.shape[0]): out[i] = x[i][np.argsort(morton(x[i], cell), kind="stable")]
    return out
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
print("random queries, random refs      : %.3f ms" % t(lambda: P.knn_point(16, ref, qry)), flush=True)
for cell in (1.0, 0.25):
    qs = torch.from_numpy(sort_rows(b, cell)).to(dev)
    print("Morton queries (cell %.2f m)     : %.3f ms" % (cell, t(lambda: P.knn_point(16, ref, qs))), flush=True)
def cell_rows(x, cell):       # counting-sort order: row-major 2-D cells, arbitrary order inside a cell
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        cx = np.floor((x[i][:, 0] + 100.0) / cell).astype(np.int64); cy = np.floor((x[i][:, 1] + 100.0) / cell).astype(np.int64)
        out[i] = x[i][np.argsort(cy * 4096 + cx, kind="stable")]
    return out
for cell in (8.0, 4.0, 2.0):
    qc = torch.from_numpy(cell_rows(b, cell)).to(dev)
    print("row-major %.0f m cells (queries)   : %.3f ms" % (cell, t(lambda: P.knn_point(16, ref, qc))), flush=True)
rs = torch.from_numpy(sort_rows(a, 0.25)).to(dev)
print("Morton queries + Morton refs     : %.3f ms" % t(lambda: P.knn_point(16, rs, qs)), flush=True)
print("random queries + Morton refs     : %.3f ms" % t(lambda: P.knn_point(16, rs, qry)), flush=True)
print("ball: random %.3f ms, Morton queries %.3f ms" % (t(lambda: P.query_ball_point(1.0, 32, ref, qry)), t(lambda: P.query_ball_point(1.0, 32, ref, qs))), flush=True)
