"""C2 ball query (r=1, nsample=32): queries in the caller's order against queries pre-sorted on the host by grid cell -- what
would a cell order of the queries buy the ball query (warps whose 32 queries are all full skip the rest of the refs)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import ops, pointnet2_utils as P, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
def cell_sorted(q, r):
    out = np.empty_like(q); perm = []
    for i in range(q.shape[0]):
        lo = r[i].min(0); h = (r[i].max(0) - lo).max() / 128
        c = np.clip(((q[i] - lo) / h).astype(np.int64), 0, 127)
        o = np.argsort((c[:, 2] * 128 + c[:, 1]) * 128 + c[:, 0], kind="stable")
        out[i] = q[i][o]; perm.append(o)
    return out, perm
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
qs_np, perm = cell_sorted(b, a)
qs = torch.from_numpy(qs_np).to(dev)
for r, ns in ((1.0, 32), (0.5, 16), (2.0, 64)):
    base = P.query_ball_point(r, ns, ref, qry); srt = P.query_ball_point(r, ns, ref, qs)
    same = all(torch.equal(base[i][torch.from_numpy(perm[i]).to(dev)], srt[i]) for i in range(8))
    print("r=%.1f nsample=%d: caller's order %.3f ms, cell-sorted queries %.3f ms, same rows=%s" % (
        r, ns, t(lambda: P.query_ball_point(r, ns, ref, qry)), t(lambda: P.query_ball_point(r, ns, ref, qs)), same), flush=True)
