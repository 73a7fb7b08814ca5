"""host-side cost of one op call (tiny problem, GPU time negligible): facade -> torch.ops dispatcher -> ctypes -> launch"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P
dev = torch.device("cuda:0")
ref = torch.randn(1, 256, 3, device=dev); qry = torch.randn(1, 64, 3, device=dev); feat = torch.randn(1, 256, 32, device=dev)
idx = P.knn_point(8, ref, qry)
def t(name, fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print("%-34s %.1f us per call" % (name, (time.perf_counter() - t0) / n * 1e6), flush=True)
with torch.no_grad():
    t("torch.add (baseline)", lambda: ref + 1.0)
    t("knn_point", lambda: P.knn_point(8, ref, qry))
    t("query_ball_point", lambda: P.query_ball_point(0.5, 8, ref, qry))
    t("index_points", lambda: P.index_points(feat, idx))
    t("group_points", lambda: P.group_points(ref, qry, feat, idx))
    t("fps", lambda: ops.fps(ref, 16, torch.zeros(1, dtype=torch.long, device=dev)))
    t("feature_propagation", lambda: P.feature_propagation(ref, qry, torch.randn(1, 64, 8, device=dev)))
    t("torch.ops.b200pc.knn direct", lambda: torch.ops.b200pc.knn(ref, qry, 8, 0, False))
    t("ops._knn impl direct", lambda: ops._knn(ref, qry, 8, 0, False))
