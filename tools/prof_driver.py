"""Tiny driver for ncu captures: runs ONE op of the hot path a few times on synthetic C2/C3 inputs.
usage: python tools/prof_driver.py {knn|ball|three_nn|knn_direct|nn1|fps|gather|interp|group|chamfer} [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np
import torch
from b200pc import ops, pointnet2_utils as P, synth

op = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
B = int(os.environ.get("PROF_B", "8"))
a, b = synth.batch_pairs(0, B, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
if op == "knn":
    fn = lambda: P.knn_point(16, ref, qry)
elif op == "ball":
    fn = lambda: P.query_ball_point(1.0, 32, ref, qry)
elif op == "three_nn":
    known = ref[:, ::4].contiguous()
    fn = lambda: P.three_nn_weights(ref, known)
elif op == "knn_direct":
    fn = lambda: ops.knn_search(ref, qry, 16, ops.FORM_DIRECT, want_dist=True)
elif op == "nn1":
    fn = lambda: ops.knn_search(ref, qry, 1, ops.FORM_DIRECT)
elif op == "fps":
    start = torch.zeros(B, dtype=torch.long, device=dev)
    fn = lambda: ops.fps(ref, 4096, start)
elif op == "gather":
    feats = torch.randn(B, 16384, 128, device=dev); idx = torch.randint(0, 16384, (B, 4096), device=dev)
    fn = lambda: P.index_points(feats, idx)
elif op == "interp":
    feats = torch.randn(B, 4096, 128, device=dev)
    known = ref[:, ::4].contiguous(); _, i3, w3 = P.three_nn_weights(ref, known)
    fn = lambda: P.three_interpolate(feats, i3, w3)
elif op == "group":
    feats = torch.randn(B, 16384, 64, device=dev); known = ref[:, ::4].contiguous()
    gidx = P.knn_point(16, ref, known)
    fn = lambda: P.group_points(ref, known, feats, gidx)
elif op == "chamfer":
    fn = lambda: ops.chamfer(ref[:, :8192].contiguous(), qry[:, :8192].contiguous())
else:
    raise SystemExit("unknown op " + op)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok", op)
