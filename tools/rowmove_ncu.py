"""driver for ncu captures of the row movers: python tools/rowmove_ncu.py <bulk 0|1> (0 / 1: register / asynchronous
group_points); one call of index_points, three_interpolate and group_points each at the C3 shapes (B=16)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["B200PC_BULK"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import numpy as np
import torch
from b200pc import ops, pointnet2_utils as P, synth

dev = torch.device("cuda:0")
B3, N = 16, 16384
a, b = synth.batch_pairs(0, 8, N)
xyz = torch.from_numpy(np.concatenate([a, b], 0)[:B3]).to(dev)
start = torch.arange(B3, device=dev, dtype=torch.long) * 7
fidx = ops.fps(xyz, 4096, start)
feats = torch.randn(B3, N, 128, device=dev)
known = P.index_points(xyz, fidx)
gidx = P.knn_point(16, xyz, known)
sfeat = torch.randn(B3, 4096, 128, device=dev)
_, i3, w3 = P.three_nn_weights(xyz, known)
gfeat = torch.randn(B3, N, 64, device=dev)
torch.cuda.synchronize()
for _ in range(2):
    P.index_points(feats, fidx)
    P.three_interpolate(sfeat, i3, w3)
    P.group_points(xyz, known, gfeat, gidx)
torch.cuda.synchronize()
print("ok")
