"""launch list driver: C2 knn_point with the grid warm start on (arg 1) or off (arg 0)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
os.environ["B200PC_GRID"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import torch
from b200pc import pointnet2_utils as P, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b).to(dev)
for _ in range(3):
    P.knn_point(16, ref, qry)
torch.cuda.synchronize()
print("ok")
