"""kernel-time breakdown of one CUDA-graph replay of the PointINet forward"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from b200pc import pointinet
dev = torch.device("cuda:0")
torch.manual_seed(0)
ins = bench.pointinet_inputs(100, 16384, dev=dev)
g = pointinet.GraphedPointINet(batch=1, npoints=16384, extra=1, t=0.5, device=dev)
g.capture(*ins[:4])
for _ in range(3): g(*ins[:4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): g(*ins[:4])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
