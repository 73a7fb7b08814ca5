"""kernel-time breakdown of one CUDA-graph replay of the PointINet forward at batch 8"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from b200pc import pointinet
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
ins = [torch.cat(x, 0) for x in zip(*[bench.pointinet_inputs(100 + i, 16384, dev=dev)[:4] for i in range(B)])]
g = pointinet.GraphedPointINet(batch=B, npoints=16384, extra=1, t=0.5, device=dev)
g.capture(*ins)
for _ in range(3): g(*ins)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): g(*ins)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
