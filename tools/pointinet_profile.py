"""Where does a PointINet frame go?  torch profiler over a few forwards (GPU kernel time + CPU side)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ROOT)
import bench
from b200pc import pointinet
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pointinet.PointINet().eval().to(dev)
ins = bench.pointinet_inputs(100, 16384, dev=dev)
for _ in range(3):
    with torch.no_grad(): net(*ins)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.no_grad(): net(*ins)
torch.cuda.synchronize()
print("wall ms/frame %.2f" % ((time.perf_counter() - t0) * 100))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        with torch.no_grad(): net(*ins)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
