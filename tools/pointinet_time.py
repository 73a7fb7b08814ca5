"""PointINet (C1: 16384 points, batch 1) as one CUDA graph: ms per frame over 50 replays"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch, bench
from b200pc import pointinet
dev = torch.device("cuda:0"); torch.manual_seed(0)
ins = bench.pointinet_inputs(100, 16384, dev=dev)
g = pointinet.GraphedPointINet(batch=1, npoints=16384, extra=1, t=0.5, device=dev); g.capture(*ins[:4])
for _ in range(5): g(*ins[:4])
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): g(*ins[:4])
e1.record(); torch.cuda.synchronize()
print("graph ms/frame %.3f  (%.1f frames/s)" % (e0.elapsed_time(e1) / 50, 50e3 / e0.elapsed_time(e1)))
