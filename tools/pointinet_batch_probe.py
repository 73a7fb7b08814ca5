"""PointINet CUDA-graph throughput for several frame pairs per GPU (batch 1, 2, 4, 8). usage: python tools/pointinet_batch_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
import bench
from b200pc import pointinet
dev = torch.device("cuda:0")
torch.manual_seed(0)
state = pointinet.PointINet().eval().state_dict()
for B in (1, 2, 4, 8):
    ins = [torch.cat(x, 0) for x in zip(*[bench.pointinet_inputs(100 + i, 16384, dev=dev)[:4] for i in range(B)])]
    g = pointinet.GraphedPointINet(state_dict=state, batch=B, npoints=16384, extra=1, t=0.5, device=dev)
    g.capture(*ins)
    for _ in range(3): g(*ins)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g(*ins); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print("batch %d: %.3f ms per replay = %.3f ms per frame = %.1f frames/s" % (B, ms, ms / B, B * 1e3 / ms), flush=True)
    del g
