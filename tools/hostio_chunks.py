"""e2e knn_point_host (C2, pinned host buffers in and out) for different chunk counts. usage: python tools/hostio_chunks.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import hostio, synth
a, b = synth.batch_pairs(0, 8, 16384)
ref = torch.from_numpy(a).pin_memory(); qry = torch.from_numpy(b).pin_memory()
out = torch.empty(8, 16384, 16, dtype=torch.int64).pin_memory()
for chunks in (1, 2, 4, 8, 2, 4):
    for _ in range(3): hostio.knn_point_host(16, ref, qry, out=out, chunks=chunks)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); hostio.knn_point_host(16, ref, qry, out=out, chunks=chunks); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("chunks=%d  median %.3f ms  min %.3f ms  -> %.4f Gq/s" % (chunks, ts[len(ts)//2], ts[0], 8*16384/ts[len(ts)//2]/1e6), flush=True)
# raw copy speeds
d = torch.empty(8, 16384, 16, dtype=torch.int64, device="cuda")
for _ in range(3): out.copy_(d, non_blocking=True)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); out.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H 16.8 MB: %.3f ms = %.1f GB/s" % (e0.elapsed_time(e1), 16.777216 / e0.elapsed_time(e1)))
