"""do four independent FPS launches (cluster kernels) on four streams overlap inside a CUDA graph?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, synth
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 2, 16384)
x = [torch.from_numpy(a[:1]).to(dev), torch.from_numpy(b[:1]).to(dev), torch.from_numpy(a[1:]).to(dev), torch.from_numpy(b[1:]).to(dev)]
st = [torch.tensor([i * 7], device=dev) for i in range(4)]
streams = [torch.cuda.Stream() for _ in range(3)]
def four_streams():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    outs = [ops.fps(x[0], 1024, st[0])]
    for i, s in enumerate(streams):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            outs.append(ops.fps(x[i + 1], 1024, st[i + 1]))
    for s in streams:
        cur.wait_stream(s)
    return outs
def one():
    return [ops.fps(x[0], 1024, st[0])]
xb = torch.cat(x, 0); sb = torch.cat(st, 0)
def batched():
    return [ops.fps(xb, 1024, sb)]
def timeit(fn, graph):
    fn(); torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            with torch.cuda.graph(g, stream=s):
                keep = fn()
        run = g.replay
    else:
        run = fn
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for name, fn in (("one FPS 16384->1024", one), ("four on four streams", four_streams), ("four as one batched launch", batched)):
    print("%-28s eager %.3f ms   graph %.3f ms" % (name, timeit(fn, False), timeit(fn, True)), flush=True)
