"""Timeline of ONE CUDA-graph replay of the PointINet forward (C1: 16384 points, batch 1): every kernel with its start,
duration and stream, grouped by op, with the gaps where nothing runs -- written as text to gpurun_out/ (copied to
profiles/r02_pointinet_graph.txt).  python tools/pointinet_trace.py [out.txt]"""
import json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from b200pc import pointinet
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pointinet_graph_timeline.txt")
dev = torch.device("cuda:0")
torch.manual_seed(0)
ins = bench.pointinet_inputs(100, 16384, dev=dev)
g = pointinet.GraphedPointINet(batch=1, npoints=16384, extra=1, t=0.5, device=dev)
g.capture(*ins[:4])
for _ in range(5): g(*ins[:4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(*ins[:4]); torch.cuda.synchronize()
tmp = out_path + ".json"
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
os.remove(tmp)
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)


def family(n):
    for key, fam in (("search_kernel", "search (kNN / ball / three-NN)"), ("small_search", "search, small path"), ("fps_kernel", "FPS"),
                     ("pack_refs", "search: pack refs"), ("grid_", "search: occupancy grid"), ("merge_", "search: merge of ref splits"),
                     ("group_points", "fused grouping"), ("fusion_features", "fusion features"), ("gather_", "index_points"),
                     ("interp_", "three_interpolate"), ("three_weights", "three-NN weights"), ("gemm", "conv / linear (library)"),
                     ("cutlass", "conv / linear (library)"), ("conv", "conv / linear (library)"), ("sgemm", "conv / linear (library)"),
                     ("elementwise", "torch elementwise / copies"), ("reduce", "torch reductions (max / softmax / sum)"), ("softmax", "torch reductions (max / softmax / sum)"),
                     ("Memcpy", "memcpy"), ("Memset", "memset"), ("cat", "torch cat / copies"), ("copy", "torch elementwise / copies")):
        if key.lower() in n.lower():
            return fam
    return "other torch kernels"


fam = collections.defaultdict(lambda: [0.0, 0])
for e in ev:
    f = family(e["name"]); fam[f][0] += e["dur"]; fam[f][1] += 1
# union of busy intervals -> wall time with at least one kernel running; idle gaps
busy = 0.0; cur_s, cur_e = ev[0]["ts"], ev[0]["ts"] + ev[0]["dur"]
gaps = []
for e in ev[1:]:
    s, d = e["ts"], e["dur"]
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((cur_e - t0, s - cur_e)); cur_s, cur_e = s, s + d
    else:
        cur_e = max(cur_e, s + d)
busy += cur_e - cur_s
with open(out_path, "w") as fh:
    w = lambda *a: print(*a, file=fh)
    w("PointINet forward, one CUDA-graph replay, C1 (16384 points, batch 1, t=0.5): %d kernels / copies, wall %.1f us, "
      "GPU busy (>= 1 kernel running) %.1f us, idle inside the replay %.1f us" % (len(ev), t1 - t0, busy, (t1 - t0) - busy))
    w("\nkernel time by family (sum over parallel branches; the replay runs up to 4 branches side by side):")
    for f, (d, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        w("  %-44s %9.1f us  %4d launches" % (f, d, n))
    w("\nlongest kernels (start relative to the replay, stream):")
    for e in sorted(ev, key=lambda e: -e["dur"])[:24]:
        w("  %8.1f us  +%7.1f  stream %-4s %s" % (e["dur"], e["ts"] - t0, e.get("args", {}).get("stream", "?"), e["name"][:110]))
    w("\nidle gaps > 3 us (no kernel running):")
    for s, d in gaps:
        if d > 3.0:
            w("  at +%.1f us: %.1f us" % (s, d))
    w("\ncritical path: the kernels that end last in each 250 us window")
    win = 250.0
    k = 0
    while k * win < t1 - t0:
        inw = [e for e in ev if k * win <= e["ts"] + e["dur"] - t0 < (k + 1) * win]
        if inw:
            e = max(inw, key=lambda e: e["dur"])
            w("  [%5.0f, %5.0f) us: %3d kernels end here; longest %7.1f us %s" % (k * win, (k + 1) * win, len(inw), e["dur"], e["name"][:90]))
        k += 1
print(open(out_path).read())
