"""A/B of two builds of libb200pc.so on the SAME box: each library runs the same search shapes in a child process.
usage: python tools/lib_ab.py path/to/libA.so path/to/libB.so   (build variants into tools/probe/, they are git-ignored)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
SHAPES = [  # (kind, B, N refs, S queries, k / nsample)
    ("knn0", 8, 16384, 16384, 16), ("knn0", 8, 16384, 16384, 64), ("knn2", 1, 16384, 16384, 16), ("knn0", 2, 16384, 16384, 16),
    ("knn0", 32, 8192, 8192, 16), ("knn0", 1, 4096, 1024, 64), ("knn0", 1, 4096, 300, 128), ("knn0", 1, 64000, 4096, 16),
    ("knn2", 8, 16384, 16384, 1), ("knn2", 1, 65536, 65536, 1), ("knn2", 4, 8192, 8192, 1), ("knn1", 16, 4096, 16384, 3),
    ("ball", 8, 16384, 16384, 32), ("ball", 1, 16384, 1024, 16), ("ball", 1, 64000, 1024, 32), ("ball", 4, 8192, 2048, 64),
    ("knn2", 1, 65536, 32768, 1), ("knn2", 4, 65536, 8192, 1), ("knn2", 4, 65536, 65536, 1), ("knn0", 1, 16384, 4096, 16), ("knn0", 1, 4096, 4096, 32),
    ("knn1", 1, 1024, 16384, 3), ("knn1", 1, 4096, 16384, 3), ("ball", 1, 4096, 1024, 32), ("knn0", 4, 2048, 2048, 16), ("knn2", 2, 8192, 8192, 16),
]
if len(sys.argv) > 2 and sys.argv[1] == "child":
    import numpy as np, torch
    from b200pc import _lib
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
    from b200pc import ops, pointnet2_utils as P, synth
    dev = torch.device("cuda:0"); out = {}
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    fr = [synth.frame_pair(500 + i, 65536) for i in range(2)]
    def t(fn, n=8):
        for _ in range(3): fn()
        torch.cuda.synchronize(); tot = 0.0
        for _ in range(n):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / n
    for kind, B, N, S, k in SHAPES:
        ref = torch.from_numpy(np.stack([fr[i % 2][0][:N] if N <= 65536 else None for i in range(B)])).to(dev)
        qry = torch.from_numpy(np.stack([fr[i % 2][1][:S] for i in range(B)])).to(dev)
        if kind == "ball": fn = lambda: P.query_ball_point(1.0, k, ref, qry)
        else: fn = lambda: ops.knn_search(ref, qry, k, int(kind[3]))
        out["%s B%d N%d S%d k%d" % (kind, B, N, S, k)] = t(fn)
    print("RESULT " + json.dumps(out))
    sys.exit(0)
res = []
for lib in sys.argv[1:]:
    o = subprocess.run([sys.executable, __file__, "child", lib], capture_output=True, text=True)
    line = [l for l in o.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        print(o.stdout[-2000:], o.stderr[-2000:]); sys.exit(1)
    res.append(json.loads(line[0][7:]))
print("%-34s " % "shape" + " ".join("%12s" % os.path.basename(a)[:12] for a in sys.argv[1:]) + "   (ratio to the first)")
for kname in res[0]:
    print("%-34s " % kname + " ".join("%12.3f" % r[kname] for r in res) + "   " + " ".join("%.3f" % (r[kname] / res[0][kname]) for r in res[1:]))
