"""search-kernel timings of this build against the v6 library (refs broadcast) on a spread of shapes.
usage: python tools/regress_vs_v6.py            (needs tools/probe/old_v6_libb200pc.so, built from commit 95d229d)"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
SHAPES = [  # (kind, B, N refs, S queries, k / nsample, extra)
    ("knn0", 8, 16384, 16384, 16), ("knn0", 1, 16384, 16384, 16), ("knn0", 2, 4096, 4096, 16), ("knn0", 1, 4096, 1024, 64),
    ("knn0", 1, 4096, 300, 128), ("knn0", 4, 8192, 8192, 8), ("knn0", 1, 2048, 2048, 32), ("knn0", 1, 64000, 4096, 16),
    ("knn2", 8, 16384, 16384, 16), ("knn2", 1, 65536, 65536, 1), ("knn2", 4, 8192, 8192, 1), ("knn1", 16, 4096, 16384, 3),
    ("ball", 8, 16384, 16384, 32), ("ball", 1, 16384, 1024, 16), ("ball", 1, 64000, 1024, 32), ("ball", 4, 8192, 2048, 64),
]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from b200pc import _lib, ops, pointnet2_utils as P, synth
    if sys.argv[2] != "new":
        _lib.LIB_PATH = sys.argv[2]
        for name in ("b200pc_group_points", "b200pc_group_points_bwd", "b200pc_poly_predict"): _lib.SIGNATURES.pop(name, None)   # newer than v6
    dev = torch.device("cuda:0"); out = {}
    a, b = synth.batch_pairs(0, 16, 16384)
    big = torch.from_numpy(a.reshape(1, -1, 3)).to(dev)
    def t(fn, n=8):
        for _ in range(3): fn()
        torch.cuda.synchronize(); tot = 0
        for _ in range(n):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / n
    for kind, B, N, S, k in SHAPES:
        if N > 16384 or S > 16384:
            ref = big[:, :N].contiguous(); qry = big[:, 70000:70000 + S].contiguous()
        else:
            ref = torch.from_numpy(a[:B, :N]).to(dev).contiguous(); qry = torch.from_numpy(b[:B, :S]).to(dev).contiguous()
        if kind == "ball": fn = lambda: P.query_ball_point(1.0, k, ref, qry)
        else: fn = lambda: ops.knn_search(ref, qry, k, int(kind[3]))
        out["%s B%d N%d S%d k%d" % (kind, B, N, S, k)] = t(fn)
    print(json.dumps(out))
else:
    res = {}
    for tag, lib in (("v6", os.path.join(ROOT, "tools", "probe", "old_v6_libb200pc.so")), ("now", "new")):
        p = subprocess.run([sys.executable, __file__, "child", lib], capture_output=True, text=True)
        if p.returncode: print(p.stderr[-2000:]); sys.exit(1)
        res[tag] = json.loads(p.stdout.strip().splitlines()[-1])
    for key in res["v6"]:
        print("%-34s v6 %8.3f ms   now %8.3f ms   x%.2f" % (key, res["v6"][key], res["now"][key], res["v6"][key] / res["now"][key]))
