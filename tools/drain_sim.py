"""Host simulation of the kNN drain's lock-step cost as a function of how many tiles one drain covers.
Counts, for 32-lane warps of consecutive queries on the C2 clouds: phase-1 iterations (max over lanes of the hit
chunks of a drain), phase-2 iterations (max over lanes of the exact candidates of a pass) and heap inserts, with the
threshold of a query only tightening at drains.  usage: python tools/drain_sim.py [n_warps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np
from b200pc import synth

K, TILE, CHUNK, CAND_CAP = 16, 512, 8, 16
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 24
a, b = synth.batch_pairs(0, 1, 16384)
ref, qry = a[0].astype(np.float64), b[0].astype(np.float64)
N = ref.shape[0]; NT = N // TILE
# tile t holds refs t, t+NT, ...: slot s of tile t is ref t + s*NT
order = np.arange(N).reshape(TILE, NT).T.reshape(-1)   # order[t*TILE + s] = ref index

def tau0_grid(q):
    """stand-in for grid_tau0: farthest corner of the smallest 3^3 box of 2^l * 1 m cells around q holding >= K refs"""
    for l in range(0, 8):
        h = 0.5 * (1 << l)
        c = np.floor(q / h)
        lo, hi = (c - 1) * h, (c + 2) * h
        if np.count_nonzero(np.all((ref >= lo) & (ref < hi), axis=1)) >= K:
            return float(np.sum(np.maximum(np.abs(q - lo), np.abs(hi - q)) ** 2))
    return np.inf

def run(M, warm=True):
    p1 = p2 = ins = sift = cands = hitc = 0
    for w in range(nw):
        qs = qry[w * 32:(w + 1) * 32]
        d = ((qs[:, None, :] - ref[None, order, :]) ** 2).sum(-1)       # [32][N] in visiting order
        tau = np.array([tau0_grid(q) if warm else np.inf for q in qs])
        heaps = [[] for _ in range(32)]
        for t0 in range(0, NT, M):
            seg = d[:, t0 * TILE:(t0 + M) * TILE]
            thr = tau.copy()
            hit = (seg < thr[:, None]).reshape(32, -1, CHUNK).any(-1)   # filter: chunk masks at the drain-start threshold
            lists = [list(np.nonzero(hit[l])[0]) for l in range(32)]
            hitc += sum(len(x) for x in lists)
            while any(lists):
                take = [x[:CAND_CAP] for x in lists]
                lists = [x[CAND_CAP:] for x in lists]
                p1 += max(len(x) for x in take)
                per = []
                for l in range(32):
                    n = 0
                    for c in take[l]:
                        dd = seg[l, c * CHUNK:(c + 1) * CHUNK]
                        for v in dd[dd < thr[l]]:
                            n += 1
                            if v < tau[l]:
                                ins += 1
                                h = heaps[l]; h.append(v); h.sort()
                                if len(h) > K: h.pop()
                                if len(h) == K: tau[l] = min(tau[l], h[-1])
                    per.append(n)
                cands += sum(per); p2 += max(per)
                thr = tau.copy()
    n = nw * 32
    return dict(M=M, hit_chunks_q=hitc / n, cands_q=cands / n, ins_q=ins / n, p1_warp=p1 / nw, p2_warp=p2 / nw)

for warm in (True,):
    for M in (1, 2, 4, 8):
        print(run(M, warm), flush=True)
