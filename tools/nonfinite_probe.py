"""crash / hang safety on non-finite and degenerate coordinates (no parity claim for NaN rows; the reference's own
result for them is unspecified).  usage: timeout 120 python tools/nonfinite_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import numpy as np, torch
from b200pc import ops, pointnet2_utils as P, synth
from oracle import strict
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(3, 1, 6000)
ref, qry = a.copy(), b[:, :1500].copy()
ref[0, 17] = np.nan; ref[0, 4000, 1] = np.inf; qry[0, 5] = np.nan; qry[0, 9, 0] = -np.inf
clean_ref = ref.copy(); clean_ref[0, 17] = 1e6; clean_ref[0, 4000] = 1e6      # the oracle is not NaN-safe: compare with far-away finite points
clean = np.ones(1500, bool); clean[[5, 9]] = False
for form in (0, 1, 2):
    idx, dist = ops.knn_search(torch.from_numpy(ref).to(dev), torch.from_numpy(qry).to(dev), 16, form, want_dist=True)
    torch.cuda.synchronize()
    qc = qry.copy(); qc[0, [5, 9]] = 0
    oi, od = strict.knn(clean_ref, qc, 16, form)
    ok = np.array_equal(idx.cpu().numpy()[0][clean], oi[0][clean])
    print("form %d: finite-query rows equal to the oracle on the cleaned cloud: %s; NaN-query row idx[:3] = %s" % (form, ok, idx[0, 5, :3].tolist()), flush=True)
ball = P.query_ball_point(1.0, 16, torch.from_numpy(ref).to(dev), torch.from_numpy(qry).to(dev)); torch.cuda.synchronize()
print("ball rows equal:", np.array_equal(ball.cpu().numpy()[0][clean], strict.query_ball_point(1.0, 16, clean_ref, qc)[0][clean]), flush=True)
same = np.tile(np.float32([[1.5, -2.0, 0.25]]), (1, 5000, 1))
i2 = P.knn_point(16, torch.from_numpy(same).to(dev), torch.from_numpy(same[:, :300]).to(dev)); torch.cuda.synchronize()
print("all-identical cloud -> lowest indices:", bool((i2.cpu() == torch.arange(16)).all()))
f = ops.fps(torch.from_numpy(same).to(dev), 8, torch.zeros(1, dtype=torch.long, device=dev)); torch.cuda.synchronize()
print("fps on identical points:", f.tolist(), "oracle", strict.farthest_point_sample(same, 8, np.zeros(1, np.int64)).tolist())
print("done")
