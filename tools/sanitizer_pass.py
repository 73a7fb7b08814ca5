"""small shapes through every kernel family, for `compute-sanitizer --tool memcheck`"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))
import torch
from b200pc import ops, pointnet2_utils as P, pytorch3d_shim as S3, synth
import b200pc.ops as _b200pc_ops; _b200pc_ops.TUNING_AUTORELOAD = True   # this probe flips B200PC_* knobs between calls (the library caches them)
dev = torch.device("cuda:0")
a, b = synth.batch_pairs(0, 2, 3000)
ref = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b[:, :700].copy()).to(dev)
P.knn_point(16, ref, qry); ops.knn_search(ref, qry, 5, 2, want_dist=True); P.query_ball_point(1.0, 32, ref, qry)
os.environ["B200PC_SMALL_PATH"] = "0"; P.knn_point(8, ref[:, :600].contiguous(), qry)
os.environ["B200PC_SMALL_PATH"] = "1"; P.knn_point(8, ref[:, :600].contiguous(), qry); P.query_ball_point(0.5, 8, ref[:, :300].contiguous(), qry[:, :50].contiguous())
os.environ["B200PC_FORCE_SPLIT"] = "3"; P.knn_point(4, ref, qry[:, :40].contiguous()); P.query_ball_point(2.0, 16, ref, qry[:, :40].contiguous()); os.environ.pop("B200PC_FORCE_SPLIT")
st = torch.tensor([1, 2], device=dev)
fi = ops.fps(ref, 64, st); ops.fps(torch.cat([ref] * 4, 1), 32, st)      # 12000 points -> cluster of 2
feats = torch.randn(2, 3000, 32, device=dev, requires_grad=True)
g = P.index_points(feats, fi); g.sum().backward()
P.index_points(ref, fi)
known = P.index_points(ref, fi); d, i3, w = P.three_nn_weights(ref, known)
sf = torch.randn(2, 64, 32, device=dev, requires_grad=True); ww = w.clone().requires_grad_(True)
P.three_interpolate(sf, i3, ww).sum().backward()
P.square_distance(ref[:, :100], qry[:, :33])
x = ref[:, :500].clone().requires_grad_(True); S3.chamfer_distance(x, qry)[0].backward()
torch.cuda.synchronize(); print("sanitizer pass ok")
