"""Generate tests/golden/*.npz by EXECUTING THE REAL REFERENCE on CPU (build container only).

Run:  python tests/golden/make_golden.py          (needs /root/reference; ~1 minute)

Every input is regenerated from seeds by b200pc.synth, so the fixtures hold outputs only (plus a
few small inputs for self-containment).  The reference has no tests of its own, so
these vectors are the pin for oracle/strict.c and oracle/ref_torch.py:

  sqdist      Utils.Pointnet2Utils.square_distance                         (file :20)
  fps         Utils.Pointnet2Utils.farthest_point_sample under manual_seed (file :64)
  ball        Utils.Pointnet2Utils.query_ball_point                        (file :88)
  gather      Utils.Pointnet2Utils.index_points                            (file :44)
  knn         Utils.Layers.Group(knn=True).forward, indices recovered by passing the point index
              as a feature channel                                         (Layers.py:42-66)
  group       Utils.Layers.Group.forward, whole output tensor [B,3+D,ns,S] as a digest      (Layers.py:42-66)
  fp_a        Utils.Layers.FeaturePropagation.forward with its conv stack replaced by Identity on the
              instance, identity-matrix features -> per-point weight rows  (Layers.py:174-192)
  fp_b        Utils.Pointnet2Utils.PointNetFeaturePropagation(mlp=[]).forward   (file :279-313)
  polyfit     PolyPCI.Models.Models_V1.PolyPCI.fitting_and_predict (function source executed) (:116-124)

Outputs that are large are stored as int32 / as SHA-256 digests of their bytes.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-interpolation-_b200"))

from b200pc import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    assert ref_loader.available(), "needs the reference checkout"
    torch.set_num_threads(8)
    R = ref_loader.pointnet2_utils()
    L = ref_loader.layers()
    out = {}

    # ---- square_distance -------------------------------------------------------------------
    a, b = synth.batch_pairs(100, 2, 4096)
    A, Bt = torch.from_numpy(a), torch.from_numpy(b)
    d_small = R.square_distance(A[:, :96], Bt[:, :64]).numpy()
    out["sqdist_small"] = d_small                                        # pair 100, src[:96], dst[:64]
    out["sqdist_4096x1024_sha"] = digest(R.square_distance(A, Bt[:, :1024]).numpy())
    out["sqdist_1024x4096_sha"] = digest(R.square_distance(A[:, :1024], Bt).numpy())
    # non-contiguous (permuted) input, as SA-MSG passes it (Pointnet2Utils.py:234)
    out["sqdist_permuted_sha"] = digest(R.square_distance(A.permute(0, 2, 1).contiguous().permute(0, 2, 1), Bt[:, :256]).numpy())

    # ---- farthest_point_sample -------------------------------------------------------------
    torch.manual_seed(3000)
    out["fps_4096_512"] = R.farthest_point_sample(A, 512).numpy().astype(np.int32)
    a16, _ = synth.batch_pairs(101, 1, 16384)
    torch.manual_seed(3001)
    out["fps_16384_1024"] = R.farthest_point_sample(torch.from_numpy(a16), 1024).numpy().astype(np.int32)
    dup = np.concatenate([a[:, :1500], a[:, :548]], 1)                   # padded-duplicate cloud
    torch.manual_seed(3002)
    out["fps_dup_2048_700"] = R.farthest_point_sample(torch.from_numpy(dup), 700).numpy().astype(np.int32)

    # ---- query_ball_point ------------------------------------------------------------------
    for r, ns, nq in ((1.0, 32, 512), (0.5, 16, 512), (0.1, 16, 256), (4.0, 8, 64)):
        g = R.query_ball_point(r, ns, A, Bt[:, :nq]).numpy().astype(np.int32)
        out["ball_r%g_ns%d_q%d" % (r, ns, nq)] = g
    # queries taken from the refs themselves (SetConv): self distance ~0, possibly negative
    out["ball_self_r0.5_ns16"] = R.query_ball_point(0.5, 16, A, A[:, ::8].contiguous()).numpy().astype(np.int32)

    # ---- index_points ----------------------------------------------------------------------
    rng = np.random.default_rng(5)
    feats = rng.normal(size=(2, 4096, 24)).astype(np.float32)
    idx3 = rng.integers(-4096, 4096, size=(2, 50, 4))                    # negatives wrap
    out["gather_idx"] = idx3.astype(np.int32)
    out["gather_out_sha"] = digest(R.index_points(torch.from_numpy(feats), torch.from_numpy(idx3)).numpy())

    # ---- kNN through the real Group.forward ------------------------------------------------
    for k, nq, nr in ((16, 512, 4096), (8, 256, 64), (64, 256, 256)):
        pts = A[:, :nr].permute(0, 2, 1).contiguous()                    # [B,3,N]
        new = Bt[:, :nq].permute(0, 2, 1).contiguous()                   # [B,3,S]
        fidx = torch.arange(nr, dtype=torch.float32).view(1, 1, nr).repeat(2, 1, 1)
        grp = L.Group(None, k, knn=True)
        o = grp(pts, new, fidx)                                          # [B,3+1,k,S]
        out["knn_k%d_q%d_r%d" % (k, nq, nr)] = o[:, 3].permute(0, 2, 1).round().numpy().astype(np.int32)

    # ---- the full Group.forward output (gather, centre, cat, permute) with real feature channels ----
    featg = torch.from_numpy(np.random.default_rng(6).normal(size=(2, 5, 4096)).astype(np.float32))     # [B,D=5,N]
    new512 = Bt[:, :512].permute(0, 2, 1).contiguous()
    ptsg = A.permute(0, 2, 1).contiguous()
    out["group_feat_seed6"] = np.array([6, 5], np.int32)                 # rng seed, D
    # kNN variant: topk's order among exactly tied distances is unspecified, so the tensor itself is stored (48 centres)
    out["group_knn16_q48"] = L.Group(None, 16, knn=True)(ptsg, new512[:, :, :48].contiguous(), featg).numpy()   # [2,8,16,48]
    self512 = A[:, ::8].permute(0, 2, 1).contiguous()                    # centres taken from the refs (SetConv): no empty ball
    out["group_ball_self_r1_ns32_sha"] = digest(L.Group(1.0, 32, knn=False)(ptsg, self512, featg).numpy())   # [2,8,32,512]

    # ---- three-NN + interpolation, variant A: FeaturePropagation ---------------------------
    S, N = 64, 1024
    sparse = A[:, :N:N // S][:, :S].contiguous()                         # 64 sparse points
    dense = A[:, :N].contiguous()
    fp = L.FeaturePropagation(S, 0, [8])
    fp.conv = torch.nn.Identity()
    eye = torch.eye(S).unsqueeze(0).repeat(2, 1, 1)                      # features1 [B,D1=S,S]
    f2 = torch.zeros(2, 0, N)
    w_rows = fp(sparse.permute(0, 2, 1).contiguous(), dense.permute(0, 2, 1).contiguous(), eye, f2)
    out["fp_a_weight_rows"] = w_rows.permute(0, 2, 1).contiguous().numpy()   # [B,N,S]: 3 non-zeros per row
    featC = torch.from_numpy(rng.normal(size=(2, 32, S)).astype(np.float32))
    out["fp_a_feat"] = featC.numpy()
    out["fp_a_out"] = fp(sparse.permute(0, 2, 1).contiguous(), dense.permute(0, 2, 1).contiguous(), featC, f2).numpy()

    # ---- variant B: PointNetFeaturePropagation with an empty MLP ---------------------------
    pfp = R.PointNetFeaturePropagation(32, [])
    out["fp_b_out"] = pfp(dense.permute(0, 2, 1), sparse.permute(0, 2, 1), None, featC).numpy()

    # ---- PolyPCI polynomial fit: the REAL fitting_and_predict, extracted from the file's AST (the module itself
    # imports datasets / visualisers that are not installed) and executed on seeded frames ---------------------------
    import ast
    import types
    from sklearn.preprocessing import PolynomialFeatures
    src = open(os.path.join(ref_loader.REF_ROOT, "PolyPCI", "Models", "Models_V1.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "fitting_and_predict")
    ns = {"np": np, "PolynomialFeatures": PolynomialFeatures}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "Models_V1.py", "exec"), ns)
    prng = np.random.default_rng(9)
    for tag, T, tq, deg in (("f5_d3", [0.0, -1.0, 1.0, -2.0, 2.0], 0.5, 3), ("f7_d2", [0.0, -1.0, 1.0, -2.0, 2.0, -3.0, 3.0], -0.25, 2)):
        frames = (prng.normal(size=(len(T), 96)) * 30).astype(np.float32)              # [F,N], seed 9, drawn in this order
        val = ns["fitting_and_predict"](types.SimpleNamespace(degree=deg), np.array(T).reshape(-1), frames, torch.tensor(tq))
        out["polyfit_%s" % tag] = torch.tensor(val).to(torch.float32).numpy()           # like Models_V1.py:198
        out["polyfit_%s_T" % tag] = np.array(T, np.float64); out["polyfit_%s_t" % tag] = np.array([tq, deg], np.float64)

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    for k_, v in out.items():
        print("  %-28s %s %s" % (k_, v.dtype, v.shape))


if __name__ == "__main__":
    main()
