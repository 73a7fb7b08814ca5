"""Parity at the largest sizes SURVEY section 8a lists (ISAPCInet: 2*field*N = 64 000 points; C5: 65 536 points),
sampled where the oracle would take too long, plus the host-pointer C-ABI wrappers."""
import ctypes as C

import numpy as np
import pytest
import torch

from b200pc import _lib, ops, pointnet2_utils as P, pytorch3d_shim as S3, synth
from oracle import strict

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.fixture(scope="module")
def cloud64k():
    a, b = synth.batch_pairs(300, 4, 16384)
    return np.ascontiguousarray(np.concatenate([a[i] for i in range(4)], 0)[None][:, :64000])   # [1,64000,3]


def test_fps_64000_to_1024_cluster_of_8(cuda_dev, cloud64k):
    start = np.array([12345])
    out = P.farthest_point_sample_from(_t(cloud64k, cuda_dev), 1024, _t(start, cuda_dev))
    np.testing.assert_array_equal(out.cpu().numpy(), strict.farthest_point_sample(cloud64k, 1024, start))


def test_sa_msg_ball_queries_on_64000_points(cuda_dev, cloud64k):
    # Pointnet2FeatureAbstract.sa1: 1 024 FPS centres, radii 0.1 / 0.2, nsample 16 / 32 (Utils/Layers.py:502)
    centres = cloud64k[:, ::62][:, :1024].copy()
    for r, ns in ((0.1, 16), (0.2, 32)):
        out = P.query_ball_point(r, ns, _t(cloud64k, cuda_dev), _t(centres, cuda_dev))
        np.testing.assert_array_equal(out.cpu().numpy(), strict.query_ball_point(r, ns, cloud64k, centres))


def test_three_nn_64000_from_1024_variant_b(cuda_dev, cloud64k):
    known = cloud64k[:, ::62][:, :1024].copy()
    dist, idx, w = P.three_nn_weights(_t(cloud64k, cuda_dev), _t(known, cuda_dev), variant=1)
    od, oi = strict.three_nn(cloud64k, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(dist.cpu().numpy().view(np.int32), od.view(np.int32))
    np.testing.assert_allclose(w.cpu().numpy(), strict.three_weights(od, 1), rtol=1e-5, atol=1e-12)


def test_transformer_self_knn_64000_sampled(cuda_dev, cloud64k):
    # TransformerLayer: knn_points(xyz, xyz, K=16) over 64 000 points (Utils/Layers.py:430); sampled check
    x = _t(cloud64k, cuda_dev)
    r = S3.knn_points(x, x, K=16, return_nn=True)
    sel = np.arange(0, 64000, 331)
    od, oi = strict.knn_points(cloud64k[:, sel], cloud64k, 16)
    np.testing.assert_array_equal(r.idx.cpu().numpy()[:, sel], oi)
    np.testing.assert_array_equal(r.dists.cpu().numpy()[:, sel].view(np.int32), od.view(np.int32))
    assert (r.idx[:, :, 0].cpu().numpy()[0] == np.arange(64000)).mean() > 0.99      # self is the nearest (duplicates aside)
    assert torch.equal(r.knn, S3.knn_gather(x, r.idx))


def test_c5_rebuild_k1_65536_sampled(cuda_dev):
    # PolyPCI.rebuild: knn_points(K=1, return_nn=True) on a full 65 536-point sweep (Models_V1.py:102-114)
    a, b = synth.batch_pairs(310, 4, 16384)
    ref = np.ascontiguousarray(np.concatenate(list(a), 0)[None]); qry = np.ascontiguousarray(np.concatenate(list(b), 0)[None])
    r = S3.knn_points(_t(qry, cuda_dev), _t(ref, cuda_dev), K=1, return_nn=True)
    sel = np.arange(0, 65536, 257)
    od, oi = strict.knn_points(qry[:, sel], ref, 1)
    np.testing.assert_array_equal(r.idx.cpu().numpy()[:, sel], oi)
    np.testing.assert_array_equal(r.knn.cpu().numpy()[:, sel, 0], ref[0][oi[0, :, 0]][None])


def test_host_pointer_abi_wrappers(cuda_dev):
    lib = _lib.load()
    a, b = synth.batch_pairs(320, 2, 1500)
    qry = np.ascontiguousarray(b[:, :300])
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    idx = np.empty((2, 300, 8), np.int64); dist = np.empty((2, 300, 8), np.float32)
    _lib.check(lib.b200pc_knn_host(p(a), p(qry), 2, 1500, 300, 8, 0, p(idx), p(dist)))
    oi, od = strict.knn(a, qry, 8, 0)
    np.testing.assert_array_equal(idx, oi); np.testing.assert_array_equal(dist.view(np.int32), od.view(np.int32))
    ball = np.empty((2, 300, 16), np.int64)
    _lib.check(lib.b200pc_ball_query_host(p(a), p(qry), 2, 1500, 300, C.c_float(float(strict.radius_sq(0.7))), 16, p(ball)))
    np.testing.assert_array_equal(ball, strict.query_ball_point(0.7, 16, a, qry))
    start = np.array([3, 1400], np.int64); fps = np.empty((2, 100), np.int64)
    _lib.check(lib.b200pc_fps_host(p(a), 2, 1500, 100, p(start), p(fps)))
    np.testing.assert_array_equal(fps, strict.farthest_point_sample(a, 100, start))


def test_hostio_pipelined_knn_equals_plain(cuda_dev):
    from b200pc import hostio
    a, b = synth.batch_pairs(330, 5, 3000)
    h_ref = torch.from_numpy(a).pin_memory(); h_qry = torch.from_numpy(b[:, :1111].copy()).pin_memory()
    for chunks in ("auto", 1, 2, 3):
        out = hostio.knn_point_host(16, h_ref, h_qry, device=cuda_dev, chunks=chunks)
        ball = hostio.query_ball_point_host(1.0, 8, h_ref, h_qry, device=cuda_dev, chunks=chunks)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.numpy(), strict.knn_point(16, a, b[:, :1111]))
        np.testing.assert_array_equal(ball.numpy(), strict.query_ball_point(1.0, 8, a, b[:, :1111]))
