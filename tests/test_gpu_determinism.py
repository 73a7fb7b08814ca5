"""Run-to-run determinism as a race smoke test (racecheck is not available on the GPU pool): every op is a pure function of
its inputs -- the cell order of the gridded searches depends on atomics, the results must not -- so 25 repetitions, interleaved
with other kernels on a second stream, have to reproduce the first result bit for bit."""
import numpy as np
import pytest
import torch

from b200pc import ops, pointnet2_utils as P, synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    if isinstance(a, (tuple, list)):
        return all(_same(x, y) for x, y in zip(a, b))
    if a is None:
        return b is None
    return torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)


@pytest.mark.parametrize("grid", ["0", "2", "3"])
def test_searches_reproduce_bit_for_bit(cuda_dev, monkeypatch, grid):
    monkeypatch.setenv("B200PC_GRID", grid); monkeypatch.setenv("B200PC_SMALL_PATH", "0"); ops.reload_tuning()
    a, b = synth.batch_pairs(31, 2, 16384)
    ref = torch.from_numpy(a).to(cuda_dev); qry = torch.from_numpy(b[:, :6000].copy()).to(cuda_dev)
    noise = torch.empty(1 << 24, device=cuda_dev)
    side = torch.cuda.Stream(device=cuda_dev)
    fns = [lambda: ops.knn_search(ref, qry, 16, 0, want_dist=True), lambda: ops.knn_search(ref, qry[:, :1000].contiguous(), 8, 2, want_dist=True),
           lambda: P.query_ball_point(1.0, 32, ref, qry), lambda: P.three_nn(qry, ref[:, :4096].contiguous())]
    for fn in fns:
        first = fn()
        for _ in range(25):
            with torch.cuda.stream(side):
                noise.normal_()                      # something else keeps the SMs and the L2 busy
            assert _same(first, fn())
    torch.cuda.synchronize()


def test_fps_grouping_and_fusion_reproduce_bit_for_bit(cuda_dev, monkeypatch):
    a, b = synth.batch_pairs(32, 2, 16384)
    xyz = torch.from_numpy(a).to(cuda_dev)
    start = torch.tensor([3, 11], device=cuda_dev)
    feat = torch.randn(2, 16384, 64, device=cuda_dev)
    fi = ops.fps(xyz, 2048, start)
    new_xyz = P.index_points(xyz, fi)
    gi = P.knn_point(16, xyz, new_xyz)
    fns = [lambda: ops.fps(xyz, 2048, start), lambda: P.group_points(xyz, new_xyz, feat, gi, xyz_first=True),
           lambda: P.fusion_group(new_xyz, xyz, 16, feat[:, :, :13].contiguous()), lambda: P.feature_propagation(xyz, new_xyz, feat[:, :2048].contiguous())]
    for bulk in ("0", "1"):
        monkeypatch.setenv("B200PC_BULK", bulk); ops.reload_tuning()
        for fn in fns:
            first = fn()
            for _ in range(15):
                assert _same(first, fn())
    torch.cuda.synchronize()
