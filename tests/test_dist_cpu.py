"""N>1 host logic on CPU: world_size-2 gloo processes, a CPU stand-in for the search op."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200pc import dist as bdist, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 16384, 65537):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import ref_torch
    a, b = synth.batch_pairs(3, 2, 1500)
    refs = torch.from_numpy(a); qry = torch.from_numpy(b[:, :701].copy())     # odd count -> ragged shards
    full = ref_torch.knn_topk(8, refs, qry)
    got = bdist.query_sharded(lambda s: ref_torch.knn_topk(8, refs, s), qry)
    ok1 = torch.equal(got, full)
    d_full, i_full = ref_torch.knn_points_dense(qry, refs, 4)
    d, i = bdist.query_sharded(lambda s: ref_torch.knn_points_dense(s, refs, 4), qry)
    ok2 = torch.equal(i, i_full) and torch.equal(d, d_full)
    # C5 rebuild: S/world queries per rank, ONE all_gather_into_tensor of the [S/world, B, 4] record slabs
    def pack_cpu(r, q, s_offset, peer_ptrs):
        d, i = ref_torch.knn_points_dense(q, r, 1)                                  # [B,s,1]
        nn = ref_torch.gather_rows(r, i)[:, :, 0]                                  # [B,s,3]
        rec = torch.cat([i.to(torch.int32).view(torch.float32), nn], dim=-1)       # {index bits, x, y, z}
        return rec.transpose(0, 1).contiguous()                                    # s-major like the CUDA entry
    q700 = qry[:, :700].contiguous()
    full = pack_cpu(refs, q700, 0, ())
    got_r = bdist.rebuild_sharded(refs, q700, mode="nccl", pack=pack_cpu)
    ok1 = ok1 and torch.equal(got_r.view(torch.int32), full.view(torch.int32))
    try:
        bdist.rebuild_sharded(refs, qry, mode="nccl", pack=pack_cpu)               # 701 queries do not divide over 2 ranks
        ok1 = False
    except ValueError:
        pass
    mine = bdist.batch_shard([refs, qry])
    ok3 = mine[0].shape[0] == 1 and torch.equal(mine[0][0], refs[rank])
    q.put((rank, ok1, ok2, ok3))
    dist.barrier(); dist.destroy_process_group()


def test_query_sharded_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_bench_reference_arm_under_torchrun_world2():
    """rank 0 alone measures and prints ONE json line; the other rank exits 0 without work."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--ref-queries", "512"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "knn_point_gqueries_per_s" and d["value"] > 0
    from oracle import ref_loader
    # the reference's own code when its checkout (or the staged copy oracle/_ref) is present, else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "sample" not in d["config"] and d["config"]["workload"].startswith("C2")


def test_hostio_chunk_plan():
    # C2 (8 x 16384 queries) is exactly one resident wave of the search kernel: not split; bigger batches are
    from b200pc import hostio
    assert hostio.plan_chunks(8, 16384) == 1
    assert hostio.plan_chunks(32, 16384) == 3
    assert hostio.plan_chunks(64, 16384) == 7
    assert hostio.plan_chunks(2, 1024) == 1
    assert hostio.plan_chunks(8, 16384, chunks=4) == 4 and hostio.plan_chunks(2, 16384, chunks=5) == 2
    assert hostio.plan_chunks(0, 0) == 1
