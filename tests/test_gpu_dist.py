"""Query-sharded kNN over NCCL (one process per GPU).  Needs >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from b200pc import dist as bdist, pointnet2_utils as P, synth
from oracle import strict
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
a, b = synth.batch_pairs(4, 1, 8192)
refs = torch.from_numpy(a).to(dev); qry = torch.from_numpy(b[:, :5001].copy()).to(dev)
idx = bdist.query_sharded(lambda q: P.knn_point(16, refs, q), qry)
ball = bdist.query_sharded(lambda q: P.query_ball_point(1.0, 32, refs, q), qry)
ok = np.array_equal(idx.cpu().numpy(), strict.knn_point(16, a, b[:, :5001])) and \
     np.array_equal(ball.cpu().numpy(), strict.query_ball_point(1.0, 32, a, b[:, :5001]))
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("SHARDED_OK" if flag.item() == 1 else "SHARDED_MISMATCH")
dist.barrier(); dist.destroy_process_group()
'''


def test_query_sharded_knn_nccl(cuda_dev, tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT, os.path.join(ROOT, "point-cloud-interpolation-_b200")))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 2)), "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
