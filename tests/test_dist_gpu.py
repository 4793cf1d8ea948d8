"""Multi-GPU parity (needs >= 2 GPUs; skipped on a one-GPU box): the NCCL constraint-sharded solve equals the
single-GPU solve pass by pass.  The check itself is tools/gpu_dist_check.py, launched one rank per GPU."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("shots,native", [("auto", "1"), ("push", "0"), ("1", "1"), ("2", "0"), ("mc", "1")])
def test_sharded_equals_whole_on_two_gpus(shots, native):
    """auto picks the push exchange (multimem.red inside the pass) for these small instances; the pull forms are forced by name."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "gpu_dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env={**os.environ, "BDDB200_EXCHANGE_SHOTS": shots, "BDDB200_SHARD_NATIVE": native})
    assert r.returncode == 0 and "DIST PARITY OK" in r.stdout, r.stdout[-3000:]
