"""Multi-GPU (NCCL) check, only when the box has >= 2 GPUs: scripts/dp_check.py under torchrun -- data-parallel
FlatTrainer steps equal the single-GPU step on the full batch (fp32 and bf16), ranks stay bit-identical, and
self-play shards games over the ranks.  The world_size-2 gloo test (tests/test_dp_gloo.py) covers the same
host logic on CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_data_parallel_and_sharded_self_play():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "dp_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
