"""Exactly N leaf-evaluation steps at the bench shape (B = 16384, bf16) bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off ... python scripts/profile_step.py [N]
Used for the committed ncu captures under profiles/ (same kernels and inputs as bench.py's timed region)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import _lib, positions
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 16384
torch.manual_seed(0)
net = GNNNetwork().cuda().eval()
net.precision = os.environ.get("AQ_PRECISION", "bf16")
_, batches = positions.mixed_batches(4, B, seed=1)  # bench.py's workload
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
for i in range(3):
    net.predict_batch(batches[i % 4])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(N):
    flush.fill_(i)
    net.predict_batch(batches[i % 4])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
