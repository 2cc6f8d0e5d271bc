"""Host-side timeline of HostLeafEvaluator.evaluate at the bench shape (AQ_HOST_TRACE=1 prints it per call)."""
import os
import sys
import time

os.environ["AQ_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import positions
from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork, HostLeafEvaluator

B = 16384
torch.manual_seed(0)
net = GNNNetwork().cuda().eval()
net.precision = "bf16"
_, batches = positions.mixed_batches(2, B, seed=1)
hst = [torch.from_numpy(gl.pack_rows_host(*[t.cpu().numpy() for t in gl.unpack_rows(b)])).pin_memory() for b in batches]
ev = HostLeafEvaluator(net, B)
for i in range(8):
    t0 = time.perf_counter()
    out = ev.evaluate(B, states=hst[i % 2])
    print(f"python evaluate() {1e6 * (time.perf_counter() - t0):.0f} us, legal/board {out['offsets'][B] / B:.1f}", file=sys.stderr)
