"""Whole lock-step self-play games (bench.py's extra.mcts workload), first call (graph captures) and repeats."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import self_play
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

games, sims = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 200
torch.manual_seed(0)
net = GNNNetwork().cuda().eval()
net.precision = "bf16"
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = self_play.play_batch_device(net, games, "cuda", sims=sims, seed=7 + rep, policy_dtype=torch.float32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt:.3f} s, {rec['sims'] / dt / 1e6:.1f} M simulations/s, {int(rec['states'].shape[0])} positions")
