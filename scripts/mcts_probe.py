"""MCTS simulations/s at the bench shape, with the state of the CUDA-graph path printed (scripts/mcts_probe.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import pv_mcts
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

torch.manual_seed(0)
net = GNNNetwork().cuda().eval()
net.precision = "bf16"
dev = torch.device("cuda", 0)
out = pv_mcts.bench_sims_per_sec(net, dev, 1, timed_barrier=torch.cuda.synchronize)
print(os.environ.get("AQ_LEGAL_LANES"), out)
