"""One lock-step search (4,096 roots x 200 simulations) by simulations per CUDA-graph launch: host time spent
launching against total time (is the search host-bound?), repeated alternately."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import positions, pv_mcts
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

G, sims = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 200
torch.manual_seed(0)
dev = torch.device("cuda", 0)
net = GNNNetwork().to(dev).eval()
net.precision = "bf16"
print("host cpus:", os.cpu_count(), open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"))
roots = positions.random_positions(G, seed=5, games=G)
searchers = {}
for chunk in (1, 5, 25, 100):
    pv_mcts.GRAPH_CHUNK = chunk
    searchers[chunk] = pv_mcts.BatchedMCTS(net, sims, device=dev)
    searchers[chunk].search(roots)          # captures
for rep in range(3):
    for chunk, m in searchers.items():
        pv_mcts.GRAPH_CHUNK = chunk
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        m.search(roots)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize(dev)
        print(f"rep {rep} chunk {chunk:3d}: search() returned after {1e3 * (t1 - t0):7.2f} ms, device span {e0.elapsed_time(e1):7.2f} ms "
              f"-> {e0.elapsed_time(e1) / sims * 1e3:6.1f} us per simulation step")
