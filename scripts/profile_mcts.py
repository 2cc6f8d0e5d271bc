"""Where does a lock-step MCTS simulation step go? (events around select / leaf eval / expand)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphaquoridorgnn_b200 import _lib, positions, pv_mcts
from alphaquoridorgnn_b200 import game_logic as gl
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

G, sims = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 200
torch.manual_seed(0)
net = GNNNetwork().cuda().eval(); net.precision = "bf16"
L = _lib.load(); P = _lib.ptr
roots = positions.random_positions(G, seed=5, games=G)
max_nodes = 1 + sims * pv_mcts.MAX_CHILDREN
ws = torch.empty((L.aq_mcts_ws_bytes(G, max_nodes),), dtype=torch.uint8, device="cuda")
leaf = torch.empty((G, 32), dtype=torch.uint8, device="cuda"); kind = torch.empty((G,), dtype=torch.int32, device="cuda")
st = _lib.stream_ptr()
L.aq_mcts_reset(P(ws), P(roots), G, max_nodes, st)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = [0.0, 0.0, 0.0]
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(sims):
    ev[0].record()
    L.aq_mcts_select(P(ws), G, max_nodes, 1.25, P(leaf), P(kind), st)
    ev[1].record()
    out = net.predict_batch(leaf)
    ev[2].record()
    L.aq_mcts_expand_backup(P(ws), G, max_nodes, P(out["priors"]), P(out["value"]), P(out["mask"]), P(out["pawn"]), st)
    ev[3].record()
    if s >= 20:
        torch.cuda.synchronize()
        for k in range(3): acc[k] += ev[k].elapsed_time(ev[k + 1])
torch.cuda.synchronize(); wall = time.perf_counter() - t0
n = sims - 20
print(f"G={G}: select {acc[0]/n*1e3:.1f} us, leaf_eval {acc[1]/n*1e3:.1f} us, expand+backup {acc[2]/n*1e3:.1f} us per simulation step; wall/step {(wall/sims)*1e6:.1f} us (with per-step sync)")
# without syncs
mcts = pv_mcts.BatchedMCTS(net, sims)
mcts.search(roots[:64], 4)
torch.cuda.synchronize(); t0 = time.perf_counter()
mcts.search(roots)
torch.cuda.synchronize(); wall = time.perf_counter() - t0
print(f"search(): {wall/sims*1e6:.1f} us per simulation step -> {G*sims/wall/1e6:.2f} M sims/s")
