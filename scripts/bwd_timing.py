"""Per-phase cycle accounting of gcn_backward_tc2_kernel (debug variant built with -DTC2B_TIMING=1):
python scripts/build_debug_lib.py timing <source>.cu -D...=1;  python scripts/bwd_timing.py alphaquoridorgnn_b200/debug/libaqgnn_timing.so"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphaquoridorgnn_b200 import _lib, positions
import alphaquoridorgnn_b200.build as _b
_lib.LIB_PATH = os.path.abspath(sys.argv[1])   # the debug copy, loaded explicitly
_b.needs_build = lambda: False
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork
TB = 4096
L = _lib.load(); P = _lib.ptr
net = GNNNetwork().cuda(); flat = net.flat_parameters().clone()
_, (tb,) = positions.mixed_batches(1, TB, seed=3)
saved = torch.empty((L.aq_gnn_saved_floats(TB),), device="cuda"); bws = torch.empty((L.aq_gnn_backward_ws_floats(TB),), device="cuda")
tp = torch.empty((TB, 209), device="cuda"); tv = torch.empty((TB,), device="cuda"); dp = torch.randn_like(tp) * 1e-4; dv = torch.randn_like(tv) * 1e-4
grads = torch.empty_like(flat); st = _lib.stream_ptr()
dbg = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_longlong * 16)()
for it in range(2):
    L.aq_gnn_forward(P(flat), P(tb), None, None, TB, P(tp), P(tv), P(saved), 1, st)
    L.aq_gnn_backward(P(flat), P(saved), P(dp), P(dv), TB, P(grads), P(bws), 1, st)
    torch.cuda.synchronize()
    dbg.aq_debug_bwd_timing(out)
names = ["loop top: mask3 / dg loads", "cp.async wait + sync + next prefetch + adjacency scatter", "ReLU masks of layers 2, 1 from the tiles",
         "dY: ld / mask / st (x2)", "aggregation MMA wait (x2)", "dZ -> bf16 tile (x2)", "dX + dW MMAs wait (x2)", "layer 1: dY1 -> tile", "dW1 MMA wait"]
boards = (TB + 147) // 148
tot = sum(out[i] for i in range(9))
for i, n in enumerate(names):
    print(f"{n:58s} {out[i]/boards:8.0f} cycles/board {100*out[i]/tot:5.1f}%")
print("total per board", tot / boards)
