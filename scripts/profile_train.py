"""Per-call device time of one training step (events), fp32 vs bf16 trunk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphaquoridorgnn_b200 import _lib, positions
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

TB = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = _lib.load(); P = _lib.ptr
torch.manual_seed(0)
net = GNNNetwork().cuda(); flat = net.flat_parameters().clone()
tb = positions.random_positions(TB, seed=3, games=max(64, TB // 8))
pt = torch.softmax(torch.randn(TB, 209, device="cuda"), 1); vt = torch.randint(-1, 2, (TB,), device="cuda").float()
saved = torch.empty((L.aq_gnn_saved_floats(TB),), device="cuda"); bws = torch.empty((L.aq_gnn_backward_ws_floats(TB),), device="cuda")
tp = torch.empty((TB, 209), device="cuda"); tv = torch.empty((TB,), device="cuda"); dp = torch.empty_like(tp); dv = torch.empty_like(tv)
grads = torch.empty_like(flat); m1 = torch.zeros_like(flat); m2 = torch.zeros_like(flat); loss = torch.zeros(2, device="cuda")
st = _lib.stream_ptr()
for prec in (0, 1):
    calls = [
        ("forward", lambda: L.aq_gnn_forward(P(flat), P(tb), None, None, TB, P(tp), P(tv), P(saved), prec, st)),
        ("loss_grad", lambda: L.aq_loss_grad(P(tp), P(tv), P(pt), P(vt), TB, TB, P(loss), P(dp), P(dv), st)),
        ("backward", lambda: L.aq_gnn_backward(P(flat), P(saved), P(dp), P(dv), TB, P(grads), P(bws), prec, st)),
        ("adam", lambda: L.aq_adam_step(P(flat), P(grads), P(m1), P(m2), flat.numel(), 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, st)),
    ]
    out = []
    for name, fn in calls:
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
    print(f"B={TB} precision={prec}: " + ", ".join(out))
