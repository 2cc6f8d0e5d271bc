"""Two training steps (bf16 tensor-core pair) at batch TB between cudaProfilerStart/Stop, for an ncu launch list:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --csv python scripts/train_launches.py 4096"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphaquoridorgnn_b200 import _lib, positions

TB = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = _lib.load(); P = _lib.ptr
torch.manual_seed(0)
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork
net = GNNNetwork().cuda(); flat = net.flat_parameters().clone()
_, (tb,) = positions.mixed_batches(1, TB, seed=3)
pt = torch.softmax(torch.randn(TB, 209, device="cuda"), 1); vt = torch.randint(-1, 2, (TB,), device="cuda").float()
saved = torch.empty((L.aq_gnn_saved_floats(TB),), device="cuda"); bws = torch.empty((L.aq_gnn_backward_ws_floats(TB),), device="cuda")
tp = torch.empty((TB, 209), device="cuda"); tv = torch.empty((TB,), device="cuda"); dp = torch.empty_like(tp); dv = torch.empty_like(tv)
grads = torch.empty_like(flat); m1 = torch.zeros_like(flat); m2 = torch.zeros_like(flat); loss = torch.zeros(2, device="cuda")
st = _lib.stream_ptr()

def step(i):
    L.aq_gnn_forward(P(flat), P(tb), None, None, TB, P(tp), P(tv), P(saved), 1, st)
    L.aq_loss_grad(P(tp), P(tv), P(pt), P(vt), TB, TB, P(loss), P(dp), P(dv), st)
    L.aq_gnn_backward(P(flat), P(saved), P(dp), P(dv), TB, P(grads), P(bws), 1, st)
    L.aq_adam_step(P(flat), P(grads), P(m1), P(m2), flat.numel(), i + 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, st)

for i in range(3): step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(2): step(3 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
