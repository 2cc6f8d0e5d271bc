"""Per-tensor gradient error of the tensor-core training path vs the fp64 oracle (debug helper)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork
from oracle import gnn_oracle

d = np.load("tests/golden/trajectories.npz")
rng = np.random.default_rng(31)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rows = d["rows"][rng.choice(len(d["rows"]), B, replace=False)]
torch.manual_seed(9)
ref = gnn_oracle.GraphPolicyValueNetworkOracle()
with torch.no_grad():
    for layer in ref.gcn_layers: layer.bias.uniform_(-0.1, 0.1)
ref64 = copy.deepcopy(ref).double()
x, ei, batch = gnn_oracle.graph_inputs_from_rows(rows, dtype=torch.float64)
torch.manual_seed(5)
pt = torch.softmax(2 * torch.randn(B, 209), 1); vt = torch.randint(-1, 2, (B,)).float()
p64, v64 = ref64(x, ei, batch)
gnn_oracle.training_loss(p64, v64, pt.double(), vt.double())[0].backward()
gref = {n: q.grad for n, q in ref64.named_parameters()}
for prec in ("fp32", "bf16"):
    net = GNNNetwork(); net.load_state_dict(ref.state_dict()); net = net.cuda().train(); net.train_precision = prec
    p, v = net(torch.from_numpy(rows))
    (torch.nn.CrossEntropyLoss()(p, pt.cuda()) + torch.nn.MSELoss()(v.squeeze(), vt.cuda())).backward()
    print(prec, "max|dp|", (p.detach().cpu() - p64.float()).abs().max().item(), "max|dv|", (v.detach().cpu() - v64.float()).abs().max().item())
    for n, q in net.named_parameters():
        g, r = q.grad.cpu().double(), gref[n]
        print(f"   {n:28s} rel-L2 {((g - r).norm() / r.norm()).item():.3e}   |ref| {r.norm().item():.3e}")
