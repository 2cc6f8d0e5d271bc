"""The GNN oracle has no reference output to pin against (torch_geometric is not vendored, pinned or
installed -- parity unpinned, see oracle/gnn_oracle.py).  What can be checked on the CPU: the
edge-list (index_add) restatement agrees with an independent dense A_hat restatement in float64
built from the reference-derived golden open masks, and the structural facts of SURVEY.md."""
import numpy as np
import torch

from oracle import gnn_oracle


def test_edge_list_and_dense_restatements_agree(traj):
    rows = traj["rows"][::2500][:12]
    torch.manual_seed(0)
    net = gnn_oracle.GraphPolicyValueNetworkOracle().double()
    with torch.no_grad():
        for layer in net.gcn_layers:
            layer.bias.uniform_(-0.1, 0.1)
    x, ei, batch = gnn_oracle.graph_inputs_from_rows(rows, dtype=torch.float64)
    with torch.no_grad():
        p, v = net(x, ei, batch)
    p2, v2 = gnn_oracle.dense_forward_fp64(net.state_dict(), rows)
    assert np.abs(p.numpy() - p2).max() < 1e-12 and np.abs(v.numpy() - v2).max() < 1e-12
    assert p.shape == (len(rows), 209) and v.shape == (len(rows), 1)


def test_parameter_inventory():
    net = gnn_oracle.GraphPolicyValueNetworkOracle()
    sd = net.state_dict()
    assert sum(t.numel() for t in sd.values()) == 64082
    assert tuple(sd["gcn_layers.0.lin.weight"].shape) == (128, 6)
    assert tuple(sd["policy_head.2.weight"].shape) == (209, 64)
    assert float(sd["gcn_layers.1.bias"].abs().max()) == 0.0  # GCNConv bias init: zeros


def test_gcn_norm_handles_existing_self_loops_and_isolated_nodes():
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 1, 2]])  # node 1 has a self loop, node 2 only a self loop, node 3 isolated
    src, dst, w = gnn_oracle.gcn_norm(ei, 4, torch.float64)
    assert src.numel() == 2 + 4  # existing self loops dropped, one per node appended
    deg = torch.zeros(4, dtype=torch.float64).index_add_(0, dst, torch.ones(6, dtype=torch.float64))
    assert deg.tolist() == [2.0, 2.0, 1.0, 1.0]
    assert torch.allclose(w[-1], torch.tensor(1.0, dtype=torch.float64))


def test_loss_restates_double_softmax():
    torch.manual_seed(1)
    p = torch.softmax(torch.randn(4, 209), 1)
    t = torch.softmax(torch.randn(4, 209), 1)
    v, vt = torch.tanh(torch.randn(4, 1)), torch.randn(4)
    loss, lp, lv = gnn_oracle.training_loss(p, v, t, vt)
    want_p = -(t * torch.log_softmax(p, 1)).sum(1).mean()
    want_v = ((v.squeeze() - vt) ** 2).mean()
    assert torch.allclose(lp, want_p) and torch.allclose(lv, want_v) and torch.allclose(loss, want_p + want_v)
