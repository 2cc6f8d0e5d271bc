"""world_size-2 CPU (gloo) test of the multi-GPU host logic: contiguous batch sharding, the
"each rank's loss is divided by the GLOBAL batch, gradients are summed with one all-reduce" rule of
train_network.FlatTrainer, and the game sharding / history gather of self_play.  The CUDA kernels are
replaced by the CPU oracle here (test infrastructure); the arithmetic identity is what is checked."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, rows, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alphaquoridorgnn_b200.train_network import shard_bounds
    from oracle import gnn_oracle
    torch.manual_seed(0)
    model = gnn_oracle.GraphPolicyValueNetworkOracle().double()
    B = rows.shape[0]
    torch.manual_seed(1)
    pt = torch.softmax(torch.randn(B, 209, dtype=torch.float64), 1)
    vt = torch.randint(-1, 2, (B,)).double()
    lo, hi = shard_bounds(B, rank, world)
    x, ei, batch = gnn_oracle.graph_inputs_from_rows(rows[lo:hi], dtype=torch.float64)
    p, v = model(x, ei, batch)
    # per-rank loss divided by the GLOBAL batch (aq_loss_grad's B_total), then summed by all-reduce
    lp = -(pt[lo:hi] * torch.log_softmax(p, 1)).sum() / B
    lv = ((v.squeeze(1) - vt[lo:hi]) ** 2).sum() / B
    (lp + lv).backward()
    flat = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    dist.all_reduce(flat)
    # game sharding + history gather as in self_play.self_play
    games = 7
    mine = games // world + (1 if rank < games % world else 0)
    history = [[rank, g] for g in range(mine)]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(history, gathered, dst=0)
    if rank == 0:
        out["flat"] = flat.clone()
        out["games"] = [h for part in gathered for h in part]
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_identity_and_game_sharding(traj):
    from alphaquoridorgnn_b200.train_network import shard_bounds
    from oracle import gnn_oracle
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]  # empty shards are legal
    rows = traj["rows"][::4000][:9]
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, rows, out), nprocs=2, join=True)
    torch.manual_seed(0)
    model = gnn_oracle.GraphPolicyValueNetworkOracle().double()
    B = rows.shape[0]
    torch.manual_seed(1)
    pt = torch.softmax(torch.randn(B, 209, dtype=torch.float64), 1)
    vt = torch.randint(-1, 2, (B,)).double()
    x, ei, batch = gnn_oracle.graph_inputs_from_rows(rows, dtype=torch.float64)
    p, v = model(x, ei, batch)
    loss, _, _ = gnn_oracle.training_loss(p, v, pt, vt)
    loss.backward()
    want = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    assert (out["flat"] - want).abs().max().item() <= 1e-12
    assert sorted(out["games"]) == [[0, 0], [0, 1], [0, 2], [0, 3], [1, 0], [1, 1], [1, 2]]
