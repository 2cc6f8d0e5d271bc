"""Multi-GPU check (run under torchrun, NCCL): data-parallel training with one all-reduce of the flat gradient
per step gives the same parameters as the single-GPU run on the full batch, and self-play shards games over
the ranks with the history gathered on rank 0.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from alphaquoridorgnn_b200 import game_logic as gl, positions, pv_mcts, self_play, train_network
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
B = 250  # not divisible by 4/8: shards differ in size
packed = positions.random_positions(B, seed=7, games=64)
torch.manual_seed(1)
pt = torch.softmax(torch.randn(B, 209), 1).cuda(); vt = torch.randint(-1, 2, (B,)).float().cuda()
for prec in ("fp32", "bf16"):
    torch.manual_seed(2)
    dp_net = GNNNetwork().cuda().train()
    ref_net = GNNNetwork().cuda().train()
    ref_net.load_state_dict(dp_net.state_dict())
    dp = train_network.FlatTrainer(dp_net, rank=rank, world_size=world, precision=prec)
    one = train_network.FlatTrainer(ref_net, precision=prec)
    lo, hi = train_network.shard_bounds(B, rank, world)
    for step in range(3):
        l_dp = dp.step(packed[lo:hi].contiguous(), pt[lo:hi].contiguous(), vt[lo:hi].contiguous(), B).clone()
        dist.all_reduce(l_dp)
        l_one = one.step(packed, pt, vt, B)
        gerr = ((dp.grads - one.grads).norm() / one.grads.norm()).item()
        if rank == 0:
            print(f"{prec} step {step}: loss dp {l_dp.sum().item():.6f} single {l_one.sum().item():.6f}  grad rel-L2 {gerr:.2e}")
        assert abs(l_dp.sum().item() - l_one.sum().item()) < 1e-5 and gerr < (1e-5 if prec == "fp32" else 2e-2)
    # every rank holds identical parameters after the steps
    mine = dp.flat.clone(); ref0 = mine.clone(); dist.broadcast(ref0, 0)
    assert torch.equal(mine, ref0), "ranks diverged"
# sharded self-play: 10 games over the ranks, rank 0 gets the merged history
pv_mcts.PV_EVALUATE_COUNT = 8
torch.manual_seed(3)
net = GNNNetwork().cuda().eval()
mine = 10 // world + (1 if rank < 10 % world else 0)
hist, info = self_play.play_batch(net, mine, sims=8, seed=100 + rank)
gathered = [None] * world if rank == 0 else None
dist.gather_object(hist, gathered, dst=0)
if rank == 0:
    total = sum(len(h) for h in gathered)
    print(f"self-play: {world} ranks, 10 games, {total} positions gathered on rank 0")
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("dp_check ok")
