"""Execution of the learning cycle -- drop-in for reference train_cycle.py:18-44.

Single process:   python -m alphaquoridorgnn_b200.train_cycle
One box, N GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
                      -m alphaquoridorgnn_b200.train_cycle
(self-play games are sharded over the ranks without communication; training is data parallel with one
NCCL all-reduce of the flat gradient per step; rank 0 evaluates and promotes.)"""
import os

import torch

from .constants import BOARD_SIZE, PV_NETWORK_NAME, PV_NETWORK_PATH
from .evaluate_network import evaluate_network
from .pv_network_gnn import create_network
from .self_play import self_play
from .train_network import train_network

NUM_TRAIN_CYCLE = 1000  # Number of training cycles (train_cycle.py:18)


def main(num_cycles=NUM_TRAIN_CYCLE):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    if rank == 0:
        print(f'Model {PV_NETWORK_NAME} on board size {BOARD_SIZE}')
        create_network(PV_NETWORK_PATH + 'best.pth')
    for i in range(num_cycles):
        if world > 1:
            torch.distributed.barrier()
        if rank == 0:
            print(f'\nBegin training cycle {i + 1}/{num_cycles} ====================')
        self_play(rank=rank, world_size=world)
        if world > 1:
            torch.distributed.barrier()
        train_network(rank=rank, world_size=world)
        if rank == 0:
            evaluate_network()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
