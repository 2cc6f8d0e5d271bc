"""Time of the fused all-reduce + Adam kernel (peer memory) against torch.distributed.all_reduce + aq_adam_step on the same flat
gradient, ranks aligned by a barrier before every call (CUDA events, per rank; rank 0 prints the max over ranks of the means).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_adam_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from alphaquoridorgnn_b200 import _lib, train_network

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L, P = _lib.load(), _lib.ptr
n = L.aq_param_count()
torch.manual_seed(rank)
params = torch.randn(n, device=dev)
grads = torch.randn(n, device=dev) * 0.01
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
comm = train_network.PeerCommunicator(rank, world, dev)
st = _lib.stream_ptr(dev)


def run(fn, reps=60):
    ts = []
    for i in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[10:])
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fused(i):
    _lib.check(L.aq_dp_adam_step(comm.handle, P(params), P(grads), P(m), P(v), 1e-3, 0.9, 0.999, 1e-8, st), "aq_dp_adam_step")


def nccl(i):
    if world > 1:
        dist.all_reduce(grads)
    _lib.check(L.aq_adam_step(P(params), P(grads), P(m), P(v), n, i + 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, st), "aq_adam_step")


t_f, t_n = run(fused), run(nccl)
if rank == 0:
    print(f"world {world}: fused peer-memory all-reduce + Adam {t_f:.1f} us; torch.distributed.all_reduce + aq_adam_step {t_n:.1f} us "
          f"(median of 50, ranks aligned by a barrier before every call; includes the wait for the slowest rank's arrival)")
steps, status = comm.status()
assert status == 0, "time-out in the peer exchange"
if world > 1:
    dist.destroy_process_group()
