"""Training-step time of FlatTrainer's variants at the bench shapes (also under torchrun: per-GPU batch, data parallel) (CUDA events, L2 flushed): fused step replayed as a CUDA graph,
fused step eager, and the separate entry points (collective='nccl' at world 1 = aq_loss_grad + aq_gnn_backward + aq_adam_step).
   python scripts/train_variants.py [profile]      ('profile': one eager fused step and one unfused step between cudaProfilerStart/Stop)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import positions, train_network
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
allpos, _ = positions.mixed_batches(1, 8192, seed=1, device=dev)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
profile = len(sys.argv) > 1 and sys.argv[1] == "profile"


def timed(fn, n=20, warm=5):
    for i in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for i in range(n):
        flush.fill_(i)
        evs[i][0].record()
        fn()
        evs[i][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / n * 1e3


for TB in (256, 4096):
    tb = allpos[:TB].contiguous()
    torch.manual_seed(1)
    pt = torch.softmax(torch.randn(TB, 209, device=dev), 1)
    vt = torch.randint(-1, 2, (TB,), device=dev).float()
    row = {}
    for name, kw in (("fused+graph", {}), ("fused eager", {"use_graph": False}), ("separate kernels", {"collective": "nccl"})):
        torch.manual_seed(0)
        net = GNNNetwork().to(dev).train()
        tr = train_network.FlatTrainer(net, precision="bf16", rank=rank, world_size=world, **kw)
        if profile:
            if name == "fused+graph":
                continue
            for _ in range(3):
                tr.step(tb, pt, vt, TB * world)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            flush.fill_(1)
            tr.step(tb, pt, vt, TB * world)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        row[name] = timed(lambda: tr.step(tb, pt, vt, TB * world))
    if not profile and rank == 0:
        print(f"B={TB}: " + ", ".join(f"{k} {v:.1f} us" for k, v in row.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
print("done")
