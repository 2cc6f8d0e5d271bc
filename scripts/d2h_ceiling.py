"""What the box's device -> host path gives N ranks at once, with nothing else running: every rank loops plain pinned
cudaMemcpyAsync D2H copies of the given sizes (the host path's result sizes), barrier-bracketed; rank 0 prints GB/s per GPU and in
aggregate.  The denominator for the end-to-end scaling of the host path.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_ceiling.py
    python scripts/d2h_ceiling.py            (one GPU)"""
import json
import os
import sys
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = {"world": world, "sizes": {}}
for name, nbytes in (("3.5MB_f16_wire", 3_500_000), ("7.5MB_f32_wire_mask", 7_500_000), ("14.4MB_dense", 14_400_000), ("64MB", 64 << 20)):
    src = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    dst = torch.empty((nbytes,), dtype=torch.uint8).pin_memory()
    h2d_src = torch.empty((524288,), dtype=torch.uint8).pin_memory()
    h2d_dst = torch.empty((524288,), dtype=torch.uint8, device=dev)
    reps = max(20, int(2e9 / nbytes))
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        h2d_dst.copy_(h2d_src, non_blocking=True)   # the 512 KB of packed states that go the other way each step
        dst.copy_(src, non_blocking=True)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    gbs = nbytes * reps / float(dt.item()) / 1e9
    out["sizes"][name] = {"bytes": nbytes, "gbs_per_gpu": gbs, "gbs_aggregate": gbs * world, "us_per_copy": float(dt.item()) / reps * 1e6}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
