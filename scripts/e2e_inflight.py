"""End-to-end host path (bench.py's e2e loop) by number of batches in flight and wire format, all ranks at once:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/e2e_inflight.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import game_logic as gl, positions
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork, HostLeafEvaluator

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
B, K, nb = 16384, 40, 4
torch.manual_seed(0)
net = GNNNetwork().to(dev).eval()
net.precision = "bf16"
_, batches = positions.mixed_batches(nb, B, seed=1 + rank, device=dev)
hst = [torch.from_numpy(gl.pack_rows_host(*[t.cpu().numpy() for t in gl.unpack_rows(b)])).pin_memory() for b in batches]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def run(inflight, **kw):
    evs = [HostLeafEvaluator(net, B, **kw) for _ in range(inflight)]

    def loop(n):
        for j in range(min(inflight - 1, n)):
            evs[j].submit(B, states=hst[j % nb])
        for i in range(n):
            j = i + inflight - 1
            if j < n:
                evs[j % inflight].submit(B, states=hst[j % nb])
            evs[i % inflight].wait()

    loop(2 * inflight + 2)
    barrier()
    t0 = time.perf_counter()
    loop(K)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    for ev in evs:
        ev.close()
    return world * B * K / float(dt.item()) / 1e6


out = {}
if os.environ.get("TRUNK_CTAS"):
    import ctypes
    from alphaquoridorgnn_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for cap in [int(c) for c in os.environ["TRUNK_CTAS"].split(",")]:
        lib.aq_debug_trunk_ctas(cap)
        for inflight in (3, 4):
            out[f"f16 x{inflight} trunk {cap} CTAs"] = round(run(inflight, wire="f16", with_mask=False), 1)
    print(json.dumps(out))
    sys.exit(0)
for inflight in (2, 3, 4, 6, 8):
    out[f"f16 x{inflight}"] = round(run(inflight, wire="f16", with_mask=False), 1)
out["f32+mask x3"] = round(run(3, wire="f32", with_mask=True), 1)
out["f32+mask x6"] = round(run(6, wire="f32", with_mask=True), 1)
if rank == 0:
    print(json.dumps({"world": world, "M board-evals/s": out}))
if world > 1:
    dist.destroy_process_group()
