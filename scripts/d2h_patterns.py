"""Which part of the end-to-end pattern costs device -> host bandwidth when every GPU of the box copies at once?  All ranks together:
 (a) plain pinned D2H loop, 3.5 MB copies, one stream
 (b) the same with a concurrent H2D loop of 0.5 MB copies on a second stream
 (c) three streams round-robin, each copy behind a ~190 us kernel on its stream (the shape of the pipelined host path, no library)
 (d) as (c) plus the 0.5 MB H2D copy in front of each kernel
 (e) as (d) with an event record per step and the host waiting on the event of the step three back (what HostLeafEvaluator.wait does)
 (f) as (e), but all D2H copies on one copy stream and all H2D copies on another, ordered with the kernels by events
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_patterns.py"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
NB, HB, STEPS = 3_490_000, 524_288, 200
src = [torch.empty((NB,), dtype=torch.uint8, device=dev) for _ in range(3)]
dst = [torch.empty((NB,), dtype=torch.uint8).pin_memory() for _ in range(3)]
hsrc = [torch.empty((HB,), dtype=torch.uint8).pin_memory() for _ in range(3)]
hdst = [torch.empty((HB,), dtype=torch.uint8, device=dev) for _ in range(3)]
work = torch.empty((1 << 22,), dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]


def busy(n=1):   # ~190 us of kernel time (calibrated below)
    for _ in range(n):
        work.mul_(1.0001)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(fn):
    fn(10)
    barrier()
    t0 = time.perf_counter()
    fn(STEPS)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


# calibrate the busy kernel to ~190 us
torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
busy(3); e0.record(); busy(10); e1.record(); torch.cuda.synchronize(dev)
REP = max(1, round(190.0 / (e0.elapsed_time(e1) * 100.0)))   # elapsed ms for 10 -> us each = ms * 100


def a(n):
    with torch.cuda.stream(streams[0]):
        for _ in range(n):
            dst[0].copy_(src[0], non_blocking=True)


def b(n):
    for _ in range(n):
        with torch.cuda.stream(streams[0]):
            dst[0].copy_(src[0], non_blocking=True)
        with torch.cuda.stream(streams[3]):
            hdst[0].copy_(hsrc[0], non_blocking=True)


def c(n, h2d=False, events=False):
    evs = []
    for i in range(n):
        s = i % 3
        with torch.cuda.stream(streams[s]):
            if h2d:
                hdst[s].copy_(hsrc[s], non_blocking=True)
            busy(REP)
            dst[s].copy_(src[s], non_blocking=True)
            if events:
                ev = torch.cuda.Event(blocking=True)
                ev.record()
                evs.append(ev)
        if events and i >= 2:
            evs[i - 2].synchronize()


d2h_stream, h2d_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def f(n, events=True):
    """as (e), but every D2H copy goes through ONE copy stream and every H2D copy through another (event-ordered with the kernels)"""
    done = []
    for i in range(n):
        s = i % 3
        with torch.cuda.stream(h2d_stream):
            hdst[s].copy_(hsrc[s], non_blocking=True)
            up = torch.cuda.Event()
            up.record()
        with torch.cuda.stream(streams[s]):
            streams[s].wait_event(up)
            busy(REP)
            k = torch.cuda.Event()
            k.record()
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(k)
            dst[s].copy_(src[s], non_blocking=True)
            ev = torch.cuda.Event(blocking=True)
            ev.record()
            done.append(ev)
        if events and i >= 2:
            done[i - 2].synchronize()


out = {"world": world, "busy_kernel_reps": REP}
for name, fn in (("a_plain_d2h", a), ("b_d2h_and_h2d", b), ("c_three_streams_kernel_then_d2h", c),
                 ("d_plus_h2d", lambda n: c(n, True)), ("e_plus_event_wait", lambda n: c(n, True, True)), ("f_shared_copy_streams", f), ("a_again", a)):
    dt = timed(fn)
    out[name] = {"d2h_gbs_per_gpu": round(NB * STEPS / dt / 1e9, 2), "us_per_step": round(dt / STEPS * 1e6, 1)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
