#!/usr/bin/env python
"""Benchmark of the AlphaQuoridorGNN hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batched leaf evaluation (BaseNetwork.predict semantics for B states: legal-move
mask + graph build + GNN forward + restriction to legal actions) over one batch of B=16384
synthetic legal positions per GPU  -- BASELINE.json configs[2].  Headline metric: GNN board-evals/s.
The other two BASELINE metrics (legal-mask positions/s for 1M positions, train samples/s at B=256)
and, when available, MCTS sims/s are measured after the headline and reported under "extra".

`--impl reference` times the CPU port of the reference path (oracle/: C restatement of game_logic
+ torch restatement of pv_network_gnn) on the host cores; the reference itself is pure Python with
un-vendored dependencies and cannot travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_BOARD_FWD = 5.79e6
FLOP_PER_BOARD_FWDBWD = 16.9e6
BYTES_PER_POSITION_LEGAL = 72    # 32 B packed state in + 32 B mask + 8 B ordered pawn list out
BYTES_PER_BOARD_HEADS = 128 * 4 + 209 * 4 + 4 + 32
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu captures (profiles/*.csv)
NCU_TRAFFIC_BYTES = {("gcn_forward_kernel", 1, 16384): 632064}  # profiles/r1_v3_kernels_ncu.csv: dram read 0.632064 MB + write 0 of gcn_forward_tc2_kernel (outputs stay in L2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device) BEFORE any pinned host buffer is
    allocated, so that the buffers of the host-buffer path live on the GPU's NUMA node.  Best effort: returns a description."""
    if os.environ.get("AQ_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out  # nvidia-smi prints an 8-digit domain; sysfs uses 4
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {index} at {bdf}: no usable local cpus ({cpulist})"
        os.sched_setaffinity(0, cpus)
        return f"gpu {index} at {bdf}: bound to {len(cpus)} local cpus ({cpulist})"
    except Exception as e:  # sysfs not visible in this container, single-socket host, ...
        return f"unavailable ({type(e).__name__}: {e})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.t = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=15)  # a query in flight (nvidia-smi takes seconds on a fresh box) still belongs to the timed region
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# reference arm: CPU port (oracle/) of the same step
# ------------------------------------------------------------------------------------------------
def oracle_random_positions(n, seed):
    """Lock-step random games on the CPU with the C oracle (same move mix as positions.py)."""
    from oracle import quoridor_oracle as qo
    rng = np.random.default_rng(seed)
    out_rows, out_plies, total = [], [], 0
    while total < n:
        G = max(64, n // 32)  # ~32 plies deep, the same mix of game phases as positions.random_positions in run_ours
        rows = np.zeros((G, 68), np.uint8)
        rows[:, [0, 2]] = 76
        rows[:, [1, 3]] = 10
        plies = np.zeros(G, np.int16)
        for _ in range(116):
            if len(rows) == 0 or total >= n:
                break
            la = qo.legal_actions_batch(rows, plies)
            ok = la["n"] > 0
            rows, plies = rows[ok], plies[ok]
            acts, cnt = la["actions"][ok], la["n"][ok]
            out_rows.append(rows)
            out_plies.append(plies)
            total += len(rows)
            is_wall = (acts >= 81)
            valid = acts >= 0
            nw, npn = (is_wall & valid).sum(1), (~is_wall & valid).sum(1)
            use_wall = (nw > 0) & ((rng.random(len(rows)) < 0.5) | (npn == 0))
            pick = np.where(use_wall, npn + (rng.random(len(rows)) * nw).astype(np.int64),
                            (rng.random(len(rows)) * np.maximum(npn, 1)).astype(np.int64))
            pick = np.minimum(pick, cnt - 1)
            a = acts[np.arange(len(rows)), pick]
            rows, plies, flags = qo.next_batch(rows, plies, a)
            rows, plies = rows[flags == 0], plies[flags == 0]
    return np.concatenate(out_rows)[:n], np.concatenate(out_plies)[:n]


def oracle_leaf_eval(model, rows, plies, threads):
    """CPU port of one step: predict() semantics for a batch (pv_network_cnn.py:117-137 behaviour)."""
    from oracle import gnn_oracle, quoridor_oracle as qo
    la = qo.legal_actions_batch(rows, plies, nthreads=threads)
    x, ei, batch = gnn_oracle.graph_inputs_from_rows(rows)
    with torch.inference_mode():
        p, v = model(x, ei, batch)
    dense = torch.from_numpy(np.unpackbits(la["mask"].view(np.uint8), axis=1, bitorder="little")[:, :209].astype(bool))
    p = torch.where(dense, p, torch.zeros_like(p))
    s = p.sum(1, keepdim=True)
    p = p / torch.where(s == 0, torch.ones_like(s), s)
    return p, v


def time_cpu_port(sample, steps, warmup, seed=1):
    from oracle import gnn_oracle, quoridor_oracle as qo
    qo.build()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = gnn_oracle.GraphPolicyValueNetworkOracle().eval()
    rows, plies = oracle_random_positions(sample, seed)
    for _ in range(warmup):
        oracle_leaf_eval(model, rows, plies, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_leaf_eval(model, rows, plies, threads)
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 2048
    value, sec_per_step, threads = time_cpu_port(sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "gnn_board_evals_per_sec", "value": value, "unit": "board-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "leaf_eval (BASELINE configs[2]: batched predict = legal mask + GNN forward + legal renorm)",
                   "batch_per_step": sample, "note": "CPU port of the reference path (oracle/); the reference is pure "
                   "Python with un-vendored torch_geometric and cannot run on the GPU box"},
        "cpu_baseline": {"value": value, "unit": "board-evals/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} synthetic positions per step, C oracle legal_actions (OpenMP) + torch CPU GNN forward"},
        "e2e": {"value": value, "unit": "board-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from alphaquoridorgnn_b200 import _lib, positions
    from alphaquoridorgnn_b200 import game_logic as gl
    from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork, PRECISIONS

    numa = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = _lib.load()
    pk = peaks()
    B, K, Wm = args.batch, args.steps, args.warmup
    prec = PRECISIONS[args.precision]

    torch.manual_seed(0)
    net = GNNNetwork().to(dev).eval()
    net.precision = args.precision
    flat = net.flat_parameters()
    # inference weights prepared once, as prep_for_inference does (the weights do not change between evaluations)
    prep = net.prepared_weights() if prec == 1 else None
    nb = 4
    allpos, batches = positions.mixed_batches(nb, B, seed=1 + rank, device=dev)
    priors = torch.empty((B, 209), dtype=torch.float32, device=dev)
    value = torch.empty((B,), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 8), dtype=torch.int32, device=dev)
    pawn = torch.empty((B, 8), dtype=torch.uint8, device=dev)
    pooled = torch.empty((L.aq_leaf_eval_ws_floats(B),), dtype=torch.float32, device=dev)  # leaf-eval workspace: pooled [B,128] | legal-mask task list
    lws = torch.empty((L.aq_legal_mask_ws_bytes(B),), dtype=torch.uint8, device=dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)
    P = _lib.ptr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i):
        _lib.check(L.aq_leaf_eval(P(flat), P(prep), P(batches[i % nb]), B, P(priors), P(value), P(mask), P(pawn), P(pooled), prec, st),
                   "aq_leaf_eval")

    def timed(fn, n, warm):
        """n launches of fn(i), L2 flushed before each, per-launch CUDA events on the launch stream."""
        for i in range(warm):
            fn(i)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        barrier()
        for i in range(n):
            flush.fill_(i & 0xFF)
            evs[i][0].record()
            fn(i)
            evs[i][1].record()
        barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / n  # max over ranks of the mean ms per launch

    sampler = ClockSampler(local_rank)  # samples clocks / throttle reasons during ALL timed regions below
    sampler.start()
    ms_step = timed(step, K, Wm)
    value_main = world * B / (ms_step * 1e-3)

    # ---- per-kernel durations (same inputs, same stream) for the roofline ------------------------
    def k_legal(i):
        _lib.check(L.aq_legal_mask_ws(P(batches[i % nb]), B, P(mask), P(pawn), P(lws), lws.numel(), st), "aq_legal_mask_ws")

    def k_trunk(i):
        _lib.check(L.aq_gcn_trunk_forward(P(flat), P(prep), P(batches[i % nb]), B, P(pooled), prec, st), "aq_gcn_trunk_forward")

    def k_heads(i):
        _lib.check(L.aq_heads_forward(P(flat), P(prep), P(pooled), B, P(priors), P(value), P(mask), prec, st), "aq_heads_forward")

    kms = {"legal_mask_kernels": timed(k_legal, K, 2), "gcn_forward_kernel": timed(k_trunk, K, 2),
           "heads_forward_kernel": timed(k_heads, K, 2)}
    ksum = sum(kms.values())
    kinfo = {
        "legal_mask_kernels": {"bound": "hbm", "achieved": BYTES_PER_POSITION_LEGAL * B / (kms["legal_mask_kernels"] * 1e-3) / 1e9,
                              "peak": pk["hbm"], "unit": "GB/s"},
        "gcn_forward_kernel": {"bound": "tensor", "achieved": FLOP_PER_BOARD_FWD * B / (kms["gcn_forward_kernel"] * 1e-3) / 1e12,
                               "peak": pk["tensor"], "unit": "TFLOP/s"},
        "heads_forward_kernel": {"bound": "hbm", "achieved": BYTES_PER_BOARD_HEADS * B / (kms["heads_forward_kernel"] * 1e-3) / 1e9,
                                 "peak": pk["hbm"], "unit": "GB/s"},
    }
    for k, v in kinfo.items():
        v["ms"] = kms[k]
        v["share_of_step"] = kms[k] / ksum
        v["frac"] = v["achieved"] / v["peak"]
    dom = max(kms, key=kms.get)
    # DRAM traffic per launch of the dominant kernel from `ncu --set full` (profiles/): the trunk reads the packed
    # states and 256 KB of weights and writes pooled [B,128]; everything else stays in shared memory / TMEM.
    traffic = NCU_TRAFFIC_BYTES.get((dom, prec, B))
    roofline = {"kernel": dom, "bound": kinfo[dom]["bound"], "achieved": kinfo[dom]["achieved"], "peak": kinfo[dom]["peak"],
                "unit": kinfo[dom]["unit"], "frac": kinfo[dom]["frac"], "traffic": traffic, "peak_source": pk["source"],
                "arith": "fp32 FFMA" if prec == 0 else "bf16 tcgen05 node transforms + fp16 tcgen05 aggregation, fp32 accumulate in TMEM",
                "kernel_symbol": "gcn_forward_fp32_kernel" if prec == 0 else "gcn_forward_tc2_kernel"}

    # ---- end to end through host buffers, every step: pinned packed states H2D, kernels, results D2H, stream sync ----
    # Headline e2e = the repo's public host API (pv_network_gnn.HostLeafEvaluator -> aq_leaf_eval_host_compact): results in
    # the shape BaseNetwork.predict returns them (probabilities of the LEGAL actions only, legal_actions() order, ragged),
    # plus value / legal mask / pawn list.  The dense [B,209] flavour (aq_leaf_eval_host) is reported in extra.
    from alphaquoridorgnn_b200.pv_network_gnn import HostLeafEvaluator
    hst = [torch.from_numpy(gl.pack_rows_host(*[t.cpu().numpy() for t in gl.unpack_rows(b)])).pin_memory() for b in batches]
    for h, b in zip(hst, batches):
        assert torch.equal(h, b.cpu())

    def e2e_run(dense):
        ev = HostLeafEvaluator(net, B, dense=dense)
        d2h = []

        def e2e_step(i):
            d2h.append(ev.d2h_bytes(ev.evaluate(B, states=hst[i % nb])))  # synchronous: results are on the host when it returns

        for i in range(3):
            e2e_step(i)
        del d2h[:]
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(i)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ev.close()
        return world * B * K / float(dt.item()), int(sum(d2h) / len(d2h))

    def e2e_pipelined():
        """Two evaluators (two pools of games) alternate: while the host waits for one batch, the other batch's kernels run and
        the first one's results cross PCIe.  Every step still moves its own states H2D and its own results D2H."""
        evs = [HostLeafEvaluator(net, B), HostLeafEvaluator(net, B)]
        d2h = []

        def run(n):
            evs[0].submit(B, states=hst[0])
            for i in range(n):
                if i + 1 < n:
                    evs[(i + 1) & 1].submit(B, states=hst[(i + 1) % nb])
                out = evs[i & 1].wait()
                d2h.append(evs[i & 1].d2h_bytes(out))

        run(4)
        del d2h[:]
        barrier()
        t0 = time.perf_counter()
        run(K)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        for ev in evs:
            ev.close()
        return world * B * K / float(dt.item()), int(sum(d2h) / len(d2h))

    e2e_v, e2e_d2h = e2e_pipelined()
    e2e = {"value": e2e_v, "unit": "board-evals/s", "h2d_bytes_per_step": B * 32, "d2h_bytes_per_step": e2e_d2h,
           "api": "HostLeafEvaluator.submit / wait -> aq_leaf_eval_host_compact_submit / _wait: predict()-shaped ragged priors (legal "
                  "actions only) + value + legal mask + pawn list; two batches in flight (two evaluators used alternately)"}
    sync_v, sync_d2h = e2e_run(dense=False)
    dense_v, dense_d2h = e2e_run(dense=True)

    extra = {"e2e_one_batch_in_flight": {"value": sync_v, "unit": "board-evals/s", "d2h_bytes_per_step": sync_d2h,
                                         "api": "HostLeafEvaluator.evaluate (synchronous, one caller): aq_leaf_eval_host_compact"},
             "e2e_dense_priors": {"value": dense_v, "unit": "board-evals/s", "d2h_bytes_per_step": dense_d2h,
                                  "api": "aq_leaf_eval_host (synchronous): dense priors [B,209]"}}
    if not args.skip_extra:
        # legal mask, BASELINE configs[1]: 1M positions resident in HBM (32 MB in, 40 MB out > L2? no: flushed)
        M = 1_000_000
        big = positions.random_positions(M, seed=101 + rank, games=16384, device=dev)
        bmask = torch.empty((M, 8), dtype=torch.int32, device=dev)
        bpawn = torch.empty((M, 8), dtype=torch.uint8, device=dev)
        bws = torch.empty((L.aq_legal_mask_ws_bytes(M),), dtype=torch.uint8, device=dev)
        ms = timed(lambda i: _lib.check(L.aq_legal_mask_ws(P(big), M, P(bmask), P(bpawn), P(bws), bws.numel(), st), "aq_legal_mask_ws"), 5, 2)
        extra["legal_mask_positions_per_sec"] = world * M / (ms * 1e-3)
        extra["legal_mask_ms_per_1M"] = ms
        extra["legal_mask_hbm_frac"] = BYTES_PER_POSITION_LEGAL * M / (ms * 1e-3) / 1e9 / pk["hbm"]
        del big, bmask, bpawn, bws
        # training step (forward + loss + backward + gradient all-reduce + Adam), random targets:
        #   B=256  -- BASELINE configs[0] shape (what one optimizer step of the reference looks like)
        #   B=4096 -- the same step at a throughput-sized per-GPU batch
        def train_bench(TB):
            tb = allpos[:TB].contiguous()
            torch.manual_seed(1)
            pt = torch.softmax(torch.randn(TB, 209, device=dev), 1)
            vt = torch.randint(-1, 2, (TB,), device=dev).float()
            tflat = flat.clone()
            saved = torch.empty((L.aq_gnn_saved_floats(TB),), dtype=torch.float32, device=dev)
            bws = torch.empty((L.aq_gnn_backward_ws_floats(TB),), dtype=torch.float32, device=dev)
            tp = torch.empty((TB, 209), dtype=torch.float32, device=dev)
            tv = torch.empty((TB,), dtype=torch.float32, device=dev)
            dp, dv = torch.empty_like(tp), torch.empty_like(tv)
            grads, m1, m2 = torch.empty_like(tflat), torch.zeros_like(tflat), torch.zeros_like(tflat)
            loss = torch.zeros(2, device=dev)
            stepno = [0]
            tprec = prec  # training arithmetic follows --precision (bf16 = tcgen05 trunk forward/backward, fp32 accumulate)

            def train_step(i):
                stepno[0] += 1
                _lib.check(L.aq_gnn_forward(P(tflat), P(tb), None, None, TB, P(tp), P(tv), P(saved), tprec, st), "fwd")
                _lib.check(L.aq_loss_grad(P(tp), P(tv), P(pt), P(vt), TB, TB * world, P(loss), P(dp), P(dv), st), "loss")
                _lib.check(L.aq_gnn_backward(P(tflat), P(saved), P(dp), P(dv), TB, P(grads), P(bws), tprec, st), "bwd")
                if world > 1:
                    dist.all_reduce(grads)
                _lib.check(L.aq_adam_step(P(tflat), P(grads), P(m1), P(m2), tflat.numel(), stepno[0], 1e-3, 0.9, 0.999, 1e-8,
                                          1.0, st), "adam")

            return timed(train_step, 20, 5)

        ms = train_bench(256)
        extra["train_samples_per_sec"] = world * 256 / (ms * 1e-3)
        extra["train_ms_per_step_B256"] = ms
        ms = train_bench(4096)
        extra["train_samples_per_sec_B4096"] = world * 4096 / (ms * 1e-3)
        extra["train_ms_per_step_B4096"] = ms
        extra["train_tensor_frac_B4096"] = FLOP_PER_BOARD_FWDBWD * 4096 / (ms * 1e-3) / 1e12 / pk["tensor"]
        try:
            from alphaquoridorgnn_b200 import pv_mcts
            extra.update(pv_mcts.bench_sims_per_sec(net, dev, world, timed_barrier=barrier))
        except Exception as e:  # MCTS is a "next" row; absence must not break the headline
            extra["mcts"] = f"unavailable: {type(e).__name__}: {e}"

    clocks = sampler.stop()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_steps = 5
        v, sec, threads = time_cpu_port(B, cpu_steps, 1)
        cpu_baseline = {"value": v, "unit": "board-evals/s", "cores": threads, "kind": "port",
                        "sample": f"{cpu_steps} steps (1 warm-up) of {B} positions: C oracle legal_actions (OpenMP, {threads} threads) + "
                                  f"torch CPU GNN forward + legal renorm; {sec:.1f} s per step"}

    if rank == 0:
        line = {
            "metric": "gnn_board_evals_per_sec", "value": value_main, "unit": "board-evals/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if prec == 0 else "bf16", "data": "synthetic",
            "config": {"workload": "leaf_eval: BASELINE configs[2], batched predict (legal mask + graph + GNN forward + legal "
                                   "renorm) on random legal 9x9 positions, random-init weights",
                       "batch_per_gpu": B, "precision": args.precision, "parallelism": f"independent leaf batches x{world}",
                       "l2": "flushed before every timed launch (256 MiB write, outside the per-launch CUDA events)",
                       "host_numa": numa},
            "roofline": roofline, "kernels": kinfo,
            # the kernels of a step are chained by programmatic dependent launches (prologues overlap the predecessor's tail), so the
            # step is shorter than the sum of its kernels timed alone
            "sum_of_kernels_timed_alone_ms": ksum,
            "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": 4 * K,  # legal_prepare + legal_search + trunk + heads per step (plus one 4-byte memset)
            "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--precision", default=os.environ.get("AQ_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="GNN inference arithmetic: bf16 = tcgen05 tensor cores (default), fp32 = FFMA")
    ap.add_argument("--skip-extra", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
