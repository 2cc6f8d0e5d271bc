"""Where does the end-to-end leaf-evaluation step spend its time?  PCIe copy bandwidth at the step's payload
sizes, kernel time per chunk size, and the host-buffer call for several chunk counts (run on the GPU box)."""
import ctypes
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphaquoridorgnn_b200 import _lib, positions  # noqa: E402
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    L, P = _lib.load(), _lib.ptr
    if len(sys.argv) > 1 and sys.argv[1] == "host":
        B = 16384
        net = GNNNetwork().to(dev).eval()
        net.precision = "bf16"
        flat, prep = net.flat_parameters(), net.prepared_weights()
        pos = positions.random_positions(B, seed=1, games=8192, device=dev)
        hst = pos.cpu().pin_memory()
        h_pri = torch.empty((B, 209), dtype=torch.float32).pin_memory()
        h_val = torch.empty((B,), dtype=torch.float32).pin_memory()
        h_msk = torch.empty((B, 8), dtype=torch.int32).pin_memory()
        h_pwn = torch.empty((B, 8), dtype=torch.uint8).pin_memory()
        ws = torch.empty((L.aq_leaf_eval_host_ws_bytes(B),), dtype=torch.uint8, device=dev)
        hctx = ctypes.c_void_p()
        _lib.check(L.aq_host_ctx_create(ctypes.byref(hctx)), "ctx")
        st = _lib.stream_ptr(dev)
        for i in range(25):
            if i == 5:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(L.aq_leaf_eval_host(P(flat), P(prep), P(hst), B, P(h_pri), P(h_val), P(h_msk), P(h_pwn), P(ws), 1, hctx, st), "host")
        dt = (time.perf_counter() - t0) / 20
        print(f"chunks={os.environ.get('AQ_HOST_CHUNKS')} graph={os.environ.get('AQ_HOST_GRAPH')}: {dt*1e3:.3f} ms/step  {B/dt/1e6:.1f} M/s")
        return
    for mb in (0.5, 3.6, 7.2, 14.4, 64):
        n = int(mb * 1e6)
        d = torch.empty((n,), dtype=torch.uint8, device=dev)
        h = torch.empty((n,), dtype=torch.uint8).pin_memory()
        for direction in ("d2h", "h2d"):
            for _ in range(3):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"{direction} {mb:5.1f} MB: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
    net = GNNNetwork().to(dev).eval()
    net.precision = "bf16"
    flat, prep = net.flat_parameters(), net.prepared_weights()
    st = _lib.stream_ptr(dev)
    for B in (1024, 2048, 4096, 8192, 16384):
        pos = positions.random_positions(B, seed=1, games=8192, device=dev)
        pri = torch.empty((B, 209), device=dev); val = torch.empty((B,), device=dev)
        msk = torch.empty((B, 8), dtype=torch.int32, device=dev); pwn = torch.empty((B, 8), dtype=torch.uint8, device=dev)
        pooled = torch.empty((L.aq_leaf_eval_ws_floats(B),), device=dev)  # leaf-eval workspace (pooled | legal-mask task list)
        for _ in range(3):
            L.aq_leaf_eval(P(flat), P(prep), P(pos), B, P(pri), P(val), P(msk), P(pwn), P(pooled), 1, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            L.aq_leaf_eval(P(flat), P(prep), P(pos), B, P(pri), P(val), P(msk), P(pwn), P(pooled), 1, st)
        b.record(); torch.cuda.synchronize()
        print(f"leaf_eval B={B:6d}: {a.elapsed_time(b)/20*1e3:8.1f} us (warm L2)")
    for env in ({"AQ_HOST_GRAPH": "0", "AQ_HOST_CHUNKS": "1"}, {"AQ_HOST_CHUNKS": "1"}, {"AQ_HOST_CHUNKS": "2"}, {"AQ_HOST_CHUNKS": "3"},
                {"AQ_HOST_CHUNKS": "4"}, {"AQ_HOST_CHUNKS": "6"}):
        subprocess.run([sys.executable, os.path.abspath(__file__), "host"], env=dict(os.environ, **env))


if __name__ == "__main__":
    main()
