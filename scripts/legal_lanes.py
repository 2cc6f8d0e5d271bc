"""Legal-mask time by batch size, mixed game phases: the one-kernel form (aq_legal_mask: 32 / 8 / 2 lanes per state by batch size)
against the two-phase form with a workspace (aq_legal_mask_ws, which itself takes the one-kernel form up to 4,096 states).
   python scripts/legal_lanes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import _lib, positions

L = _lib.load()
P = _lib.ptr
allpos, _ = positions.mixed_batches(1, 1 << 20, seed=1)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
mask = torch.empty((1 << 20, 8), dtype=torch.int32, device="cuda")
pawn = torch.empty((1 << 20, 8), dtype=torch.uint8, device="cuda")
st = _lib.stream_ptr()
lws = torch.empty((L.aq_legal_mask_ws_bytes(1 << 20),), dtype=torch.uint8, device="cuda")
print("B, aq_legal_mask (one kernel, lanes by batch size) us, aq_legal_mask_ws (two-phase above 4096 states) us")
for B in (256, 1024, 2048, 4096, 4097, 8192, 16384, 32768, 65536, 131072, 262144, 1 << 20):
    row = []
    for two_phase in (False, True):
        ts = []
        for it in range(8):
            x = allpos[(it * B) % ((1 << 20) - B + 1):][:B]
            flush.fill_(it)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if two_phase:
                _lib.check(L.aq_legal_mask_ws(P(x), B, P(mask), P(pawn), P(lws), lws.numel(), st), "aq_legal_mask_ws")
            else:
                _lib.check(L.aq_legal_mask(P(x), B, P(mask), P(pawn), st), "aq_legal_mask")
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        row.append(sorted(ts[2:])[len(ts[2:]) // 2])
    print(B, ", ".join(f"{t:.1f}" for t in row), flush=True)
