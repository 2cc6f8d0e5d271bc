"""Two-phase legal mask (aq_legal_mask_ws) at 16 K / 128 K / 1 M states for a library variant (scripts/with_lib.py):
registers-per-thread / resident-CTA trade-off of legal_prepare_kernel and legal_search_kernel (-DAQ_PREP_MIN_CTAS, -DAQ_SEARCH_MIN_CTAS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import _lib, positions

L, P = _lib.load(), _lib.ptr
allpos, _ = positions.mixed_batches(1, 1 << 20, seed=1)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
mask = torch.empty((1 << 20, 8), dtype=torch.int32, device="cuda")
pawn = torch.empty((1 << 20, 8), dtype=torch.uint8, device="cuda")
st = _lib.stream_ptr()
lws = torch.empty((L.aq_legal_mask_ws_bytes(1 << 20),), dtype=torch.uint8, device="cuda")
out = []
for B in (16384, 131072, 1 << 20):
    ts = []
    for it in range(12):
        x = allpos[(it * B) % ((1 << 20) - B + 1):][:B]
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.aq_legal_mask_ws(P(x), B, P(mask), P(pawn), P(lws), lws.numel(), st), "aq_legal_mask_ws")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    out.append(f"{B}: {sorted(ts[2:])[5]:.1f} us")
print(os.path.basename(_lib.LIB_PATH), " | ".join(out), "| checksum", int(mask.sum().item()), flush=True)
