"""Legal-mask launches bracketed by cudaProfilerStart/Stop for ncu (bench.py's workloads: B = 16,384 mixed game phases, and 1,000,000
positions):   ncu --profile-from-start off --set full -k regex:legal_ ... python scripts/profile_legal.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from alphaquoridorgnn_b200 import _lib, positions

L, P = _lib.load(), _lib.ptr
B = 16384
_, batches = positions.mixed_batches(4, B, seed=1)
big = positions.random_positions(1_000_000, seed=101, games=16384)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
st = _lib.stream_ptr()


def run(x):
    n = x.shape[0]
    mask = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    pawn = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
    ws = torch.empty((L.aq_legal_mask_ws_bytes(n),), dtype=torch.uint8, device="cuda")
    _lib.check(L.aq_legal_mask_ws(P(x), n, P(mask), P(pawn), P(ws), ws.numel(), st), "aq_legal_mask_ws")
    return mask


for i in range(2):
    run(batches[i])
run(big)
torch.cuda.synchronize()
torch.cuda.profiler.start()
flush.fill_(1)
run(batches[2])
flush.fill_(2)
run(batches[3])
flush.fill_(3)
run(big)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
