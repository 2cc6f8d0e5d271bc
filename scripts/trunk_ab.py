"""A/B of a library variant on the leaf-evaluation step (bench.py's workload and timing: CUDA events, L2 flushed between steps):
    python scripts/trunk_ab.py out.npz                                                   (the in-tree library)
    python scripts/with_lib.py alphaquoridorgnn_b200/debug/libaqgnn_x.so scripts/trunk_ab.py out_x.npz
    python scripts/trunk_ab.py --compare out.npz out_x.npz"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        print(f"{k}: max |a - b| = {np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max():.3e}")
    sys.exit(0)

import torch

from alphaquoridorgnn_b200 import _lib, positions
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork

torch.manual_seed(0)
net = GNNNetwork().cuda().eval()
net.precision = "bf16"
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
out = {}
for B in (16384, 4096, 256):
    _, batches = positions.mixed_batches(4, B, seed=1)
    for i in range(5):
        r = net.predict_batch(batches[i % 4])
    K = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.fill_(i)
        ev[i][0].record()
        r = net.predict_batch(batches[i % 4])
        ev[i][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    print(f"{os.path.basename(_lib.LIB_PATH)} B={B}: step median {t[K // 2]:.1f} us, min {t[0]:.1f} us")
    r = net.predict_batch(batches[0])
    out[f"priors{B}"] = r["priors"].float().cpu().numpy()
    out[f"value{B}"] = r["value"].float().cpu().numpy()
np.savez(sys.argv[1], **out)
