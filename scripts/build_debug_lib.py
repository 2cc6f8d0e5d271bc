"""Build a DEBUG copy of libaqgnn.so with extra -D defines for one source file (the per-phase clock64 accounting of the tensor-core
kernels: -DTC2_TIMING=1 in gnn_tc2.cu, -DTC2B_TIMING=1 in gnn_tc2_bwd.cu):
    python scripts/build_debug_lib.py timing gnn_tc2.cu -DTC2_TIMING=1      ->  alphaquoridorgnn_b200/debug/libaqgnn_timing.so
The timing scripts load it explicitly (scripts/tc2_timing.py <lib>, scripts/bwd_timing.py <lib>); the product always loads
alphaquoridorgnn_b200/libaqgnn.so.  All other objects come from the regular build."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alphaquoridorgnn_b200 import build as B  # noqa: E402

name, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
ddir = os.path.join(B.PKG_DIR, "debug")
os.makedirs(ddir, exist_ok=True)
obj = os.path.join(B.PKG_DIR, "build", f"{src[:-3]}_dbg_{name}.o")
flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
subprocess.check_call([B.nvcc_path()] + flags + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
regular = {os.path.basename(s)[:-3] + ".o" for s in B.sources()}
objs = [o for o in glob.glob(os.path.join(B.PKG_DIR, "build", "*.o")) if os.path.basename(o) in regular and os.path.basename(o) != src[:-3] + ".o"]
out = os.path.join(ddir, f"libaqgnn_{name}.so")
subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + [obj, "-lcuda"])
print(out)
