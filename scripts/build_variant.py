"""Build a variant of libaqgnn.so with extra -D defines for ONE source file (kernel tuning experiments):
    python scripts/build_variant.py NAME gnn_tc2.cu -DTC2_X=1 ...   ->  alphaquoridorgnn_b200/variants/libaqgnn_NAME.so
Select it at run time with AQ_LIB_PATH=<that path>.  All other objects come from the regular build."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alphaquoridorgnn_b200 import build as B  # noqa: E402

name, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.PKG_DIR, "variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(B.PKG_DIR, "build", f"{src[:-3]}_{name}.o")
flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
subprocess.check_call([B.nvcc_path()] + flags + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
objs = [o for o in glob.glob(os.path.join(B.PKG_DIR, "build", "*.o"))
        if "_" + name + ".o" not in o and os.path.basename(o) in {os.path.basename(s)[:-3] + ".o" for s in B.sources()} and os.path.basename(o) != src[:-3] + ".o"]
out = os.path.join(vdir, f"libaqgnn_{name}.so")
subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + [obj, "-lcuda"])
print(out)
