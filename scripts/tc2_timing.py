"""Per-phase cycle accounting of the tc2 trunk (debug variant built with -DTC2_TIMING=1):
python scripts/build_debug_lib.py timing <source>.cu -D...=1;  python scripts/tc2_timing.py alphaquoridorgnn_b200/debug/libaqgnn_timing.so"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphaquoridorgnn_b200 import _lib, positions
import alphaquoridorgnn_b200.build as _b
_lib.LIB_PATH = os.path.abspath(sys.argv[1])   # the debug copy, loaded explicitly
_b.needs_build = lambda: False
from alphaquoridorgnn_b200.pv_network_gnn import GNNNetwork
B = 16384
net = GNNNetwork().cuda().eval(); net.precision = "bf16"
_, (pos,) = positions.mixed_batches(1, B, seed=1)
L = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_longlong * 16)()
net.predict_batch(pos); torch.cuda.synchronize()
L.aq_debug_tc2_timing(out)
net.predict_batch(pos); torch.cuda.synchronize()
L.aq_debug_tc2_timing(out)
names3 = {3: "epilogue/pool + sync (x3)", 4: "L1 MMA wait", 5: "(unused)", 6: "MMA issue by thread 0 (x2)", 7: "node phase of next board", 8: "transform+aggregate wait (x2)", 9: "pool exchange + store"}
names = {2: "inputs + sync", 3: "node work + sync", 4: "L1 MMA wait", 5: "epilogue X (+sync)", 6: "transform wait", 7: "epilogue Z (+sync)",
         8: "aggregate wait", 9: "pool / store", 10: "final sync"}
groups = 4
boards = (B // 148 + groups - 1) // groups
tot = sum(out[i] for i in range(16))
for i in sorted(names):
    print(f"{names[i]:32s} {out[i]/boards:9.0f} cycles/board  {100*out[i]/tot:5.1f}%")
print("total per board", tot / boards)
