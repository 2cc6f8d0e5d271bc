"""Run a script of this repo against a DEBUG copy of the library (scripts/build_debug_lib.py):
    python scripts/with_lib.py alphaquoridorgnn_b200/debug/libaqgnn_x.so scripts/train_variants.py [args]"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphaquoridorgnn_b200 import _lib
import alphaquoridorgnn_b200.build as _b

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
_b.needs_build = lambda: False
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
