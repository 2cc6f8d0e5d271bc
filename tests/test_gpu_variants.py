"""The alternate tensor-core kernels stay correct: the kernel version is an environment knob read once per process
(AQ_TC_VERSION: inference trunk 1 / 2 / 3; AQ_TRAIN_TC_VERSION: training pair 1 / 2), so the bf16 tolerance tests of
tests/test_gpu_gnn.py are re-run in a subprocess for every non-default setting."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"AQ_TC_VERSION": "1"}, {"AQ_TC_VERSION": "3"}, {"AQ_TRAIN_TC_VERSION": "1"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_alternate_kernel_versions_within_tolerance(env):
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_gnn.py"), "-m", "gpu", "-x", "-q", "-k",
           "bf16_tensor_core or prepared_inference or empty_and_tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-1000:]


def test_tmem_resident_aggregation_operand_variant_within_tolerance():
    """gnn_tc2.cu built with -DTC2_AGG_TMEM=1 (Z^T as a tensor-memory A operand, 3 boards in flight; measured slower, kept as
    a documented alternative) passes the same bf16 tolerance tests."""
    lib = os.path.join(ROOT, "alphaquoridorgnn_b200", "variants", "libaqgnn_agg1.so")
    main = os.path.join(ROOT, "alphaquoridorgnn_b200", "libaqgnn.so")
    # the variant links the regular objects of every other source: rebuild it whenever the main library is newer (ABI additions)
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(main):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "build_variant.py"), "agg1", "gnn_tc2.cu", "-DTC2_AGG_TMEM=1"],
                              cwd=ROOT, timeout=900)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_gnn.py"), "-m", "gpu", "-x", "-q", "-k",
           "bf16_tensor_core or prepared_inference or empty_and_tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, AQ_LIB_PATH=lib))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-1000:]
