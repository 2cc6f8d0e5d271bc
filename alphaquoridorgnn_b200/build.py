"""Builds libaqgnn.so (all CUDA kernels + the C ABI of include/aqgnn.h) in-tree with nvcc for
sm_100a.  nvcc cross-compiles without a GPU, so this runs in the build container; the .so then
travels to the GPU box with the repo snapshot."""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaqgnn.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def nvcc_path():
    p = find_nvcc()
    if p is None:
        raise RuntimeError("nvcc not found")
    return p


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


STAMP_PATH = LIB_PATH + ".stamp"


def source_digest():
    """sha256 over every file the library is built from (content, not mtime: the snapshot that carries the repo to the
    GPU box does not keep modification times)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(PKG_DIR, "..", "include", "aqgnn.h")]
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """True if libaqgnn.so is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc_path()] + flags + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                            env=dict(os.environ, CC="/usr/bin/gcc"))))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(obj_dir, "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see " + os.path.join(obj_dir, "nvcc.log"))
    link = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcuda"]
    subprocess.check_call(link)
    with open(STAMP_PATH, "w") as f:
        f.write(source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
