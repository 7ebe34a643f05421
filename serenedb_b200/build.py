"""Builds libsdbg.so (the sm_100a kernels + C ABI) in-tree with nvcc. No JIT cache: the .so lives
under serenedb_b200/_lib/ so it travels to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsdbg.so")
SOURCES = [os.path.join(CSRC, "sdbg_abi.cu"), os.path.join(CSRC, "posting_format.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("bm25_kernels.cuh", "column_kernels.cuh", "device_common.cuh",
                                                   "posting_format.hpp")] + [
    os.path.join(os.path.dirname(HERE), "include", "sdbg.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "static"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a. Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # GPU box without sources changed: use the shipped build
        raise RuntimeError("nvcc not found and no prebuilt libsdbg.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
