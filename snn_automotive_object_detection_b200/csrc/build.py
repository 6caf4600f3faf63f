"""Builds libsnn_heads_b200.so in-tree with nvcc for sm_100a (no torch types in the ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libsnn_heads_b200.so")
SOURCES = [os.path.join(HERE, "snn_abi.cu")]
HEADERS = [os.path.join(HERE, f) for f in ("ptx.cuh", "spike_gemm_lif.cuh", "aux_kernels.cuh")] + [
    os.path.join(os.path.dirname(PKG), "include", "snn_heads.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsnn_heads_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
