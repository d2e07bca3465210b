"""Build the in-tree native library ``vkhr_b200/lib/libvkhr_b200.so`` with nvcc for sm_100a.

The library is the product: CUDA kernels + the C ABI of include/vkhr_b200.h.
It is built in-tree (never JIT-cached) so that it travels with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libvkhr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact fp32: no FMA contraction, IEEE division/sqrt, denormals kept
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden,-fno-fast-math,-fopenmp",
    "--shared", "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cc")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", "vkhr_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """``defines`` / ``out``: A/B builds of tuning macros into a side library (tools/), never the product path."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libvkhr_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out or LIB, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
