"""harness -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Host-side helpers the tests, bench.py and tests/golden/make_golden.py share:

* ``harness.synth``    -- the synthetic strand generator (harness/synth.cc, plain C++ + OpenMP, no CUDA):
  the reference's .hair assets are Git-LFS pointers, so every workload is generated.
  The same buffers feed the CPU oracle and the GPU path; nothing here influences parity.
* ``harness.selftest`` -- device self tests of the walk's numerics building blocks (harness/selftest.cu includes
  vkhr_b200/csrc/walk.cuh): the FMA division and the reciprocal against the IEEE instruction sequences, bit for bit.

Nothing under vkhr_b200/ imports this package, and it does not import vkhr_b200: bench.py's reference arm
uses it next to ``oracle`` without mapping the product library.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")
SYNTH_SO = os.path.join(BUILD_DIR, "libvkhr_harness.so")
SELFTEST_SO = os.path.join(BUILD_DIR, "libvkhr_selftest.so")
WALK_HEADER = os.path.join(_HERE, "..", "vkhr_b200", "csrc", "walk.cuh")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(selftest: bool = True) -> None:
    """Compile the host generator (g++) and, when nvcc is there, the device self-test library (sm_100a)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    src = os.path.join(_HERE, "synth.cc")
    if _stale(SYNTH_SO, [src]):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math", "-o", SYNTH_SO, src])
    cu = os.path.join(_HERE, "selftest.cu")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if selftest and os.path.exists(nvcc) and _stale(SELFTEST_SO, [cu, WALK_HEADER]):
        # the product's numerics flags (vkhr_b200/build.py): no FMA contraction, IEEE division, denormals kept
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                               "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                               "-Xcompiler", "-fPIC", "--shared", "-cudart", "static", "-o", SELFTEST_SO, cu])
