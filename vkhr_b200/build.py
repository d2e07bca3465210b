"""Build the in-tree native library ``vkhr_b200/lib/libvkhr_b200.so`` with nvcc for sm_100a.

The library is the product: CUDA kernels + the C ABI of include/vkhr_b200.h.
It is built in-tree (never JIT-cached) so that it travels with the repo snapshot.

Translation units (compiled in parallel, one object each, then linked):
  vkhr_b200.cu               walk / splat / combine / volume kernels, the frame kernel, the C ABI
  prefilter_variants.cu x 10 the tiled prefilter kernel, one (tap-offset pair) per object -- these
                             instantiations used to take 1 m 45 s in one translation unit
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libvkhr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact fp32: no FMA contraction, IEEE division/sqrt, denormals kept
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden,-fno-fast-math",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static"]

# (NO0, PO0) tap-offset pairs of the column-form prefilter; group 0 also carries the row-wise instantiation.
# Must match the declarations in vkhr_b200.cu (pf_variants_0 .. pf_variants_8).
PF_GROUPS = [(0, 0), (-1, 0), (-1, 1), (-2, 1), (-2, 2), (-3, 2), (-3, 3), (-4, 3), (-4, 4)]


def units():
    """(object name, source, extra defines)."""
    u = [("vkhr_b200.o", os.path.join(CSRC, "vkhr_b200.cu"), [])]
    for g, (n, p) in enumerate(PF_GROUPS):
        u.append((f"prefilter_variants_{g}.o", os.path.join(CSRC, "prefilter_variants.cu"),
                  [f"VKHR_PF_G={g}", f"VKHR_PF_NO0=({n})", f"VKHR_PF_PO0=({p})"]))
    return u


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", "vkhr_b200.h"), os.path.abspath(__file__)]
    return deps


def _stale(target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """``defines`` / ``out``: A/B builds of tuning macros into a side library (tools/), never the product path."""
    side = out is not None or bool(defines)
    target = out or LIB
    if not side and not force and not _stale(LIB):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libvkhr_b200.so")
    obj_dir = OBJ_DIR if not side else os.path.join(OBJ_DIR, "side_" + os.path.basename(target))
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(unit):
        name, src, defs = unit
        obj = os.path.join(obj_dir, name)
        # the prefilter objects do not depend on the tuning macros of an A/B build of the walk: reuse the product's
        if side and name.startswith("prefilter_variants_") and os.path.exists(os.path.join(OBJ_DIR, name)) \
                and not _stale(os.path.join(OBJ_DIR, name)) and not any(d.startswith("VKHR_B200_PF") for d in defines):
            return os.path.join(OBJ_DIR, name)
        if not force and not side and not _stale(obj):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in (*defs, *defines)], "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, 12)) as pool:
        objs = list(pool.map(compile_one, units()))
    subprocess.check_call([nvcc, *LINK_FLAGS, "-o", target, *objs])
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
