"""ctypes door to harness/_build/libvkhr_selftest.so (device self tests; GPU tests only)."""
from __future__ import annotations

import ctypes as C
import os

from . import SELFTEST_SO, build

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SELFTEST_SO):
            build(selftest=True)
        _lib = C.CDLL(SELFTEST_SO)
        _lib.vkhr_selftest_division.restype = C.c_int
        _lib.vkhr_selftest_division.argtypes = [C.c_int, C.c_float, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib.vkhr_selftest_rcp.restype = C.c_int
        _lib.vkhr_selftest_rcp.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(C.c_uint64)]
    return _lib


def division(divisor: float, n_trials: int = 1 << 26, seed: int = 1, device: int = 0) -> int:
    """Mismatches between the kernels' FMA division by ``divisor`` and the IEEE division (must be 0)."""
    bad = C.c_uint64(0)
    rc = lib().vkhr_selftest_division(int(device), float(divisor), int(n_trials), int(seed), C.byref(bad))
    if rc != 0:
        raise RuntimeError(f"vkhr_selftest_division failed ({rc})")
    return int(bad.value)


def rcp(lo: float, hi: float, device: int = 0) -> int:
    """Mismatches of the walk's reciprocal / `direction /= steps` division over EVERY float steps in [lo, hi] (must be 0)."""
    bad = C.c_uint64(0)
    rc = lib().vkhr_selftest_rcp(int(device), float(lo), float(hi), C.byref(bad))
    if rc != 0:
        raise RuntimeError(f"vkhr_selftest_rcp failed ({rc})")
    return int(bad.value)
