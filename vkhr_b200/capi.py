"""ctypes binding of ``include/vkhr_b200.h`` (the C ABI of libvkhr_b200.so).

This is the only door into the native code from Python; it declares exactly
the prototypes of the header.  The library is required: if it is missing and
cannot be built, importing this module raises -- there is no Python or CPU
fallback for the voxelisation path.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

LIB_PATH = _build.LIB
HEADER_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "vkhr_b200.h")

# status codes / flags (include/vkhr_b200.h)
OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_CUDA = -2
ERR_OUT_OF_MEMORY = -3
ERR_NO_DEVICE = -4
ERR_UNSUPPORTED = -5

INDEX_EXACT = 1 << 0
NORMALIZE = 1 << 1
STRATEGY_COUNT32 = 1 << 8
STRATEGY_PACKED8 = 1 << 9
STRATEGY_BRICK8 = 1 << 10
BRICK8_SPLIT = 1 << 11

DOWNSAMPLE_MAX, DOWNSAMPLE_MEAN, DOWNSAMPLE_SUM, DOWNSAMPLE_MIN = 0, 1, 2, 3

c_ctx = C.c_void_p
_f32p = C.POINTER(C.c_float)
_vec3 = C.c_float * 3


class Instance(C.Structure):
    """``vkhr_b200_instance``."""
    _fields_ = [
        ("d_vertices", C.c_void_p),
        ("d_indices", C.c_void_p),
        ("n_indices", C.c_uint64),
        ("n_vertices", C.c_uint32),
        ("segs_per_strand", C.c_uint32),
        ("aabb_origin", C.c_float * 3),
        ("aabb_size", C.c_float * 3),
        ("d_densities_out", C.c_void_p),
    ]


class HostInstance(C.Structure):
    """``vkhr_b200_host_instance``."""
    _fields_ = [
        ("vertices", C.c_void_p),
        ("indices", C.c_void_p),
        ("n_indices", C.c_uint64),
        ("n_vertices", C.c_uint32),
        ("segs_per_strand", C.c_uint32),
        ("aabb_origin", C.c_float * 3),
        ("aabb_size", C.c_float * 3),
        ("densities_out", C.c_void_p),
    ]


class ShardPeers(C.Structure):
    """``vkhr_b200_shard_peers``."""
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("partials", C.POINTER(C.c_void_p)), ("bitmaps", C.POINTER(C.c_void_p)),
                ("outs", C.POINTER(C.c_void_p)), ("signals", C.POINTER(C.c_void_p))]


class PrefilterParams(C.Structure):
    """``vkhr_b200_prefilter_params``."""
    _fields_ = [
        ("ao_radius", C.c_float), ("ao_exponent", C.c_float), ("ao_max", C.c_float),
        ("strand_alpha", C.c_float), ("thickness", C.c_float), ("gauss_width", C.c_float),
        ("flags", C.c_uint32),
    ]


PREFILTER_GENERIC = 1 << 0
PREFILTER_ROWWISE = 1 << 1
PREFILTER_DENSE = 1 << 2


class AdsmParams(C.Structure):
    """``vkhr_b200_adsm_params``."""
    _fields_ = [("light", C.c_float * 3), ("steps", C.c_float), ("strand_alpha", C.c_float), ("thickness", C.c_float)]


class VkhrB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"vkhr_b200 error {code}: {message}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH) or os.environ.get("VKHR_B200_REBUILD"):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise ImportError(
                f"libvkhr_b200.so is missing at {LIB_PATH} and could not be built ({e}); "
                "the voxelisation path has no fallback") from e
    return C.CDLL(LIB_PATH)


lib = _load()

_P = C.c_void_p
_u32, _u64, _int, _sz = C.c_uint32, C.c_uint64, C.c_int, C.c_size_t

_PROTOTYPES = {
    "vkhr_b200_create": (_int, [_int, C.POINTER(c_ctx)]),
    "vkhr_b200_destroy": (None, [c_ctx]),
    "vkhr_b200_last_error": (C.c_char_p, [c_ctx]),
    "vkhr_b200_version": (C.c_char_p, []),
    "vkhr_b200_stream": (_P, [c_ctx]),
    "vkhr_b200_synchronize": (_int, [c_ctx]),
    "vkhr_b200_launch_count": (_u64, [c_ctx]),
    "vkhr_b200_last_strategy": (C.c_uint32, [c_ctx]),
    "vkhr_b200_set_scratch_ring_bytes": (_int, [c_ctx, _sz]),
    "vkhr_b200_profile_enable": (_int, [c_ctx, _int]),
    "vkhr_b200_profile_read": (_int, [c_ctx, C.c_double * 4, C.c_uint32 * 4]),
    "vkhr_b200_voxelize_segments": (_int, [c_ctx, _P, _u32, _P, _u64, _u32, _P, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P]),
    "vkhr_b200_voxelize_vertices": (_int, [c_ctx, _P, _u32, _P, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P]),
    "vkhr_b200_voxelize_segments_dev": (_int, [c_ctx, _P, _u32, _P, _u64, _u32, _P, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P, _P]),
    "vkhr_b200_voxelize_vertices_dev": (_int, [c_ctx, _P, _u32, _P, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P, _P]),
    "vkhr_b200_voxelize_segments_batch_dev": (_int, [c_ctx, C.POINTER(Instance), _u32, _u32, _u32, _u32, _u32, _P]),
    "vkhr_b200_voxelize_segments_batch": (_int, [c_ctx, C.POINTER(HostInstance), _u32, _u32, _u32, _u32, _u32]),
    "vkhr_b200_host_register": (_int, [c_ctx, _P, _sz]),
    "vkhr_b200_host_unregister": (_int, [c_ctx, _P]),
    "vkhr_b200_count_segments_dev": (_int, [c_ctx, _P, _u32, _P, _u64, _u32, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P]),
    "vkhr_b200_count_vertices_dev": (_int, [c_ctx, _P, _u32, _vec3, _vec3, _u32, _u32, _u32, _u32, _P, _P]),
    "vkhr_b200_saturating_sum_u8_dev": (_int, [c_ctx, _P, _u32, _u64, _P, _P]),
    "vkhr_b200_combine_peer_u8_dev": (_int, [c_ctx, _P, _P, _u32, _u64, _u64, _P]),
    "vkhr_b200_chunk_bitmap_dev": (_int, [c_ctx, _P, _u64, _P, _P]),
    "vkhr_b200_combine_peer_u8_sparse_dev": (_int, [c_ctx, _P, _P, _P, _u32, _u64, _u64, _P]),
    "vkhr_b200_sharded_volume_bytes": (_u64, [_u32, _u32, _u32, _u32]),
    "vkhr_b200_voxelize_segments_sharded_dev": (_int, [c_ctx, _P, _u32, _P, _u64, _u32, _vec3, _vec3, _u32, _u32, _u32, _u32,
                                                 C.POINTER(ShardPeers), _P]),
    "vkhr_b200_clamp_counts_dev": (_int, [c_ctx, _P, _u64, _u32, _P, _P]),
    "vkhr_b200_normalize_dev": (_int, [c_ctx, _P, _u64, _P]),
    "vkhr_b200_normalize": (_int, [c_ctx, _P, _u64]),
    "vkhr_b200_downsample_dev": (_int, [c_ctx, _P, _u32, _u32, _u32, _int, _P, _P]),
    "vkhr_b200_downsample": (_int, [c_ctx, _P, _u32, _u32, _u32, _int, _P]),
    "vkhr_b200_generate_bounding_box_dev": (_int, [c_ctx, _P, _u32, _P, _P]),
    "vkhr_b200_generate_bounding_box": (_int, [c_ctx, _P, _u32, C.c_float * 6]),
    "vkhr_b200_profile_read_ex": (_int, [c_ctx, C.POINTER(C.c_double), C.POINTER(C.c_uint32), _u32]),
    "vkhr_b200_volume_save": (_int, [C.c_char_p, _P, _u64]),
    "vkhr_b200_prefilter_defaults": (None, [C.POINTER(PrefilterParams)]),
    "vkhr_b200_prefilter_dev": (_int, [c_ctx, _P, _u32, _u32, _u32, C.POINTER(PrefilterParams), _P, _P, _P, _P]),
    "vkhr_b200_prefilter": (_int, [c_ctx, _P, _u32, _u32, _u32, C.POINTER(PrefilterParams), _P, _P, _P]),
    "vkhr_b200_voxelize_hair": (_int, [c_ctx, _P, _sz, _u32, _u32, _u32, _u32, _P, _P, C.c_float * 6]),
    "vkhr_b200_adsm_dev": (_int, [c_ctx, _P, _u32, _u32, _u32, _vec3, _vec3, C.POINTER(AdsmParams), _P, _P]),
    "vkhr_b200_adsm": (_int, [c_ctx, _P, _u32, _u32, _u32, _vec3, _vec3, C.POINTER(AdsmParams), _P]),
    "vkhr_b200_malloc": (_int, [c_ctx, _sz, C.POINTER(_P)]),
    "vkhr_b200_free": (_int, [c_ctx, _P]),
    "vkhr_b200_memset": (_int, [c_ctx, _P, _int, _sz, _P]),
    "vkhr_b200_upload": (_int, [c_ctx, _P, _P, _sz, _P]),
    "vkhr_b200_download": (_int, [c_ctx, _P, _P, _sz, _P]),
}

for _name, (_res, _args) in _PROTOTYPES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def header_symbols() -> list[str]:
    """Every function include/vkhr_b200.h declares (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"VKHR_B200_API\s+[\w\s\*]+?\b(vkhr_b200_\w+)\s*\(", text)))


def vec3(a) -> C.Array:
    return _vec3(float(a[0]), float(a[1]), float(a[2]))


def last_error(ctx) -> str:
    msg = lib.vkhr_b200_last_error(ctx)
    return msg.decode() if msg else ""


def check(ctx, rc: int) -> None:
    if rc != OK:
        raise VkhrB200Error(rc, last_error(ctx))
