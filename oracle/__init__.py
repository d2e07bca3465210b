"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers of the strand-voxelisation path:

* ``port``  -- oracle/_build/libvoxel_oracle.so, the plain-C restatement in
  oracle/voxel_oracle.c (+ oracle/prefilter_oracle.c) of the reference's
  ``HairStyle::voxelize_segments`` / ``voxelize_vertices`` / ``Volume``
  (reference src/vkhr/scene_graph/hair_style.cc:215-357, hair_style.hh:228-257).
* ``ref``   -- oracle/_ref/libvkhr_ref.so, the UNMODIFIED reference
  hair_style.cc behind oracle/ref_driver.cc (prebuilt where /root/reference
  exists; travels to the GPU box as a binary).

Only tests/, tests/golden/make_golden.py, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  Nothing under vkhr_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libvoxel_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libvkhr_ref.so")
REFERENCE_ROOT = os.environ.get("VKHR_REFERENCE", "/root/reference")

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u16p = C.POINTER(C.c_uint16)
_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)


def build(ref: bool = True) -> None:
    """Compile the C restatement, and the reference driver when the reference tree is present."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REFERENCE={REFERENCE_ROOT}"])


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _vec3(a):
    return (C.c_float * 3)(*[float(x) for x in a])


class _Port:
    """The C restatement (cpu_baseline.kind == "port")."""

    def __init__(self):
        try:
            build(ref=False)              # incremental: a no-op when the .so is up to date
        except Exception:                 # noqa: BLE001  (no make/gcc: use the prebuilt library)
            if not os.path.exists(PORT_SO):
                raise
        L = self.lib = C.CDLL(PORT_SO)
        L.oracle_generate_indices.restype = C.c_uint64
        L.oracle_count_samples.restype = C.c_uint64
        L.oracle_fnv1a64.restype = C.c_uint64

    def generate_bounding_box(self, xyz):
        xyz = _f32(xyz)
        lo, hi = (C.c_float * 3)(), (C.c_float * 3)()
        self.lib.oracle_generate_bounding_box(_ptr(xyz, _f32p), C.c_uint64(xyz.size // 3), lo, hi)
        return np.array(lo, dtype=np.float32), np.array(hi, dtype=np.float32)

    def get_bounding_box(self, lo, hi):
        out = (C.c_float * 8)()
        self.lib.oracle_get_bounding_box(_vec3(lo), _vec3(hi), out)
        return np.array(out, dtype=np.float32)

    def generate_indices(self, n_strands, default_segments, segments=None):
        n_seg = int(np.sum(segments)) if segments is not None else n_strands * default_segments
        out = np.empty(2 * n_seg, dtype=np.uint32)
        seg = None if segments is None else np.ascontiguousarray(segments, dtype=np.uint16)
        n = self.lib.oracle_generate_indices(C.c_uint32(n_strands), C.c_uint32(default_segments),
                                             _ptr(seg, _u16p), _ptr(out, _u32p))
        assert n == out.size
        return out

    def generate_tangents(self, xyz, n_strands, default_segments, segments=None):
        xyz = _f32(xyz)
        out = np.zeros_like(xyz)
        seg = None if segments is None else np.ascontiguousarray(segments, dtype=np.uint16)
        self.lib.oracle_generate_tangents(_ptr(xyz, _f32p), C.c_uint32(n_strands),
                                          C.c_uint32(default_segments), _ptr(seg, _u16p), _ptr(out, _f32p))
        return out

    def voxelize_segments(self, xyz, indices, origin, size, W, H, D, tangents=None, flags=0):
        xyz = _f32(xyz)
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        dens = np.empty(W * H * D, dtype=np.uint8)
        tin = _f32(tangents)
        tout = np.empty((W * H * D, 4), dtype=np.int8) if tin is not None else None
        rc = self.lib.oracle_voxelize_segments(_ptr(xyz, _f32p), _ptr(idx, _u32p), C.c_uint64(idx.size),
                                               _ptr(tin, _f32p), _vec3(origin), _vec3(size),
                                               C.c_uint64(W), C.c_uint64(H), C.c_uint64(D), C.c_uint32(flags),
                                               _ptr(dens, _u8p), _ptr(tout, _i8p))
        if rc != 0:
            raise MemoryError("oracle_voxelize_segments")
        return (dens, tout) if tin is not None else dens

    def voxelize_vertices(self, xyz, origin, size, W, H, D, tangents=None, flags=0):
        xyz = _f32(xyz)
        dens = np.empty(W * H * D, dtype=np.uint8)
        tin = _f32(tangents)
        tout = np.empty((W * H * D, 4), dtype=np.int8) if tin is not None else None
        rc = self.lib.oracle_voxelize_vertices(_ptr(xyz, _f32p), C.c_uint64(xyz.size // 3),
                                               _ptr(tin, _f32p), _vec3(origin), _vec3(size),
                                               C.c_uint64(W), C.c_uint64(H), C.c_uint64(D), C.c_uint32(flags),
                                               _ptr(dens, _u8p), _ptr(tout, _i8p))
        if rc != 0:
            raise MemoryError("oracle_voxelize_vertices")
        return (dens, tout) if tin is not None else dens

    def count_segments(self, xyz, indices, origin, size, W, H, D, counts=None, flags=0):
        xyz = _f32(xyz)
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        if counts is None:
            counts = np.zeros(W * H * D, dtype=np.uint32)
        self.lib.oracle_count_segments(_ptr(xyz, _f32p), _ptr(idx, _u32p), C.c_uint64(idx.size),
                                       _vec3(origin), _vec3(size),
                                       C.c_uint64(W), C.c_uint64(H), C.c_uint64(D), C.c_uint32(flags), _ptr(counts, _u32p))
        return counts

    def count_vertices(self, xyz, origin, size, W, H, D, counts=None, flags=0):
        xyz = _f32(xyz)
        if counts is None:
            counts = np.zeros(W * H * D, dtype=np.uint32)
        self.lib.oracle_count_vertices(_ptr(xyz, _f32p), C.c_uint64(xyz.size // 3),
                                       _vec3(origin), _vec3(size),
                                       C.c_uint64(W), C.c_uint64(H), C.c_uint64(D), C.c_uint32(flags), _ptr(counts, _u32p))
        return counts

    def count_samples(self, xyz, indices, origin, size, W, H, D):
        xyz = _f32(xyz)
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        return int(self.lib.oracle_count_samples(_ptr(xyz, _f32p), _ptr(idx, _u32p), C.c_uint64(idx.size),
                                                 _vec3(origin), _vec3(size),
                                                 C.c_uint64(W), C.c_uint64(H), C.c_uint64(D)))

    def normalize(self, densities):
        d = np.array(densities, dtype=np.uint8, copy=True).reshape(-1)
        self.lib.oracle_normalize(_ptr(d, _u8p), C.c_uint64(d.size))
        return d

    def downsample(self, densities, W, H, D, filter=0):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty((W // 2) * (H // 2) * (D // 2), dtype=np.uint8)
        self.lib.oracle_downsample(_ptr(d, _u8p), C.c_uint32(W), C.c_uint32(H), C.c_uint32(D),
                                   C.c_int(filter), _ptr(out, _u8p))
        return out

    # ---- prefilter (oracle/prefilter_oracle.c: restatement of the GLSL consumers at voxel centres) ----
    def prefilter_ao(self, densities, W, H, D, radius=2.5, exponent=10.0, ao_max=0.16):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty(W * H * D, dtype=np.float32)
        self.lib.oracle_prefilter_ao(_ptr(d, _u8p), C.c_uint32(W), C.c_uint32(H), C.c_uint32(D),
                                     C.c_float(radius), C.c_float(exponent), C.c_float(ao_max), _ptr(out, _f32p))
        return out

    def prefilter_opacity(self, densities, strand_alpha=0.3, thickness=11.0):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty(d.size, dtype=np.float32)
        self.lib.oracle_prefilter_opacity(_ptr(d, _u8p), C.c_uint64(d.size), C.c_float(strand_alpha), C.c_float(thickness),
                                          _ptr(out, _f32p))
        return out

    def prefilter_gauss(self, densities, W, H, D, kernel_width=3.0):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty(W * H * D, dtype=np.float32)
        self.lib.oracle_prefilter_gauss(_ptr(d, _u8p), C.c_uint32(W), C.c_uint32(H), C.c_uint32(D), C.c_float(kernel_width),
                                        _ptr(out, _f32p))
        return out

    def adsm_t_table(self, steps=1024.0):
        n = int(self.lib.oracle_adsm_t_table(C.c_float(steps), None, C.c_uint32(0)))
        t = np.empty(n, dtype=np.float32)
        self.lib.oracle_adsm_t_table(C.c_float(steps), _ptr(t, _f32p), C.c_uint32(n))
        return t

    def prefilter_adsm(self, densities, W, H, D, origin, size, light, steps=1024.0, strand_alpha=0.3, thickness=11.0):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty(W * H * D, dtype=np.float32)
        o, s, l = (np.ascontiguousarray(a, dtype=np.float32) for a in (origin, size, light))
        self.lib.oracle_prefilter_adsm(_ptr(d, _u8p), C.c_uint32(W), C.c_uint32(H), C.c_uint32(D), _ptr(o, _f32p), _ptr(s, _f32p),
                                       _ptr(l, _f32p), C.c_float(steps), C.c_float(strand_alpha), C.c_float(thickness), _ptr(out, _f32p))
        return out

    def fnv1a64(self, buf):
        b = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        return int(self.lib.oracle_fnv1a64(_ptr(b, _u8p), C.c_uint64(b.size)))


class RefHairStyle:
    """A reference ``vkhr::HairStyle`` object living inside libvkhr_ref.so."""

    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def close(self):
        if self._h:
            self._lib.vkhr_ref_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def vertex_count(self):
        return int(self._lib.vkhr_ref_vertex_count(self._h))

    @property
    def strand_count(self):
        return int(self._lib.vkhr_ref_strand_count(self._h))

    @property
    def segment_count(self):
        return int(self._lib.vkhr_ref_segment_count(self._h))

    @property
    def default_segment_count(self):
        return int(self._lib.vkhr_ref_default_segment_count(self._h))

    @property
    def has_bounding_box(self):
        return bool(self._lib.vkhr_ref_has_bounding_box(self._h))

    @property
    def vertices(self):
        out = np.empty((self.vertex_count, 3), dtype=np.float32)
        self._lib.vkhr_ref_get_vertices(self._h, _ptr(out, _f32p))
        return out

    @property
    def tangents(self):
        out = np.empty((self.vertex_count, 3), dtype=np.float32)
        self._lib.vkhr_ref_get_tangents(self._h, _ptr(out, _f32p))
        return out

    @property
    def indices(self):
        out = np.empty(int(self._lib.vkhr_ref_index_count(self._h)), dtype=np.uint32)
        self._lib.vkhr_ref_get_indices(self._h, _ptr(out, _u32p))
        return out

    @property
    def segments(self):
        n = int(self._lib.vkhr_ref_get_segments(self._h, None))
        out = np.empty(n, dtype=np.uint16)
        if n:
            self._lib.vkhr_ref_get_segments(self._h, _ptr(out, _u16p))
        return out

    @property
    def aabb(self):
        """origin.xyz, radius, size.xyz, volume (struct AABB, hair_style.hh:16-21)."""
        out = (C.c_float * 8)()
        self._lib.vkhr_ref_get_aabb(self._h, out)
        return np.array(out, dtype=np.float32)

    def prepare(self):
        """SceneGraph::add_style's generate_* calls for the fields the file lacked (no shuffle)."""
        self._lib.vkhr_ref_prepare(self._h)
        return self

    def save(self, path):
        if self._lib.vkhr_ref_save(self._h, path.encode()) != 0:
            raise IOError(path)

    def voxelize(self, mode, W, H, D, normalize=False, want_tangents=False):
        """mode 'segments' | 'vertices'.  Returns (densities, tangents|None, seconds of the reference call)."""
        dens = np.empty(W * H * D, dtype=np.uint8)
        tang = np.empty((W * H * D, 4), dtype=np.int8) if want_tangents else None
        t = self._lib.vkhr_ref_voxelize(self._h, C.c_int(0 if mode == "segments" else 1),
                                        C.c_uint64(W), C.c_uint64(H), C.c_uint64(D),
                                        C.c_int(int(normalize)), _ptr(dens, _u8p), _ptr(tang, _i8p))
        if t < 0:
            raise ValueError("reference voxeliser refused the input (fewer than 2 indices)")
        return dens, tang, float(t)


class _Ref:
    """The unmodified reference (cpu_baseline.kind == "reference")."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            if os.path.isdir(REFERENCE_ROOT):
                build(ref=True)
            else:
                raise FileNotFoundError(f"{REF_SO} missing and no reference tree to build it from")
        L = self.lib = C.CDLL(REF_SO)
        for name in ("vkhr_ref_load", "vkhr_ref_create"):
            getattr(L, name).restype = C.c_void_p
        L.vkhr_ref_index_count.restype = C.c_uint64
        L.vkhr_ref_voxelize.restype = C.c_double
        L.vkhr_ref_voxelize.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _u8p, _i8p]
        L.vkhr_ref_prepare.argtypes = [C.c_void_p]
        L.vkhr_ref_prepare.restype = None
        for name in ("vkhr_ref_destroy", "vkhr_ref_vertex_count", "vkhr_ref_strand_count", "vkhr_ref_segment_count",
                     "vkhr_ref_index_count", "vkhr_ref_default_segment_count", "vkhr_ref_has_bounding_box"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.vkhr_ref_save.argtypes = [C.c_void_p, C.c_char_p]
        L.vkhr_ref_get_vertices.argtypes = [C.c_void_p, _f32p]
        L.vkhr_ref_get_tangents.argtypes = [C.c_void_p, _f32p]
        L.vkhr_ref_get_indices.argtypes = [C.c_void_p, _u32p]
        L.vkhr_ref_get_segments.argtypes = [C.c_void_p, _u16p]
        L.vkhr_ref_get_segments.restype = C.c_uint32
        L.vkhr_ref_get_aabb.argtypes = [C.c_void_p, _f32p]

    def create(self, xyz, n_strands, default_segments, segments=None, aabb_min=None, aabb_max=None):
        xyz = _f32(xyz)
        seg = None if segments is None else np.ascontiguousarray(segments, dtype=np.uint16)
        lo = None if aabb_min is None else _vec3(aabb_min)
        hi = None if aabb_max is None else _vec3(aabb_max)
        h = self.lib.vkhr_ref_create(_ptr(xyz, _f32p), C.c_uint32(xyz.size // 3), C.c_uint32(n_strands),
                                     C.c_uint32(default_segments), _ptr(seg, _u16p), lo, hi)
        if not h:
            raise RuntimeError("vkhr_ref_create failed")
        return RefHairStyle(self.lib, h)

    def volume_save(self, densities, path) -> bool:
        """``HairStyle::Volume::save`` of the unmodified reference (hair_style.cc:359-369)."""
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        self.lib.vkhr_ref_volume_save.argtypes = [_u8p, C.c_uint64, C.c_char_p]
        return bool(self.lib.vkhr_ref_volume_save(_ptr(d, _u8p), C.c_uint64(d.size), path.encode()))

    def load(self, path):
        h = self.lib.vkhr_ref_load(path.encode())
        if not h:
            raise IOError(f"reference HairStyle::load failed for {path}")
        return RefHairStyle(self.lib, h)

    def normalize(self, densities):
        d = np.array(densities, dtype=np.uint8, copy=True).reshape(-1)
        self.lib.vkhr_ref_normalize(_ptr(d, _u8p), C.c_uint64(d.size))
        return d

    def downsample(self, densities, W, H, D, filter=0):
        d = np.ascontiguousarray(densities, dtype=np.uint8).reshape(-1)
        out = np.empty((W // 2) * (H // 2) * (D // 2), dtype=np.uint8)
        self.lib.vkhr_ref_downsample(_ptr(d, _u8p), C.c_uint32(W), C.c_uint32(H), C.c_uint32(D),
                                     C.c_int(filter), _ptr(out, _u8p))
        return out


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref_available() -> bool:
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_ROOT)


def ref() -> _Ref:
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref
