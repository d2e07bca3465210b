"""Host-side mirror of the reference's ``vkhr::HairStyle`` for the voxelisation path.

Same names, argument meaning and results as
``include/vkhr/scene_graph/hair_style.hh:16-21,23-210`` /
``src/vkhr/scene_graph/hair_style.cc`` so that parity tests read like calls on
the reference object.  The data-preparation helpers (``load``/``save``,
``generate_tangents/indices/bounding_box``) are host logic in numpy float32
(IEEE, one rounding per operation, like the reference's GLM code); the
voxelisers, ``Volume.normalize`` and ``Volume.downsample`` run on the GPU through the
C ABI -- there is no CPU implementation of them in this package.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Optional

import numpy as np

from . import capi
from .voxelizer import Voxelizer, default_voxelizer

_HEADER = struct.Struct("<4sIIIIff3f64s3f3f")      # 128 bytes, hair_style.hh:147-174
assert _HEADER.size == 128

# bitfield, hair_style.hh:152-162
HAS_SEGMENTS, HAS_VERTICES, HAS_THICKNESS, HAS_TRANSPARENCY = 1, 2, 4, 8
HAS_COLOR, HAS_TANGENTS, HAS_INDICES, HAS_BOUNDING_BOX = 16, 32, 64, 128


@dataclasses.dataclass
class AABB:
    """``struct AABB`` (hair_style.hh:16-21)."""
    origin: np.ndarray
    radius: float
    size: np.ndarray
    volume: float


@dataclasses.dataclass
class Volume:
    """``HairStyle::Volume`` (hair_style.hh:89-101): x-fastest u8 densities, i8vec4 tangents."""
    resolution: np.ndarray                 # 3 floats, like glm::vec3 resolution
    bounds: AABB
    densities: np.ndarray                  # uint8, W*H*D
    tangents: Optional[np.ndarray] = None  # int8, (W*H*D, 4)
    _vox: Optional[Voxelizer] = None

    @property
    def shape(self):
        return tuple(int(r) for r in self.resolution)

    def normalize(self) -> None:
        """``Volume::normalize`` (hair_style.cc:344-357), in place, on the GPU."""
        vox = self._vox or default_voxelizer()
        self.densities = vox.normalize(self.densities)

    def save(self, path: str) -> bool:
        """``Volume::save`` (hair_style.cc:359-369): raw dump of the densities."""
        d = np.ascontiguousarray(self.densities, dtype=np.uint8).reshape(-1)
        return capi.lib.vkhr_b200_volume_save(str(path).encode(), d.ctypes.data, d.size) == capi.OK

    def downsample(self, filter: int = capi.DOWNSAMPLE_MAX) -> "Volume":
        """``Volume::downsample`` (hair_style.hh:228-257) with one of the built-in 2x2x2 functors."""
        vox = self._vox or default_voxelizer()
        W, H, D = self.shape
        d = vox.downsample(self.densities, W, H, D, filter)
        return Volume(np.asarray(self.resolution, dtype=np.float32) / np.float32(2.0), self.bounds, d, None, vox)


class HairStyle:
    """Strand geometry + the voxelisers of the reference's ``HairStyle``."""

    def __init__(self, path: Optional[str] = None, voxelizer: Optional[Voxelizer] = None):
        self.segments = np.zeros(0, dtype=np.uint16)
        self.vertices = np.zeros((0, 3), dtype=np.float32)
        self.thickness = np.zeros(0, dtype=np.float32)
        self.transparency = np.zeros(0, dtype=np.float32)
        self.color = np.zeros((0, 3), dtype=np.float32)
        self.tangents = np.zeros((0, 3), dtype=np.float32)
        self.indices = np.zeros(0, dtype=np.uint32)
        self._strand_count = 0
        self._default_segment_count = 0
        self.default_thickness = 0.0
        self.default_transparency = 0.0
        self.default_color = np.zeros(3, dtype=np.float32)
        self.information = b""
        self._bbox_min = np.zeros(3, dtype=np.float32)
        self._bbox_max = np.zeros(3, dtype=np.float32)
        self._has_bounding_box = False
        self._vox = voxelizer
        if path is not None and not self.load(path):
            raise IOError(f"cannot load .hair file {path}")

    # ---- counts (hair_style.cc:72-92) ----------------------------------------
    def get_strand_count(self) -> int:
        return int(self.segments.size) if self.segments.size else int(self._strand_count)

    def set_strand_count(self, n: int) -> None:
        self._strand_count = int(n)

    def get_vertex_count(self) -> int:
        return int(self.vertices.shape[0])

    def get_segment_count(self) -> int:
        return self.get_vertex_count() - self.get_strand_count()

    def get_default_segment_count(self) -> int:
        return int(self._default_segment_count)

    def set_default_segment_count(self, n: int) -> None:
        self._default_segment_count = int(n)

    def has_segments(self): return self.segments.size != 0
    def has_vertices(self): return self.vertices.size != 0
    def has_tangents(self): return self.tangents.size != 0
    def has_indices(self): return self.indices.size != 0
    def has_bounding_box(self): return self._has_bounding_box

    def _uniform_segments(self) -> int:
        """segs per strand when every strand has the same count (else 0)."""
        if self.has_segments():
            s0 = int(self.segments[0])
            return s0 if np.all(self.segments == s0) else 0
        return self.get_default_segment_count()

    def _segment_counts(self) -> np.ndarray:
        if self.has_segments():
            return self.segments.astype(np.int64)
        return np.full(self.get_strand_count(), self.get_default_segment_count(), dtype=np.int64)

    # ---- .hair I/O (hair_style.cc:24-70; Cem Yuksel HAIR format) ---------------
    def load(self, path: str) -> bool:
        try:
            with open(path, "rb") as f:
                raw = f.read(128)
                if len(raw) != 128:
                    return False
                (sig, strands, nverts, bits, dsegs, dthick, dtransp, c0, c1, c2, info,
                 m0, m1, m2, x0, x1, x2) = _HEADER.unpack(raw)
                if sig != b"HAIR":
                    return False

                def rd(dtype, count):
                    a = np.fromfile(f, dtype=dtype, count=count)
                    if a.size != count:
                        raise EOFError
                    return a

                self.segments = rd(np.uint16, strands) if bits & HAS_SEGMENTS else np.zeros(0, np.uint16)
                self.vertices = (rd(np.float32, 3 * nverts).reshape(-1, 3) if bits & HAS_VERTICES
                                 else np.zeros((0, 3), np.float32))
                self.thickness = rd(np.float32, nverts) if bits & HAS_THICKNESS else np.zeros(0, np.float32)
                self.transparency = rd(np.float32, nverts) if bits & HAS_TRANSPARENCY else np.zeros(0, np.float32)
                self.color = (rd(np.float32, 3 * nverts).reshape(-1, 3) if bits & HAS_COLOR
                              else np.zeros((0, 3), np.float32))
                self.tangents = (rd(np.float32, 3 * nverts).reshape(-1, 3) if bits & HAS_TANGENTS
                                 else np.zeros((0, 3), np.float32))
                self._strand_count = strands
                self._default_segment_count = dsegs
                # read_indices sizes the array from get_segment_count() (hair_style.cc:621-626)
                self.indices = (rd(np.uint32, 2 * (nverts - self.get_strand_count())) if bits & HAS_INDICES
                                else np.zeros(0, np.uint32))
        except (OSError, EOFError):
            return False
        self.default_thickness, self.default_transparency = dthick, dtransp
        self.default_color = np.array([c0, c1, c2], dtype=np.float32)
        self.information = info.rstrip(b"\0")
        self._bbox_min = np.array([m0, m1, m2], dtype=np.float32)
        self._bbox_max = np.array([x0, x1, x2], dtype=np.float32)
        self._has_bounding_box = bool(bits & HAS_BOUNDING_BOX)
        return self.has_vertices()

    def save(self, path: str) -> bool:
        if not self.has_vertices():
            return False
        bits = ((HAS_SEGMENTS if self.has_segments() else 0) | HAS_VERTICES |
                (HAS_THICKNESS if self.thickness.size else 0) | (HAS_TRANSPARENCY if self.transparency.size else 0) |
                (HAS_COLOR if self.color.size else 0) | (HAS_TANGENTS if self.has_tangents() else 0) |
                (HAS_INDICES if self.has_indices() else 0) | (HAS_BOUNDING_BOX if self._has_bounding_box else 0))
        hdr = _HEADER.pack(b"HAIR", self.get_strand_count(), self.get_vertex_count(), bits,
                           self.get_default_segment_count(), self.default_thickness, self.default_transparency,
                           *[float(c) for c in self.default_color], bytes(self.information)[:64],
                           *[float(c) for c in self._bbox_min], *[float(c) for c in self._bbox_max])
        try:
            with open(path, "wb") as f:
                f.write(hdr)
                if self.has_segments(): f.write(self.segments.astype(np.uint16).tobytes())
                f.write(np.ascontiguousarray(self.vertices, np.float32).tobytes())
                if self.thickness.size: f.write(self.thickness.astype(np.float32).tobytes())
                if self.transparency.size: f.write(self.transparency.astype(np.float32).tobytes())
                if self.color.size: f.write(np.ascontiguousarray(self.color, np.float32).tobytes())
                if self.has_tangents(): f.write(np.ascontiguousarray(self.tangents, np.float32).tobytes())
                if self.has_indices(): f.write(self.indices.astype(np.uint32).tobytes())
        except OSError:
            return False
        return True

    # ---- generators (host logic) ---------------------------------------------------
    def generate_indices(self) -> None:
        """hair_style.cc:196-213: pairs (k, k+1) inside each strand."""
        counts = self._segment_counts()
        starts = np.concatenate(([0], np.cumsum(counts + 1)[:-1]))
        first = np.repeat(starts, counts) + (np.arange(int(counts.sum())) - np.repeat(np.cumsum(counts) - counts, counts))
        idx = np.empty(2 * first.size, dtype=np.uint32)
        idx[0::2] = first
        idx[1::2] = first + 1
        self.indices = idx

    def generate_tangents(self) -> None:
        """hair_style.cc:171-194: normalize(v[k+1]-v[k]); a strand's last vertex repeats the previous tangent."""
        v = np.ascontiguousarray(self.vertices, np.float32)
        counts = self._segment_counts()
        ends = np.cumsum(counts + 1) - 1                       # last vertex of each strand
        t = np.zeros_like(v)
        d = v[1:] - v[:-1]
        dot = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.float32(1.0) / np.sqrt(dot, dtype=np.float32)
            t[:-1] = d * inv[:, None]
        t[ends] = t[np.maximum(ends - 1, 0)]
        self.tangents = t

    def generate_thickness(self, radius: float = 0.042) -> None:
        """hair_style.cc:151-169: `radius` along the strand, 0 at the tip."""
        th = np.full(self.get_vertex_count(), radius, dtype=np.float32)
        th[np.cumsum(self._segment_counts() + 1) - 1] = 0.0
        self.thickness = th

    def generate_bounding_box(self) -> None:
        """hair_style.cc:215-234 on the GPU: min/max folded from (0,0,0) (the box always contains the origin)."""
        vox = self._vox or default_voxelizer()
        self._bbox_min, self._bbox_max = vox.generate_bounding_box(self.vertices)
        self._has_bounding_box = True

    def set_bounding_box(self, bbox_min, bbox_max) -> None:
        """What a .hair header with the has_bounding_box bit carries (hair_style.hh:172-173)."""
        self._bbox_min = np.asarray(bbox_min, dtype=np.float32).copy()
        self._bbox_max = np.asarray(bbox_max, dtype=np.float32).copy()
        self._has_bounding_box = True

    def get_bounding_box(self) -> AABB:
        """hair_style.cc:236-255."""
        size = (self._bbox_max - self._bbox_min).astype(np.float32)
        radius = np.sqrt((size[0] * size[0] + size[1] * size[1]) + size[2] * size[2], dtype=np.float32)
        return AABB(self._bbox_min.copy(), float(radius), size, float((size[0] * size[1]) * size[2]))

    # ---- the hot path -----------------------------------------------------------------
    def voxelize_segments(self, width: int, height: int, depth: int, flags: int = 0) -> Volume:
        """``HairStyle::voxelize_segments`` (hair_style.cc:296-342) on the GPU."""
        vox = self._vox or default_voxelizer()
        b = self.get_bounding_box()
        # the reference always fills Volume::tangents from this->tangents (hair_style.cc:309,:323)
        tin = self.tangents if self.has_tangents() and self.tangents.shape == self.vertices.shape else None
        if tin is not None:
            d, t = vox.voxelize_segments(self.vertices, self.indices, b.origin, b.size, width, height, depth, flags=flags, tangents=tin)
        else:
            d, t = vox.voxelize_segments(self.vertices, self.indices, b.origin, b.size, width, height, depth, flags=flags), None
        return Volume(np.array([width, height, depth], dtype=np.float32), b, d, t, vox)

    def voxelize_vertices(self, width: int, height: int, depth: int, flags: int = 0) -> Volume:
        """``HairStyle::voxelize_vertices`` (hair_style.cc:257-294) on the GPU."""
        vox = self._vox or default_voxelizer()
        b = self.get_bounding_box()
        tin = self.tangents if self.has_tangents() and self.tangents.shape == self.vertices.shape else None
        if tin is not None:
            d, t = vox.voxelize_vertices(self.vertices, b.origin, b.size, width, height, depth, flags=flags, tangents=tin)
        else:
            d, t = vox.voxelize_vertices(self.vertices, b.origin, b.size, width, height, depth, flags=flags), None
        return Volume(np.array([width, height, depth], dtype=np.float32), b, d, t, vox)
