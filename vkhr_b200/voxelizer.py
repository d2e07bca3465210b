"""``Voxelizer`` -- one native context (one GPU + one stream) behind the C ABI.

Host-pointer methods take numpy arrays and are synchronous; ``*_dev`` methods
take torch CUDA tensors (torch is used for device memory and streams only),
enqueue on torch's current stream and do not synchronise.  All compute happens
in libvkhr_b200.so; nothing here touches voxels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import lib


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Voxelizer:
    def __init__(self, device: int = 0):
        h = capi.c_ctx()
        rc = lib.vkhr_b200_create(int(device), C.byref(h))
        if rc != capi.OK:
            raise capi.VkhrB200Error(rc, capi.last_error(None))
        self._h = h
        self.device = int(device)

    # ---- lifetime ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.vkhr_b200_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self) -> int:
        return int(lib.vkhr_b200_launch_count(self._h))

    @property
    def last_strategy(self) -> int:
        """Strategy flag (capi.STRATEGY_*) the last voxelisation ran with."""
        return int(lib.vkhr_b200_last_strategy(self._h))

    @property
    def stream(self) -> int:
        return int(lib.vkhr_b200_stream(self._h) or 0)

    def set_scratch_ring_bytes(self, n_bytes: int) -> None:
        """BRICK8 scratch the frame kernel keeps in flight (an L2-resident ring of volumes); a tuning knob."""
        capi.check(self._h, lib.vkhr_b200_set_scratch_ring_bytes(self._h, int(n_bytes)))

    def profile_enable(self, on: bool = True) -> None:
        capi.check(self._h, lib.vkhr_b200_profile_enable(self._h, int(bool(on))))

    def profile_read(self) -> dict:
        """Summed device milliseconds / span counts per phase since the last read."""
        ms, n = (C.c_double * 5)(), (C.c_uint32 * 5)()
        capi.check(self._h, lib.vkhr_b200_profile_read_ex(self._h, ms, n, 5))
        names = ("clear", "walk", "finish", "normalize", "prefilter")
        return {k: {"ms": ms[i], "spans": int(n[i])} for i, k in enumerate(names)}

    def synchronize(self) -> None:
        capi.check(self._h, lib.vkhr_b200_synchronize(self._h))

    @staticmethod
    def version() -> str:
        return lib.vkhr_b200_version().decode()

    # ---- host-pointer API (numpy) ------------------------------------------
    def voxelize_segments(self, vertices, indices, aabb_origin, aabb_size, W, H, D,
                          segs_per_strand: int = 0, flags: int = 0, tangents=None, want_tangents: bool = False,
                          out=None, tangents_out=None):
        """Replaces ``HairStyle::voxelize_segments`` (reference hair_style.cc:296-342); returns W*H*D uint8.

        ``want_tangents`` (or ``tangents`` given): also returns the int8 (W*H*D, 4) tangent volume
        (``Volume::tangents``); without ``tangents`` the tangent of a segment is normalize(tip - root).
        ``out`` / ``tangents_out``: caller-owned result arrays (e.g. views of pinned memory) instead of fresh ones.
        """
        v = _np(vertices, np.float32).reshape(-1, 3)
        idx = None if indices is None else _np(indices, np.uint32).reshape(-1)
        tin = None if tangents is None else _np(tangents, np.float32).reshape(-1, 3)
        if tin is not None and tin.shape != v.shape:
            raise ValueError("tangents must have one row per vertex")
        want = want_tangents or tin is not None or tangents_out is not None
        nvox = int(W) * int(H) * int(D)
        if out is None:
            out = np.empty(nvox, dtype=np.uint8)
        elif out.dtype != np.uint8 or out.size != nvox or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint8 array of W*H*D")
        tout = tangents_out if tangents_out is not None else (np.empty((nvox, 4), dtype=np.int8) if want else None)
        if tout is not None and (tout.dtype != np.int8 or tout.size != 4 * nvox or not tout.flags.c_contiguous):
            raise ValueError("tangents_out must be a contiguous int8 array of W*H*D x 4")
        rc = lib.vkhr_b200_voxelize_segments(self._h, _p(v), v.shape[0], _p(idx),
                                             0 if idx is None else idx.size, int(segs_per_strand), _p(tin),
                                             capi.vec3(aabb_origin), capi.vec3(aabb_size),
                                             int(W), int(H), int(D), int(flags), _p(out), _p(tout))
        capi.check(self._h, rc)
        return (out, tout) if want else out

    def voxelize_vertices(self, vertices, aabb_origin, aabb_size, W, H, D, flags: int = 0, tangents=None):
        """Replaces ``HairStyle::voxelize_vertices`` (reference hair_style.cc:257-294); with ``tangents`` also the tangent volume."""
        v = _np(vertices, np.float32).reshape(-1, 3)
        tin = None if tangents is None else _np(tangents, np.float32).reshape(-1, 3)
        if tin is not None and tin.shape != v.shape:
            raise ValueError("tangents must have one row per vertex")
        out = np.empty(int(W) * int(H) * int(D), dtype=np.uint8)
        tout = np.empty((out.size, 4), dtype=np.int8) if tin is not None else None
        rc = lib.vkhr_b200_voxelize_vertices(self._h, _p(v), v.shape[0], _p(tin),
                                             capi.vec3(aabb_origin), capi.vec3(aabb_size),
                                             int(W), int(H), int(D), int(flags), _p(out), _p(tout))
        capi.check(self._h, rc)
        return (out, tout) if tin is not None else out

    def make_host_batch(self, instances: Sequence[dict]):
        """``vkhr_b200_host_instance[]`` from dicts of numpy arrays: vertices (f32), out (u8, W*H*D),
        aabb_origin, aabb_size and either indices (u32) or segs_per_strand."""
        arr = (capi.HostInstance * len(instances))()
        keep = []
        for k, ins in enumerate(instances):
            v, o, idx = ins["vertices"], ins["out"], ins.get("indices")
            if v.dtype != np.float32 or o.dtype != np.uint8 or not v.flags.c_contiguous or not o.flags.c_contiguous:
                raise ValueError("vertices must be contiguous float32 and out contiguous uint8")
            if idx is not None and (idx.dtype != np.uint32 or not idx.flags.c_contiguous):
                raise ValueError("indices must be contiguous uint32")
            arr[k].vertices = v.ctypes.data
            arr[k].indices = None if idx is None else idx.ctypes.data
            arr[k].n_indices = 0 if idx is None else idx.size
            arr[k].n_vertices = v.size // 3
            arr[k].segs_per_strand = int(ins.get("segs_per_strand", 0))
            arr[k].aabb_origin = capi.vec3(ins["aabb_origin"])
            arr[k].aabb_size = capi.vec3(ins["aabb_size"])
            arr[k].densities_out = o.ctypes.data
            keep.append((v, o, idx))
        return arr, keep

    def voxelize_segments_batch(self, batch, W, H, D, flags: int = 0) -> None:
        """``voxelize_segments`` of a crowd held in host memory (pipelined upload / kernels / download)."""
        if isinstance(batch, (list, tuple)) and batch and isinstance(batch[0], dict):
            batch = self.make_host_batch(batch)
        arr = batch[0] if isinstance(batch, tuple) else batch
        if isinstance(batch, tuple):
            for v, o, _ in batch[1]:
                if o.size != int(W) * int(H) * int(D):
                    raise ValueError(f"out holds {o.size} bytes, the call needs {int(W) * int(H) * int(D)}")
        capi.check(self._h, lib.vkhr_b200_voxelize_segments_batch(self._h, arr, len(arr), int(W), int(H), int(D), int(flags)))

    def host_register(self, array: np.ndarray) -> None:
        capi.check(self._h, lib.vkhr_b200_host_register(self._h, C.c_void_p(array.ctypes.data), array.nbytes))

    def host_unregister(self, array: np.ndarray) -> None:
        capi.check(self._h, lib.vkhr_b200_host_unregister(self._h, C.c_void_p(array.ctypes.data)))

    def normalize(self, densities) -> np.ndarray:
        """``Volume::normalize`` (reference hair_style.cc:344-357) on a host grid; returns a new array."""
        d = np.array(densities, dtype=np.uint8, copy=True).reshape(-1)
        capi.check(self._h, lib.vkhr_b200_normalize(self._h, _p(d), d.size))
        return d

    def downsample(self, densities, W, H, D, filter: int = capi.DOWNSAMPLE_MAX) -> np.ndarray:
        """``Volume::downsample`` (reference hair_style.hh:228-257)."""
        d = _np(densities, np.uint8).reshape(-1)
        if d.size != W * H * D:
            raise ValueError("densities does not match W*H*D")
        out = np.empty((W // 2) * (H // 2) * (D // 2), dtype=np.uint8)
        capi.check(self._h, lib.vkhr_b200_downsample(self._h, _p(d), int(W), int(H), int(D), int(filter), _p(out)))
        return out

    @staticmethod
    def prefilter_params(**kw) -> "capi.PrefilterParams":
        """``vkhr_b200_prefilter_params`` with the reference's defaults (interface.hh:101-105), overridden by ``kw``."""
        p = capi.PrefilterParams()
        lib.vkhr_b200_prefilter_defaults(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown prefilter parameter {k}")
            setattr(p, k, v)
        return p

    def prefilter(self, densities, W, H, D, ao=True, opacity=False, gauss=False, **params) -> dict:
        """Density -> AO / opacity / Gaussian float volumes (the shaders' local_ambient_occlusion,
        volume_approximated_deep_shadows step and filter_volume at every voxel centre), host arrays."""
        d = _np(densities, np.uint8).reshape(-1)
        if d.size != W * H * D:
            raise ValueError("densities does not match W*H*D")
        p = self.prefilter_params(**params)
        out = {k: np.empty(d.size, dtype=np.float32) for k, on in (("ao", ao), ("opacity", opacity), ("gauss", gauss)) if on}
        capi.check(self._h, lib.vkhr_b200_prefilter(self._h, _p(d), int(W), int(H), int(D), C.byref(p),
                                                    _p(out.get("ao")), _p(out.get("opacity")), _p(out.get("gauss"))))
        return out

    def voxelize_hair(self, hair_bytes, W, H, D, flags: int = 0, want_tangents: bool = False):
        """``.hair`` file image -> (densities, tangents | None, aabb_origin, aabb_size): HairStyle::load +
        SceneGraph::add_style's missing-field generation + voxelize_segments over get_bounding_box()."""
        buf = np.frombuffer(bytes(hair_bytes), dtype=np.uint8)
        dens = np.empty(int(W) * int(H) * int(D), dtype=np.uint8)
        tang = np.empty((dens.size, 4), dtype=np.int8) if want_tangents else None
        box = (C.c_float * 6)()
        capi.check(self._h, lib.vkhr_b200_voxelize_hair(self._h, _p(buf), buf.size, int(W), int(H), int(D), int(flags),
                                                        _p(dens), _p(tang), box))
        b = np.array(box, dtype=np.float32)
        return dens, tang, b[:3].copy(), b[3:].copy()

    @staticmethod
    def adsm_params(light, steps=1024.0, strand_alpha=0.3, thickness=11.0) -> "capi.AdsmParams":
        """``vkhr_b200_adsm_params``: lights[0].origin + raycast_steps / hair_alpha / thickness (volume.frag:72-78)."""
        p = capi.AdsmParams()
        p.light = (C.c_float * 3)(*[float(x) for x in light])
        p.steps, p.strand_alpha, p.thickness = float(steps), float(strand_alpha), float(thickness)
        return p

    def adsm(self, densities, W, H, D, aabb_origin, aabb_size, light, **params) -> np.ndarray:
        """Volumetric ADSM transmittance volume (``volume_approximated_deep_shadows`` at every voxel centre), host arrays."""
        d = _np(densities, np.uint8).reshape(-1)
        if d.size != W * H * D:
            raise ValueError("densities does not match W*H*D")
        out = np.empty(d.size, dtype=np.float32)
        p = self.adsm_params(light, **params)
        capi.check(self._h, lib.vkhr_b200_adsm(self._h, _p(d), int(W), int(H), int(D), capi.vec3(aabb_origin), capi.vec3(aabb_size),
                                               C.byref(p), _p(out)))
        return out

    def generate_bounding_box(self, vertices) -> tuple[np.ndarray, np.ndarray]:
        """``HairStyle::generate_bounding_box`` (reference hair_style.cc:215-234): (min, max) folded from (0,0,0)."""
        v = _np(vertices, np.float32).reshape(-1, 3)
        out = (C.c_float * 6)()
        capi.check(self._h, lib.vkhr_b200_generate_bounding_box(self._h, _p(v), v.shape[0], out))
        a = np.array(out, dtype=np.float32)
        return a[:3].copy(), a[3:].copy()

    # ---- device-pointer API (torch CUDA tensors) ----------------------------
    def _torch_stream(self, stream):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        # handle 0 is torch's legacy default stream; the C ABI reads NULL as "the context's own stream",
        # so name the legacy stream explicitly (cudaStreamLegacy == 1)
        return C.c_void_p(int(stream) or 1)

    def _check_dev(self, t, dtype, name):
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"{name} must be a CUDA tensor on cuda:{self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous {dtype}")

    def _check_size(self, t, n, name):
        if t.numel() != int(n):
            raise ValueError(f"{name} holds {t.numel()} elements, the call needs {int(n)}")

    def voxelize_segments_dev(self, vertices, indices, aabb_origin, aabb_size, W, H, D,
                              segs_per_strand: int = 0, flags: int = 0, out=None, stream=None, tangents_out=None):
        """Device-resident ``HairStyle::voxelize_segments``; ``tangents_out`` (int8, W*H*D*4) also asks for
        ``Volume::tangents`` of normalize(tip - root) per segment (the reference's default mode, hair_style.cc:309-327)."""
        import torch
        self._check_dev(vertices, torch.float32, "vertices")
        if indices is not None:
            self._check_dev(indices, torch.int32, "indices")
        n = int(W) * int(H) * int(D)
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=vertices.device)
        self._check_dev(out, torch.uint8, "out")
        self._check_size(out, n, "out")
        if tangents_out is not None:
            self._check_dev(tangents_out, torch.int8, "tangents_out")
            self._check_size(tangents_out, 4 * n, "tangents_out")
        rc = lib.vkhr_b200_voxelize_segments_dev(
            self._h, C.c_void_p(vertices.data_ptr()), vertices.numel() // 3,
            None if indices is None else C.c_void_p(indices.data_ptr()),
            0 if indices is None else indices.numel(), int(segs_per_strand), None,
            capi.vec3(aabb_origin), capi.vec3(aabb_size), int(W), int(H), int(D), int(flags),
            C.c_void_p(out.data_ptr()), None if tangents_out is None else C.c_void_p(tangents_out.data_ptr()),
            self._torch_stream(stream))
        capi.check(self._h, rc)
        return out

    def voxelize_vertices_dev(self, vertices, aabb_origin, aabb_size, W, H, D, flags: int = 0, out=None, stream=None):
        import torch
        self._check_dev(vertices, torch.float32, "vertices")
        n = int(W) * int(H) * int(D)
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=vertices.device)
        self._check_dev(out, torch.uint8, "out")
        self._check_size(out, n, "out")
        rc = lib.vkhr_b200_voxelize_vertices_dev(
            self._h, C.c_void_p(vertices.data_ptr()), vertices.numel() // 3, None,
            capi.vec3(aabb_origin), capi.vec3(aabb_size), int(W), int(H), int(D), int(flags),
            C.c_void_p(out.data_ptr()), None, self._torch_stream(stream))
        capi.check(self._h, rc)
        return out

    def make_batch(self, instances: Sequence[dict]):
        """Build the ``vkhr_b200_instance[]`` array once (per-frame calls then reuse it).

        Each dict: vertices (cuda f32), out (cuda u8), aabb_origin, aabb_size,
        and either indices (cuda i32) or segs_per_strand.
        """
        import torch
        arr = (capi.Instance * len(instances))()
        keep = []
        for k, ins in enumerate(instances):
            v, o = ins["vertices"], ins["out"]
            self._check_dev(v, torch.float32, "vertices")
            self._check_dev(o, torch.uint8, "out")
            idx = ins.get("indices")
            if idx is not None:
                self._check_dev(idx, torch.int32, "indices")
            arr[k].d_vertices = v.data_ptr()
            arr[k].d_indices = None if idx is None else idx.data_ptr()
            arr[k].n_indices = 0 if idx is None else idx.numel()
            arr[k].n_vertices = v.numel() // 3
            arr[k].segs_per_strand = int(ins.get("segs_per_strand", 0))
            arr[k].aabb_origin = capi.vec3(ins["aabb_origin"])
            arr[k].aabb_size = capi.vec3(ins["aabb_size"])
            arr[k].d_densities_out = o.data_ptr()
            keep.append((v, o, idx))
        return arr, keep

    def voxelize_segments_batch_dev(self, batch, W, H, D, flags: int = 0, stream=None) -> None:
        """``voxelize_segments`` for a crowd: ``batch`` from :meth:`make_batch` (or a list of dicts)."""
        if isinstance(batch, (list, tuple)) and batch and isinstance(batch[0], dict):
            batch = self.make_batch(batch)
        arr = batch[0] if isinstance(batch, tuple) else batch
        if isinstance(batch, tuple):
            for v, o, _ in batch[1]:
                self._check_size(o, int(W) * int(H) * int(D), "out")
        rc = lib.vkhr_b200_voxelize_segments_batch_dev(self._h, arr, len(arr), int(W), int(H), int(D),
                                                       int(flags), self._torch_stream(stream))
        capi.check(self._h, rc)

    def count_segments_dev(self, vertices, indices, aabb_origin, aabb_size, W, H, D, counts,
                           segs_per_strand: int = 0, flags: int = 0, stream=None):
        """ADD this shard's hits into ``counts`` (cuda int32, W*H*D) -- the multi-GPU partial."""
        import torch
        self._check_dev(vertices, torch.float32, "vertices")
        self._check_dev(counts, torch.int32, "counts")
        if indices is not None:
            self._check_dev(indices, torch.int32, "indices")
        if counts.numel() != int(W) * int(H) * int(D):
            raise ValueError("counts does not match W*H*D")
        rc = lib.vkhr_b200_count_segments_dev(
            self._h, C.c_void_p(vertices.data_ptr()), vertices.numel() // 3,
            None if indices is None else C.c_void_p(indices.data_ptr()),
            0 if indices is None else indices.numel(), int(segs_per_strand),
            capi.vec3(aabb_origin), capi.vec3(aabb_size), int(W), int(H), int(D), int(flags),
            C.c_void_p(counts.data_ptr()), self._torch_stream(stream))
        capi.check(self._h, rc)
        return counts

    def count_vertices_dev(self, vertices, aabb_origin, aabb_size, W, H, D, counts, flags: int = 0, stream=None):
        import torch
        self._check_dev(vertices, torch.float32, "vertices")
        self._check_dev(counts, torch.int32, "counts")
        self._check_size(counts, int(W) * int(H) * int(D), "counts")
        rc = lib.vkhr_b200_count_vertices_dev(
            self._h, C.c_void_p(vertices.data_ptr()), vertices.numel() // 3,
            capi.vec3(aabb_origin), capi.vec3(aabb_size), int(W), int(H), int(D), int(flags),
            C.c_void_p(counts.data_ptr()), self._torch_stream(stream))
        capi.check(self._h, rc)
        return counts

    def clamp_counts_dev(self, counts, flags: int = 0, out=None, stream=None):
        """densities = min(counts, 255) [+ normalize]."""
        import torch
        self._check_dev(counts, torch.int32, "counts")
        if out is None:
            out = torch.empty(counts.numel(), dtype=torch.uint8, device=counts.device)
        self._check_dev(out, torch.uint8, "out")
        self._check_size(out, counts.numel(), "out")
        rc = lib.vkhr_b200_clamp_counts_dev(self._h, C.c_void_p(counts.data_ptr()), counts.numel(), int(flags),
                                            C.c_void_p(out.data_ptr()), self._torch_stream(stream))
        capi.check(self._h, rc)
        return out

    def saturating_sum_u8_dev(self, slabs, out=None, stream=None):
        """out[i] = min(sum_r slabs[r, i], 255): the combine of saturated u8 partial volumes (multi-GPU)."""
        import torch
        self._check_dev(slabs, torch.uint8, "slabs")
        if slabs.dim() != 2:
            raise ValueError("slabs must be (n_slabs, slab_bytes)")
        if out is None:
            out = torch.empty(slabs.shape[1], dtype=torch.uint8, device=slabs.device)
        self._check_dev(out, torch.uint8, "out")
        capi.check(self._h, lib.vkhr_b200_saturating_sum_u8_dev(self._h, C.c_void_p(slabs.data_ptr()), int(slabs.shape[0]),
                                                                int(slabs.shape[1]), C.c_void_p(out.data_ptr()),
                                                                self._torch_stream(stream)))
        return out

    def combine_peer_u8_dev(self, partial_ptrs, out_ptrs, slab_offset: int, slab_bytes: int, stream=None):
        """Fused peer-memory combine: ``partial_ptrs`` / ``out_ptrs`` are the device addresses of every rank's partial /
        output volume as mapped in this process (e.g. ``_SymmetricMemory.buffer_ptrs``)."""
        n = len(partial_ptrs)
        if n != len(out_ptrs):
            raise ValueError("one output pointer per partial pointer")
        pa = (C.c_void_p * n)(*[int(p) for p in partial_ptrs])
        oa = (C.c_void_p * n)(*[int(p) for p in out_ptrs])
        capi.check(self._h, lib.vkhr_b200_combine_peer_u8_dev(self._h, pa, oa, n, int(slab_offset), int(slab_bytes),
                                                              self._torch_stream(stream)))

    def chunk_bitmap_dev(self, volume, bitmap, stream=None):
        """One bit per 16-byte chunk of ``volume`` (cuda uint8, a multiple of 512 bytes) into ``bitmap`` (cuda int32)."""
        import torch
        self._check_dev(volume, torch.uint8, "volume")
        self._check_dev(bitmap, torch.int32, "bitmap")
        if bitmap.numel() * 512 < volume.numel():
            raise ValueError("bitmap too small")
        capi.check(self._h, lib.vkhr_b200_chunk_bitmap_dev(self._h, C.c_void_p(volume.data_ptr()), volume.numel(),
                                                           C.c_void_p(bitmap.data_ptr()), self._torch_stream(stream)))

    def combine_peer_u8_sparse_dev(self, partial_ptrs, bitmap_ptrs, out_ptrs, slab_offset: int, slab_bytes: int, stream=None):
        """Sparse fused peer-memory combine (outputs zeroed beforehand; see include/vkhr_b200.h)."""
        n = len(partial_ptrs)
        if n != len(out_ptrs) or n != len(bitmap_ptrs):
            raise ValueError("one bitmap and one output pointer per partial pointer")
        pa = (C.c_void_p * n)(*[int(p) for p in partial_ptrs])
        ba = (C.c_void_p * n)(*[int(p) for p in bitmap_ptrs])
        oa = (C.c_void_p * n)(*[int(p) for p in out_ptrs])
        capi.check(self._h, lib.vkhr_b200_combine_peer_u8_sparse_dev(self._h, pa, ba, oa, n, int(slab_offset), int(slab_bytes),
                                                                     self._torch_stream(stream)))

    @staticmethod
    def sharded_volume_bytes(W, H, D, world) -> int:
        return int(lib.vkhr_b200_sharded_volume_bytes(int(W), int(H), int(D), int(world)))

    def voxelize_segments_sharded_dev(self, vertices, indices, aabb_origin, aabb_size, W, H, D, rank, partial_ptrs, bitmap_ptrs,
                                      out_ptrs, signal_ptrs, segs_per_strand: int = 0, flags: int = 0, stream=None) -> None:
        """This rank's part of a strand-sharded ``voxelize_segments`` (``vkhr_b200_voxelize_segments_sharded_dev``):
        ``*_ptrs`` are the device addresses of every rank's partial / bitmap / output / signal buffers as mapped into this
        process.  On return (stream-ordered) the complete volume is in ``out_ptrs[rank]``."""
        import torch
        n_v = 0
        vp = None
        if vertices is not None and vertices.numel():
            self._check_dev(vertices, torch.float32, "vertices")
            n_v, vp = vertices.numel() // 3, C.c_void_p(vertices.data_ptr())
        if indices is not None:
            self._check_dev(indices, torch.int32, "indices")
        world = len(partial_ptrs)
        if not (world == len(bitmap_ptrs) == len(out_ptrs) == len(signal_ptrs)):
            raise ValueError("one partial, bitmap, output and signal pointer per rank")
        arr = lambda ptrs: (C.c_void_p * world)(*[int(p) for p in ptrs])   # noqa: E731
        pa, ba, oa, sa = arr(partial_ptrs), arr(bitmap_ptrs), arr(out_ptrs), arr(signal_ptrs)
        peers = capi.ShardPeers(int(rank), world, pa, ba, oa, sa)
        capi.check(self._h, lib.vkhr_b200_voxelize_segments_sharded_dev(
            self._h, vp, n_v, None if indices is None else C.c_void_p(indices.data_ptr()), 0 if indices is None else indices.numel(),
            int(segs_per_strand), capi.vec3(aabb_origin), capi.vec3(aabb_size), int(W), int(H), int(D), int(flags),
            C.byref(peers), self._torch_stream(stream)))

    def normalize_dev(self, densities, stream=None):
        import torch
        self._check_dev(densities, torch.uint8, "densities")
        capi.check(self._h, lib.vkhr_b200_normalize_dev(self._h, C.c_void_p(densities.data_ptr()),
                                                        densities.numel(), self._torch_stream(stream)))
        return densities

    def downsample_dev(self, densities, W, H, D, filter: int = capi.DOWNSAMPLE_MAX, out=None, stream=None):
        import torch
        self._check_dev(densities, torch.uint8, "densities")
        self._check_size(densities, int(W) * int(H) * int(D), "densities")
        if out is None:
            out = torch.empty((W // 2) * (H // 2) * (D // 2), dtype=torch.uint8, device=densities.device)
        self._check_dev(out, torch.uint8, "out")
        self._check_size(out, (W // 2) * (H // 2) * (D // 2), "out")
        capi.check(self._h, lib.vkhr_b200_downsample_dev(self._h, C.c_void_p(densities.data_ptr()), int(W), int(H), int(D),
                                                         int(filter), C.c_void_p(out.data_ptr()), self._torch_stream(stream)))
        return out

    def prefilter_dev(self, densities, W, H, D, ao=None, opacity=None, gauss=None, stream=None, **params):
        """Device-resident prefilter: ``ao`` / ``opacity`` / ``gauss`` are cuda float32 tensors of W*H*D (or None)."""
        import torch
        self._check_dev(densities, torch.uint8, "densities")
        if densities.numel() != int(W) * int(H) * int(D):
            raise ValueError("densities does not match W*H*D")
        for name, t in (("ao", ao), ("opacity", opacity), ("gauss", gauss)):
            if t is not None:
                self._check_dev(t, torch.float32, name)
                if t.numel() != densities.numel():
                    raise ValueError(f"{name} does not match W*H*D")
        p = self.prefilter_params(**params)
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        capi.check(self._h, lib.vkhr_b200_prefilter_dev(self._h, C.c_void_p(densities.data_ptr()), int(W), int(H), int(D),
                                                        C.byref(p), ptr(ao), ptr(opacity), ptr(gauss), self._torch_stream(stream)))

    def adsm_dev(self, densities, W, H, D, aabb_origin, aabb_size, light, out=None, stream=None, **params):
        """Device-resident ADSM transmittance volume: ``out`` is a cuda float32 tensor of W*H*D (allocated if None)."""
        import torch
        self._check_dev(densities, torch.uint8, "densities")
        if densities.numel() != int(W) * int(H) * int(D):
            raise ValueError("densities does not match W*H*D")
        if out is None:
            out = torch.empty(densities.numel(), dtype=torch.float32, device=densities.device)
        self._check_dev(out, torch.float32, "out")
        p = self.adsm_params(light, **params)
        capi.check(self._h, lib.vkhr_b200_adsm_dev(self._h, C.c_void_p(densities.data_ptr()), int(W), int(H), int(D),
                                                   capi.vec3(aabb_origin), capi.vec3(aabb_size), C.byref(p),
                                                   C.c_void_p(out.data_ptr()), self._torch_stream(stream)))
        return out

    def generate_bounding_box_dev(self, vertices, out=None, stream=None):
        import torch
        self._check_dev(vertices, torch.float32, "vertices")
        if out is None:
            out = torch.empty(6, dtype=torch.float32, device=vertices.device)
        capi.check(self._h, lib.vkhr_b200_generate_bounding_box_dev(self._h, C.c_void_p(vertices.data_ptr()),
                                                                    vertices.numel() // 3, C.c_void_p(out.data_ptr()),
                                                                    self._torch_stream(stream)))
        return out


_default: dict[int, Voxelizer] = {}


def default_voxelizer(device: int = 0) -> Voxelizer:
    """Process-wide context per device (what ``HairStyle.voxelize_*`` uses)."""
    v = _default.get(device)
    if v is None or v.handle is None:
        v = _default[device] = Voxelizer(device)
    return v
