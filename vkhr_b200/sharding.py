"""Strand sharding across the GPUs of one box (BASELINE.json configs[2]): one process per GPU.

Each rank voxelises a contiguous range of whole strands into a partial u32 hit-count grid; the
partials are summed with ONE integer collective over NVLink and clamped:
``density = min(sum_r count_r, 255)``.  That equals the single-GPU (and the reference's sequential)
result for every partition because the reference's counter only saturates
(``if (d != 255) d += 1``, hair_style.cc:277-280, :322-325).

Three exchange schedules:

* ``"allreduce"``  -- ``all_reduce(u32 grid, SUM)`` then a local clamp: 2(n-1)/n * 4 N^3 bytes per GPU.
* ``"rs_ag"``      -- ``reduce_scatter(u32)`` -> the owner clamps its slab -> ``all_gather(u8)``:
  (n-1)/n * (4 + 1) N^3 bytes per GPU, 37.5 % less NVLink traffic, and the clamp touches 1/n of the grid.
* ``"u8"``         -- every rank voxelises its shard straight into a SATURATED u8 partial (the single-GPU PACKED8
  path, no u32 grid at all); ``all_to_all(u8 slabs)`` -> the owner adds its n partial slabs with a byte-wise
  saturating add -> ``all_gather(u8)``: 2 (n-1)/n * N^3 bytes per GPU, 4x less than the all-reduce.  Exact because
  ``min(sum_r min(c_r, 255), 255) == min(sum_r c_r, 255)``.
* ``"p2p"``        -- the same arithmetic as ``"u8"`` with the exchange fused into ONE kernel over NVLink peer memory
  (``vkhr_b200_combine_peer_u8_dev``): partial and output volumes live in torch symmetric memory, every GPU reads its
  slab of all partials straight from the peers' HBM and stores the finished slab into all outputs; two device-side
  barriers of the symmetric-memory group bracket the kernel.  The exchange is sparse: one bit per 16-byte chunk of each
  partial tells the owner which peers to read, and only non-zero results are stored (outputs zeroed beforehand), so
  NVLink carries the hair and not the empty space.  CUDA + NVLink only (no gloo form).

torch.distributed is the plumbing (NCCL on the GPUs; gloo in the CPU tests of this module's host
logic).  Counting and clamping run in libvkhr_b200.so; ``count_fn`` / ``clamp_fn`` exist so that the
world_size-2 gloo tests can drive the exchange logic with the CPU oracle standing in for the kernels --
nothing in the package passes them.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def strand_range(n_strands: int, world: int, rank: int) -> Tuple[int, int]:
    """(first strand, strand count) of ``rank``: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_strands), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_vertices(vertices: np.ndarray, n_strands: int, segs_per_strand: int, world: int, rank: int) -> np.ndarray:
    """This rank's rows of a strand-major (n_strands * (segs + 1), 3) vertex array (a view, whole strands)."""
    v = np.asarray(vertices).reshape(-1, 3)
    vps = segs_per_strand + 1
    if v.shape[0] != n_strands * vps:
        raise ValueError("vertices do not match n_strands * (segs_per_strand + 1)")
    first, count = strand_range(n_strands, world, rank)
    return v[first * vps:(first + count) * vps]


def padded_voxels(n_voxels: int, world: int) -> int:
    """Grid length rounded up so that every rank owns an equal, 16-byte aligned slab (reduce-scatter layout)."""
    q = 16 * world
    return (n_voxels + q - 1) // q * q


class ShardedVoxelizer:
    """voxelize_segments / voxelize_vertices over the ranks of a torch.distributed group."""

    def __init__(self, voxelizer=None, group=None, count_fn: Optional[Callable] = None,
                 clamp_fn: Optional[Callable] = None, partial_fn: Optional[Callable] = None,
                 satsum_fn: Optional[Callable] = None):
        import torch.distributed as dist
        self.dist = dist
        self.vox = voxelizer
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._count = count_fn or self._count_cuda
        self._clamp = clamp_fn or self._clamp_cuda
        self._partial = partial_fn or self._partial_cuda
        self._satsum = satsum_fn or self._satsum_cuda
        self._counts = None

    # ---- the CUDA path (the only one the package uses) -------------------------------------------
    def _count_cuda(self, mode, vertices, indices, segs, origin, size, W, H, D, counts, flags):
        if self.vox is None:
            raise RuntimeError("ShardedVoxelizer needs a Voxelizer: there is no CPU fallback")
        if mode == "segments":
            self.vox.count_segments_dev(vertices, indices, origin, size, W, H, D, counts[:W * H * D],
                                        segs_per_strand=segs, flags=flags)
        else:
            self.vox.count_vertices_dev(vertices, origin, size, W, H, D, counts[:W * H * D], flags=flags)

    def _clamp_cuda(self, counts, out, flags=0):
        self.vox.clamp_counts_dev(counts, flags=flags, out=out)

    def _partial_cuda(self, mode, vertices, indices, segs, origin, size, W, H, D, out, flags):
        """This shard's saturated u8 volume (the ordinary single-GPU voxelisation of the shard)."""
        if self.vox is None:
            raise RuntimeError("ShardedVoxelizer needs a Voxelizer: there is no CPU fallback")
        if mode == "segments":
            self.vox.voxelize_segments_dev(vertices, indices, origin, size, W, H, D, segs_per_strand=segs, flags=flags, out=out)
        else:
            self.vox.voxelize_vertices_dev(vertices, origin, size, W, H, D, flags=flags, out=out)

    def _satsum_cuda(self, slabs, out):
        self.vox.saturating_sum_u8_dev(slabs, out=out)

    def _exchange_slabs(self, partial, slab):
        """recv[r] = rank r's partial of MY slab.  NCCL: one all_to_all; gloo (CPU tests) has none: all_gather + slice."""
        import torch
        recv = torch.empty_like(partial)
        if self.dist.get_backend(self.group) == "nccl":
            self.dist.all_to_all_single(recv, partial, group=self.group)
        else:
            every = [torch.empty_like(partial) for _ in range(self.world)]
            self.dist.all_gather(every, partial, group=self.group)
            for r in range(self.world):
                recv[r * slab:(r + 1) * slab] = every[r][self.rank * slab:(self.rank + 1) * slab]
        return recv.view(self.world, slab)

    # ---- AABB: every rank must voxelise into the same box ------------------------------------------
    def global_bounding_box(self, local_min, local_max):
        """HairStyle::generate_bounding_box (hair_style.cc:215-234) over all shards: elementwise min / max of the
        per-rank boxes (each already folded from (0,0,0))."""
        import torch
        lo = torch.as_tensor(np.asarray(local_min, dtype=np.float32)).clone()
        hi = torch.as_tensor(np.asarray(local_max, dtype=np.float32)).clone()
        dev = self._collective_device()
        lo, hi = lo.to(dev), hi.to(dev)
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN, group=self.group)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX, group=self.group)
        return lo.cpu().numpy(), hi.cpu().numpy()

    def _collective_device(self):
        import torch
        backend = self.dist.get_backend(self.group)
        return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    # ---- the sharded voxelisation --------------------------------------------------------------------
    def voxelize(self, mode: str, vertices, indices, segs_per_strand: int, aabb_origin, aabb_size,
                 W: int, H: int, D: int, flags: int = 0, out=None, schedule: str = "rs_ag"):
        """All ranks call this with THEIR shard (torch tensors on the collective's device); every rank returns the
        complete W*H*D uint8 volume.  ``mode``: "segments" | "vertices"."""
        import torch
        if schedule not in ("allreduce", "rs_ag", "u8", "p2p"):
            raise ValueError("schedule must be 'allreduce', 'rs_ag', 'u8' or 'p2p'")
        nv = W * H * D
        nvp = padded_voxels(nv, self.world)
        dev = vertices.device
        if schedule == "p2p" and self.world > 1:
            return self._voxelize_p2p(mode, vertices, indices, segs_per_strand, aabb_origin, aabb_size, W, H, D, flags, out)
        if schedule == "u8" and self.world > 1:
            if out is None:
                out = torch.empty(nvp, dtype=torch.uint8, device=dev)
            elif out.numel() < nvp:
                raise ValueError(f"out must hold the padded grid ({nvp} bytes)")
            if getattr(self, "_partial_u8", None) is None or self._partial_u8.numel() != nvp or self._partial_u8.device != dev:
                self._partial_u8 = torch.zeros(nvp, dtype=torch.uint8, device=dev)      # the pad stays zero
            partial = self._partial_u8
            if vertices.numel():
                self._partial(mode, vertices, indices, segs_per_strand, aabb_origin, aabb_size, W, H, D, partial[:nv], flags & 1)
            else:
                partial.zero_()
            slab = nvp // self.world
            slabs = self._exchange_slabs(partial, slab)
            mine_u8 = torch.empty(slab, dtype=torch.uint8, device=dev)
            self._satsum(slabs, mine_u8)
            self.dist.all_gather_into_tensor(out[:nvp], mine_u8, group=self.group)
            vol = out[:nv]
            if flags & 2:
                if self.vox is None:
                    raise RuntimeError("NORMALIZE needs the CUDA library")
                self.vox.normalize_dev(vol)
            return vol
        if self._counts is None or self._counts.numel() != nvp or self._counts.device != dev:
            self._counts = torch.zeros(nvp, dtype=torch.int32, device=dev)
        else:
            self._counts.zero_()
        counts = self._counts
        if vertices.numel():
            self._count(mode, vertices, indices, segs_per_strand, aabb_origin, aabb_size, W, H, D, counts, flags & 1)
        norm = flags & 2                                  # NORMALIZE needs the whole grid: applied after the exchange
        if out is None:
            out = torch.empty(nvp, dtype=torch.uint8, device=dev)
        elif out.numel() < nvp:
            raise ValueError(f"out must hold the padded grid ({nvp} bytes)")
        if self.world == 1:
            self._clamp(counts, out[:nvp], 0)
        elif schedule == "allreduce":
            self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM, group=self.group)
            self._clamp(counts, out[:nvp], 0)
        else:
            slab = nvp // self.world
            mine = torch.empty(slab, dtype=torch.int32, device=dev)
            self.dist.reduce_scatter_tensor(mine, counts, op=self.dist.ReduceOp.SUM, group=self.group)
            mine_u8 = torch.empty(slab, dtype=torch.uint8, device=dev)
            self._clamp(mine, mine_u8, 0)
            self.dist.all_gather_into_tensor(out[:nvp], mine_u8, group=self.group)
        vol = out[:nv]
        if norm:
            if self.vox is None:
                raise RuntimeError("NORMALIZE needs the CUDA library")
            self.vox.normalize_dev(vol)
        return vol

    def _symmetric(self, nvp, dev):
        """(partial, bitmap, out, signals and their handles) in symmetric memory for a padded grid of nvp bytes; allocated
        and exchanged once per grid size."""
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        cached = getattr(self, "_symm", None)
        if cached is None or cached[0] != nvp:
            group = self.group if self.group is not None else self.dist.group.WORLD
            partial = symm_mem.empty(nvp, dtype=torch.uint8, device=dev)
            bitmap = symm_mem.empty(nvp // 512, dtype=torch.int32, device=dev)
            outbuf = symm_mem.empty(nvp, dtype=torch.uint8, device=dev)
            signals = symm_mem.empty(2 * 16, dtype=torch.int32, device=dev)
            hp, hb, ho, hs = (symm_mem.rendezvous(t, group) for t in (partial, bitmap, outbuf, signals))
            partial.zero_()
            signals.zero_()
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)                 # every rank's pads are zero before anyone signals
            self._symm = cached = (nvp, partial, bitmap, outbuf, signals, hp, hb, ho, hs)
        return cached[1:]

    def _voxelize_p2p(self, mode, vertices, indices, segs, origin, size, W, H, D, flags, out):
        """ONE C-ABI call per rank (vkhr_b200_voxelize_segments_sharded_dev): shard -> u8 partial, chunk bitmap, device-side
        barrier over the signal pads, the fused sparse peer-memory combine, second barrier."""
        if self.vox is None:
            raise RuntimeError("the p2p schedule needs the CUDA library and NVLink peer access")
        if mode != "segments":
            raise ValueError("the p2p schedule voxelises segments")
        nv = W * H * D
        nvp = self.vox.sharded_volume_bytes(W, H, D, self.world)
        partial, bitmap, outbuf, signals, hp, hb, ho, hs = self._symmetric(nvp, vertices.device)
        self.vox.voxelize_segments_sharded_dev(vertices, indices, origin, size, W, H, D, self.rank, hp.buffer_ptrs, hb.buffer_ptrs,
                                               ho.buffer_ptrs, hs.buffer_ptrs, segs_per_strand=segs, flags=flags & 3)
        vol = outbuf[:nv]
        if out is not None:
            if out.numel() < nv:
                raise ValueError("out is too small")
            out[:nv].copy_(vol)
            vol = out[:nv]
        return vol

    def voxelize_segments(self, vertices, indices, segs_per_strand, aabb_origin, aabb_size, W, H, D, **kw):
        return self.voxelize("segments", vertices, indices, segs_per_strand, aabb_origin, aabb_size, W, H, D, **kw)

    def voxelize_vertices(self, vertices, aabb_origin, aabb_size, W, H, D, **kw):
        return self.voxelize("vertices", vertices, None, 0, aabb_origin, aabb_size, W, H, D, **kw)
