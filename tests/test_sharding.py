"""Host logic of the multi-GPU path (vkhr_b200/sharding.py) on CPU: world_size-2 gloo processes, the CPU oracle
standing in for the counting / clamping kernels (injected; the package itself never does that)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vkhr_b200 import sharding
from harness import synth


def test_strand_range_partitions_every_strand_once():
    for n in (0, 1, 7, 8, 136_320, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.strand_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sharding.strand_range(10, 2, 2)


def test_padded_voxels():
    assert sharding.padded_voxels(256 ** 3, 8) == 256 ** 3
    assert sharding.padded_voxels(30 * 20 * 10, 8) % (16 * 8) == 0
    assert sharding.padded_voxels(1, 2) == 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, res, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        P = oracle.port()
        W, H, D = res
        v, n, s = synth.shape("ponytail", seed=0xBEEF, seg_len=1.1, scale=3000 / 136320)
        mine = np.ascontiguousarray(sharding.shard_vertices(v, n, s, world, rank))
        first, count = sharding.strand_range(n, world, rank)

        def count_fn(mode, vertices, indices, segs, origin, size, W, H, D, counts, flags):
            vv = vertices.numpy().reshape(-1, 3)
            if mode == "segments":
                c = P.count_segments(vv, P.generate_indices(vv.shape[0] // (segs + 1), segs), origin, size, W, H, D, flags=flags)
            else:
                c = P.count_vertices(vv, origin, size, W, H, D, flags=flags)
            counts[:W * H * D] += torch.from_numpy(c.astype(np.int32))

        def clamp_fn(counts, out, flags=0):
            out.copy_(torch.clamp(counts, max=255).to(torch.uint8))

        def partial_fn(mode, vertices, indices, segs, origin, size, W, H, D, out, flags):
            vv = vertices.numpy().reshape(-1, 3)
            if mode == "segments":
                d = P.voxelize_segments(vv, P.generate_indices(vv.shape[0] // (segs + 1), segs), origin, size, W, H, D)
            else:
                d = P.voxelize_vertices(vv, origin, size, W, H, D)
            out.copy_(torch.from_numpy(d))

        def satsum_fn(slabs, out):
            out.copy_(torch.clamp(slabs.to(torch.int32).sum(dim=0), max=255).to(torch.uint8))

        sv = sharding.ShardedVoxelizer(None, None, count_fn=count_fn, clamp_fn=clamp_fn, partial_fn=partial_fn, satsum_fn=satsum_fn)
        lo_l, hi_l = P.generate_bounding_box(mine)
        lo, hi = sv.global_bounding_box(lo_l, hi_l)
        size = (hi - lo).astype(np.float32)
        t = torch.from_numpy(mine.reshape(-1))
        got = {}
        for schedule in ("allreduce", "rs_ag", "u8"):
            got["seg_" + schedule] = sv.voxelize_segments(t, None, s, lo, size, W, H, D, schedule=schedule).numpy().copy()
            got["ver_" + schedule] = sv.voxelize_vertices(t, lo, size, W, H, D, schedule=schedule).numpy().copy()
        results[rank] = (lo, hi, got, count)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("res", [(32, 32, 32), (9, 6, 5)])
def test_two_rank_gloo_equals_single_process(res):
    import oracle
    P = oracle.port()
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), res, results), nprocs=world, join=True)
    W, H, D = res
    v, n, s = synth.shape("ponytail", seed=0xBEEF, seg_len=1.1, scale=3000 / 136320)
    lo, hi = P.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    want_seg = P.voxelize_segments(v, P.generate_indices(n, s), lo, size, W, H, D)
    want_ver = P.voxelize_vertices(v, lo, size, W, H, D)
    assert want_seg.max() == 255 or res != (9, 6, 5)        # the coarse grid saturates: clamp-after-sum is exercised
    assert sum(results[r][3] for r in range(world)) == n
    for r in range(world):
        rlo, rhi, got, _ = results[r]
        assert np.array_equal(rlo, lo) and np.array_equal(rhi, hi)
        for schedule in ("allreduce", "rs_ag", "u8"):
            assert np.array_equal(got["seg_" + schedule], want_seg), (r, schedule)
            assert np.array_equal(got["ver_" + schedule], want_ver), (r, schedule)
