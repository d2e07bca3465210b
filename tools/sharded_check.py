"""tools/sharded_check.py -- strand-sharded voxelisation over N GPUs (torchrun): parity vs one GPU + timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        tools/sharded_check.py [--strands 1000000] [--segs 32] [--res 512] [--reps 10]
Every rank voxelises its contiguous strand range into a partial u32 grid; the grids are combined with an NCCL
integer collective (all three schedules of vkhr_b200/sharding.py).  Rank 0 also voxelises the WHOLE set alone and the
volumes must be byte-identical.  Prints one JSON object on rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch
import torch.distributed as dist

import vkhr_b200
from vkhr_b200 import sharding
from harness import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strands", type=int, default=1_000_000)
    ap.add_argument("--segs", type=int, default=32)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--seg-len", type=float, default=0.5)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    vox = vkhr_b200.Voxelizer(local)
    sv = sharding.ShardedVoxelizer(vox)
    W = args.res
    n, s = args.strands, args.segs
    # every rank generates the whole set (seeded, identical everywhere) and keeps its contiguous strand range
    full = synth.strands(n, s, seed=0x5EED, root_min=(-25.0, 50.0, -25.0), seg_len=args.seg_len)
    mine = sharding.shard_vertices(full, n, s, world, rank)
    if rank != 0:
        mine = np.ascontiguousarray(mine)
        full = None
    mine_t = torch.from_numpy(np.ascontiguousarray(mine)).to(dev).reshape(-1)
    bb = vox.generate_bounding_box_dev(mine_t).cpu().numpy()
    lo, hi = sv.global_bounding_box(bb[:3], bb[3:])
    size = (hi - lo).astype(np.float32)
    res = {"world": world, "strands": n, "segments": n * s, "resolution": W}
    out = torch.empty(sharding.padded_voxels(W ** 3, world), dtype=torch.uint8, device=dev)
    vols = {}
    for schedule in ("allreduce", "rs_ag", "u8") + (("p2p",) if os.environ.get("VKHR_B200_P2P", "1") != "0" else ()):
        for _ in range(2):
            sv.voxelize_segments(mine_t, None, s, lo, size, W, W, W, out=None if schedule == "p2p" else out, schedule=schedule)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            vol = sv.voxelize_segments(mine_t, None, s, lo, size, W, W, W, out=None if schedule == "p2p" else out, schedule=schedule)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[schedule] = {"ms": float(ms.item()), "Gseg_per_s": n * s / float(ms.item()) / 1e6}
        vols[schedule] = vol.clone()
    if rank == 0:
        full_t = torch.from_numpy(full).to(dev).reshape(-1)
        flo, fhi = vox.generate_bounding_box(full)
        assert np.array_equal(flo, lo) and np.array_equal(fhi, hi), "sharded AABB differs from the whole set's"
        for _ in range(2):
            ref = vox.voxelize_segments_dev(full_t, None, lo, size, W, W, W, segs_per_strand=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            ref = vox.voxelize_segments_dev(full_t, None, lo, size, W, W, W, segs_per_strand=s)
        e1.record()
        torch.cuda.synchronize()
        res["one_gpu"] = {"ms": e0.elapsed_time(e1) / args.reps}
        for k, v in vols.items():
            res[k]["byte_identical_to_one_gpu"] = bool(torch.equal(v, ref))
        print(json.dumps(res), flush=True)
        assert all(res[k]["byte_identical_to_one_gpu"] for k in vols)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
