"""tools/perf_probe.py -- per-phase device timings of the voxeliser on the named workloads (measurement tool).

    python tools/perf_probe.py [--reps 20] [--out gpurun_out/probe.json]
Times, with the library's per-phase CUDA events, single-instance and batched voxelisation for both
strategies at sigma ~ 2 and ~ 6.6 samples/segment, and prints one JSON object.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

import vkhr_b200
from vkhr_b200 import capi
from harness import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    vox = vkhr_b200.Voxelizer(0)
    res = {}
    cases = [("ponytail", 0.5, 256, 1), ("ponytail", 2.5, 256, 1), ("ponytail", 0.5, 256, 16), ("ponytail", 2.5, 256, 16),
             ("straight", 0.5, 512, 1), ("ponytail", 0.5, 512, 1), ("ponytail", 0.5, 1024, 1)]
    for shape, seg_len, W, inst in cases:
        items = []
        for k in range(inst):
            v, n, s = synth.shape(shape, seed=100 + k, seg_len=seg_len)
            lo, hi = synth.host_bounding_box(v)
            items.append({"vertices": torch.from_numpy(v).to(dev).reshape(-1), "segs_per_strand": s,
                          "aabb_origin": lo, "aabb_size": (hi - lo).astype(np.float32),
                          "out": torch.empty(W ** 3, dtype=torch.uint8, device=dev)})
        nseg = inst * n * s
        V = v.shape[0]
        batch = vox.make_batch(items)
        for strat_name, strat in (("packed8", capi.STRATEGY_PACKED8), ("count32", capi.STRATEGY_COUNT32)):
            if strat_name == "count32" and W >= 1024:
                continue
            for _ in range(3):
                vox.voxelize_segments_batch_dev(batch, W, W, W, flags=strat)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                vox.voxelize_segments_batch_dev(batch, W, W, W, flags=strat)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            vox.profile_enable(True)
            vox.profile_read()
            for _ in range(args.reps):
                vox.voxelize_segments_batch_dev(batch, W, W, W, flags=strat)
            prof = vox.profile_read()
            vox.profile_enable(False)
            alg = inst * (12 * V + W ** 3)
            key = f"{shape}_len{seg_len}_{W}^3_x{inst}_{strat_name}"
            res[key] = {"us_per_call": ms * 1e3, "Gseg_per_s": nseg / ms / 1e6,
                        "alg_GBs": alg / ms / 1e6, "frac_of_6530": alg / ms / 1e6 / 6530.3,
                        "phase_us": {k: prof[k]["ms"] / args.reps * 1e3 for k in prof},
                        "nonzero": int((items[0]["out"] != 0).sum()), "max": int(items[0]["out"].max())}
            print(key, json.dumps(res[key]), flush=True)
        del items, batch
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
