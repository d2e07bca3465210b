#!/usr/bin/env python
"""bench.py -- strand segments voxelised per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: the multi-character crowd of BASELINE.json configs[3] -- per-frame re-voxelisation of 64
ponytail-shaped instances (136,320 strands x 12 segments = 1,635,840 segments each, the shape of
configs[0]; ~104.7 M segments per frame) into one 256^3 u8 density volume per instance.  A "step" is one
frame of the WHOLE crowd.  The 64 instances are independent objects, so N ranks shard them (64 / N each) with
NO data-path collective, and the crowd stays the same at every N: "scaling": "strong" (this is the
configuration north_star's 85 % efficiency target is quoted on).  Inputs are synthetic (the reference's
.hair assets are Git-LFS pointers), 1.36 GB of strands + 1.07 GB of volumes per frame, larger than the L2.

Output: ONE JSON line on rank 0.  `value` is device-resident throughput (CUDA events, max over ranks); `e2e`
goes through the host-pointer C-ABI call with pinned host buffers (H2D + kernels + D2H inside the timed
region); `roofline` holds the frame kernel's algorithmic bytes against its event-timed duration;
`cpu_baseline` is the unmodified reference CPU voxeliser (whole call, all host cores) with its walk-only and
one-thread figures beside it; `strand_sharded` is BASELINE configs[2] (1 M strands x 32 segments at 512^3)
sharded by strand range over the N GPUs with an integer combine -- NCCL u32 all-reduce (north_star's form) and
the fused peer-memory kernel -- each checked byte for byte against the one-GPU volume.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified hair_style.cc; else
the C port) on a bounded sample of the same workload: one instance of the crowd per step, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "strand segments voxelised per second"
UNIT = "M seg/s"
CROWD = 64                     # instances of BASELINE configs[3]


def _emit(line: str):          # replaced in main() by a writer that keeps library chatter off stdout
    print(line, flush=True)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--instances", type=int, default=CROWD, help="crowd instances in total (sharded over the GPUs)")
    p.add_argument("--res", type=int, default=256)
    p.add_argument("--seg-len", type=float, default=0.5, help="synthetic segment length (0.5: ~2 samples/segment at 256^3)")
    p.add_argument("--strategy", default="auto", choices=["auto", "packed8", "count32", "brick8", "brick8-split"])
    p.add_argument("--ring-mib", type=int, default=0, help="BRICK8 scratch ring of the frame kernel in MiB (0 = library default)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-others", action="store_true", help="skip the short device-resident timings of the other BASELINE configs")
    p.add_argument("--no-sharded", action="store_true", help="skip the strand-sharded configs[2] block")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    p.add_argument("--cpu-seconds", type=float, default=10.0)
    return p.parse_args()


def workload_config(args) -> dict:
    """The `config` both arms print (identical: the driver compares them)."""
    W = args.res
    n_seg = 136_320 * 12
    return {
        "workload": f"multi-character crowd (BASELINE configs[3]): {args.instances} ponytail-shaped instances "
                    f"(136,320 strands x 12 segments = {n_seg} segments each, {args.instances * n_seg} segments per step), "
                    f"each re-voxelised into its own {W}^3 u8 volume every step",
        "instances": args.instances, "segments_per_instance": n_seg, "resolution": [W, W, W], "seg_len": args.seg_len,
    }


# --------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed regions run."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        mhz = sorted(m for m, _ in self.samples)
        return {"sm_mhz": float(mhz[len(mhz) // 2]), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons),
                "samples": len(mhz), "sm_mhz_min": float(mhz[0])}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(kernel_key)
    except Exception:  # noqa: BLE001
        return None


def make_instance(index: int, seg_len: float):
    """Instance `index` of the crowd (the same strands whatever the number of ranks)."""
    import numpy as np
    from harness import synth
    v, n, s = synth.shape("ponytail", seed=0x5EED + index, seg_len=seg_len)
    lo, hi = synth.host_bounding_box(v)
    return v, n, s, lo, (hi - lo).astype(np.float32)


def set_omp_threads(n: int):
    """OpenMP threads of the reference's quantise loop: pinned explicitly, the same in both arms and at every N
    (torchrun exports OMP_NUM_THREADS=1 to its workers, which used to make the N > 1 reference arm a different run)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except Exception:  # noqa: BLE001
        pass


def cpu_reference_figures(args, v0, n_strands, segs, lo, size, budget_s: float, check_against=None) -> dict:
    """The reference's CPU voxeliser on this box: the whole call with all cores (the stock behaviour), the same with one
    OpenMP thread (its quantise loop gets FASTER: SURVEY F12), and the walk alone (C port, density only, one core)."""
    import numpy as np
    import oracle
    W = args.res
    n_seg = n_strands * segs
    cores = os.cpu_count() or 1
    out = {"unit": UNIT, "cores": cores}
    P = oracle.port()
    idx = P.generate_indices(n_strands, segs)

    def timed(fn, seconds, min_runs=2):
        times = []
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end or len(times) < min_runs:
            t0 = time.perf_counter()
            r = fn()
            times.append(time.perf_counter() - t0)
        return sorted(times)[len(times) // 2], len(times), r

    if oracle.ref_available():
        hs = oracle.ref().create(v0, n_strands, segs)
        set_omp_threads(cores)
        med, runs, (d_ref, _, _) = timed(lambda: hs.voxelize("segments", W, W, W), budget_s * 0.5)
        out.update({"value": n_seg / med / 1e6, "kind": "reference",
                    "sample": f"instance 0 of the crowd ({n_seg} segments, {W}^3), unmodified reference HairStyle::voxelize_segments "
                              f"whole call (serial walk + tangent volume + OpenMP quantise on {cores} threads), median of {runs} runs"})
        set_omp_threads(1)
        med1, runs1, _ = timed(lambda: hs.voxelize("segments", W, W, W), budget_s * 0.25)
        out["whole_call_one_omp_thread"] = {"value": n_seg / med1 / 1e6, "unit": UNIT, "runs": runs1}
        set_omp_threads(cores)
        if check_against is not None:
            assert np.array_equal(check_against(), d_ref), "GPU volume differs from the reference's"
            out["sample"] += "; bit-exact vs GPU: yes"
    medw, runsw, d_port = timed(lambda: P.voxelize_segments(v0, idx, lo, size, W, W, W), budget_s * 0.25)
    out["walk_only"] = {"value": n_seg / medw / 1e6, "unit": UNIT, "cores": 1, "kind": "port", "runs": runsw,
                        "sample": "density-only C restatement of hair_style.cc:296-329 (no tangent volume, no quantise pass)"}
    if "value" not in out:
        out.update({"value": out["walk_only"]["value"], "kind": "port", "cores": 1, "sample": out["walk_only"]["sample"]})
        if check_against is not None:
            assert np.array_equal(check_against(), d_port), "GPU volume differs from the port's"
    return out


# --------------------------------------------------------------------------------------
def run_reference(args, rank: int):
    """The reference's own CPU voxeliser on this box's host cores: one instance of the crowd per step."""
    if rank != 0:
        return
    import oracle
    v, n, s, lo, size = make_instance(0, args.seg_len)
    W = args.res
    cores = os.cpu_count() or 1
    set_omp_threads(cores)
    if oracle.ref_available():
        kind = "reference"
        hs = oracle.ref().create(v, n, s)          # generate_bounding_box: the same AABB as make_instance

        def step():
            hs.voxelize("segments", W, W, W)
        threads = cores
        sample = (f"unmodified reference HairStyle::voxelize_segments({W}^3) whole call (serial walk + tangent volume + OpenMP "
                  f"quantise, OMP_NUM_THREADS pinned to {cores}), 1 instance of the crowd ({n * s} segments) per step")
    else:
        kind = "port"
        P = oracle.port()
        idx = P.generate_indices(n, s)

        def step():
            P.voxelize_segments(v, idx, lo, size, W, W, W)
        threads = 1
        sample = f"C port, density-only walk, 1 instance of the crowd ({n * s} segments) per step"
    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * s * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 walk -> u8 counts",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------
def timed_local(fn, reps):
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def other_configs(vox, dev, args, flags=0):
    """Short device-resident timings (CUDA events, 20 reps after 3 warm-ups) of the other BASELINE.json configs that
    fit one GPU; inputs are rotated over 8 copies (> L2) for the small sets.  Parity for these lives in tests/."""
    import numpy as np
    import torch
    from vkhr_b200 import capi
    from harness import synth
    out = {}
    cases = [("configs[0] ponytail 256^3, one instance", "ponytail", 256, 8, False),
             ("configs[0] ponytail 256^3, one instance, densities + tangent volume (the reference's default mode)", "ponytail", 256, 8, "tangents"),
             ("configs[1] Yuksel-straight-shaped 50,000 x 65 at 512^3", "straight", 512, 2, False),
             ("configs[1] Yuksel-curly-shaped 50,000 x 65 at 512^3", "curly", 512, 2, False),
             ("configs[4] animated ponytail frame at 1024^3 (voxelise only)", "ponytail", 1024, 1, False),
             ("configs[4] animated ponytail frame at 1024^3 (voxelise + AO/opacity prefilter)", "ponytail", 1024, 1, True)]
    for name, shape, W, copies, prefilter in cases:
        try:
            v, n, s = synth.shape(shape, seed=0x5EED, seg_len=0.5)
            lo, hi = synth.host_bounding_box(v)
            size = (hi - lo).astype(np.float32)
            vt = [torch.from_numpy(v).to(dev).reshape(-1).clone() for _ in range(copies)]
            o = [torch.empty(W ** 3, dtype=torch.uint8, device=dev) for _ in range(copies)]
            reps = 5 if prefilter else 20
            tangents = prefilter == "tangents"
            prefilter = prefilter is True
            pf = [torch.empty(W ** 3, dtype=torch.float32, device=dev) for _ in range(2)] if prefilter else None
            to = [torch.empty(4 * W ** 3, dtype=torch.int8, device=dev) for _ in range(copies)] if tangents else None

            def once(r):
                vox.voxelize_segments_dev(vt[r % copies], None, lo, size, W, W, W, segs_per_strand=s, out=o[r % copies], flags=flags,
                                          tangents_out=to[r % copies] if tangents else None)
                if prefilter:
                    vox.prefilter_dev(o[r % copies], W, W, W, ao=pf[0], opacity=pf[1])
            for r in range(3):
                once(r)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                once(r)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            alg = 12 * v.shape[0] + W ** 3 + (W ** 3 * (1 + 4 * 2) if prefilter else 0) + (4 * W ** 3 if tangents else 0)      # prefilter N^3 (1 + 4k)
            sname = {capi.STRATEGY_COUNT32: "count32", capi.STRATEGY_PACKED8: "packed8", capi.STRATEGY_BRICK8: "brick8"}.get(vox.last_strategy)
            out[name] = {"segments": n * s, "ms": ms, "value": n * s / ms / 1e3, "unit": UNIT, "strategy": sname,
                         "hbm_frac_whole_path": alg / (ms * 1e-3) / 1e9 / hbm_peak()[0]}
            del vt, o, pf, to
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    out["configs[4] animated ponytail, 120-frame sequence at 1024^3 (voxelise + AO/opacity prefilter per frame)"] = animated_sequence(vox, dev)
    return out


def animated_sequence(vox, dev, frames: int = 120, W: int = 1024):
    """BASELINE configs[4] as specified: the swayed ponytail of frame t = 0..119 re-voxelised at 1024^3 into the FIXED union
    AABB of the sequence (the reference keeps the load-time AABB, rasterizer/hair_style.cc:66), then the density ->
    AO / opacity prefilter, every frame; strands of all frames resident on the device (2.5 GB), CUDA events."""
    import numpy as np
    import torch
    from harness import synth
    try:
        v0, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
        lo, hi = synth.sway_union_bounding_box(v0, n, s, range(frames))
        size = (hi - lo).astype(np.float32)
        vt = [torch.from_numpy(synth.sway(v0, n, s, float(t))).to(dev).reshape(-1) for t in range(frames)]
        dens = torch.empty(W ** 3, dtype=torch.uint8, device=dev)
        ao = torch.empty(W ** 3, dtype=torch.float32, device=dev)
        op = torch.empty(W ** 3, dtype=torch.float32, device=dev)

        def frame(t, prefilter=True):
            vox.voxelize_segments_dev(vt[t], None, lo, size, W, W, W, segs_per_strand=s, out=dens)
            if prefilter:
                vox.prefilter_dev(dens, W, W, W, ao=ao, opacity=op)
        for t in range(3):
            frame(t)
        torch.cuda.synchronize()
        res = {}
        for name, pf in (("voxelise", False), ("voxelise_and_prefilter", True)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(frames):
                frame(t, pf)
            e1.record()
            torch.cuda.synchronize()
            res[name + "_ms_per_frame"] = e0.elapsed_time(e1) / frames
        alg = 12 * v0.shape[0] + W ** 3 + W ** 3 * (1 + 4 * 2)
        res.update({"frames": frames, "segments_per_frame": n * s, "nonzero_voxel_fraction_last_frame": float((dens != 0).float().mean().item()),
                    "value": n * s / res["voxelise_and_prefilter_ms_per_frame"] / 1e3, "unit": UNIT,
                    "hbm_frac_whole_path": alg / (res["voxelise_and_prefilter_ms_per_frame"] * 1e-3) / 1e9 / hbm_peak()[0]})
        del vt, dens, ao, op
        torch.cuda.empty_cache()
        return res
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def strand_sharded(vox, dev, rank: int, world: int, reps: int = 10) -> dict:
    """BASELINE configs[2]: 1 M strands x 32 segments (32 M segments) at 512^3, strands sharded by contiguous range over
    the ranks, partial volumes combined with ONE integer exchange.  Every schedule is asserted byte-identical to the
    volume rank 0 computes alone from the whole set."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from vkhr_b200 import sharding
    from harness import synth
    W = 512
    n, s = synth.shape_counts("big")
    first, count = sharding.strand_range(n, world, rank)
    mine, _, _ = synth.shape("big", seed=0x5EED, seg_len=0.5, first_strand=first, n_strands=count)
    mine_t = torch.from_numpy(mine).to(dev).reshape(-1)
    res = {"workload": f"BASELINE configs[2]: {n} strands x {s} segments = {n * s} segments at {W}^3, contiguous strand ranges "
                       f"over {world} GPU(s), one shared AABB, integer combine",
           "segments": n * s, "resolution": [W, W, W], "world": world}

    def timed(fn):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    if world == 1:
        lo, hi = vox.generate_bounding_box_dev(mine_t).cpu().numpy().reshape(2, 3)
        size = (hi - lo).astype(np.float32)
        out1 = torch.empty(W ** 3, dtype=torch.uint8, device=dev)
        ms, _ = timed(lambda: vox.voxelize_segments_dev(mine_t, None, lo, size, W, W, W, segs_per_strand=s, out=out1))
        res["one_gpu"] = {"ms": ms, "value": n * s / ms / 1e3, "unit": UNIT}
        return res
    sv = sharding.ShardedVoxelizer(vox)
    bb = vox.generate_bounding_box_dev(mine_t).cpu().numpy()
    lo, hi = sv.global_bounding_box(bb[:3], bb[3:])
    size = (hi - lo).astype(np.float32)
    nv = W ** 3
    out = torch.empty(sharding.padded_voxels(nv, world), dtype=torch.uint8, device=dev)
    vols = {}
    nvlink = {"allreduce": 2 * (world - 1) / world * 4 * nv, "p2p": 2 * (world - 1) / world * nv}
    for schedule in ("allreduce", "p2p"):
        try:
            ms, vol = timed(lambda: sv.voxelize_segments(mine_t, None, s, lo, size, W, W, W,
                                                         out=None if schedule == "p2p" else out, schedule=schedule))
            res[schedule] = {"ms": ms, "value": n * s / ms / 1e3, "unit": UNIT,
                             "nvlink_bytes_per_gpu" + ("_upper_bound_sparse_exchange" if schedule == "p2p" else ""): int(nvlink[schedule])}
            vols[schedule] = vol[:nv].clone()
            if schedule == "p2p":
                # where the time goes on this rank: the library's per-phase CUDA events (a separate pass)
                vox.profile_enable(True)
                vox.profile_read()
                for _ in range(reps):
                    sv.voxelize_segments(mine_t, None, s, lo, size, W, W, W, schedule="p2p")
                torch.cuda.synchronize()
                pr = vox.profile_read()
                vox.profile_enable(False)
                res[schedule]["phases_ms_rank0"] = {"shard_voxelisation": (pr["walk"]["ms"] + pr["finish"]["ms"]) / reps,
                                                    "chunk_bitmap_and_output_clear": pr["clear"]["ms"] / reps,
                                                    "first_barrier_(waiting_for_the_slowest_rank)": pr["normalize"]["ms"] / reps,
                                                    "fused_combine_and_second_barrier": pr["prefilter"]["ms"] / reps}
        except Exception as e:  # noqa: BLE001
            res[schedule] = {"error": str(e)[:300]}
    # the one-GPU volume of the whole set (rank 0 alone; the others wait), and the comparison
    ok = True
    if rank == 0:
        full, _, _ = synth.shape("big", seed=0x5EED, seg_len=0.5)
        full_t = torch.from_numpy(full).to(dev).reshape(-1)
        flo, fhi = vox.generate_bounding_box_dev(full_t).cpu().numpy().reshape(2, 3)
        res["aabb_equals_whole_set"] = bool(np.array_equal(flo, lo) and np.array_equal(fhi, hi))
        out1 = torch.empty(W ** 3, dtype=torch.uint8, device=dev)
        ms1, ref = timed_local(lambda: vox.voxelize_segments_dev(full_t, None, lo, size, W, W, W, segs_per_strand=s, out=out1), reps)
        res["one_gpu"] = {"ms": ms1, "value": n * s / ms1 / 1e3, "unit": UNIT}
        for name, v in vols.items():
            same = bool(torch.equal(v, ref))
            res[name]["byte_identical_to_one_gpu"] = same
            res[name]["speedup_vs_one_gpu"] = ms1 / res[name]["ms"]
            ok = ok and same
        del full_t, ref
    dist.barrier()
    torch.cuda.synchronize()
    assert ok, f"a sharded volume differs from the one-GPU volume: {res}"
    del out, vols
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vkhr_b200
    from vkhr_b200 import capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: unbound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    vox = vkhr_b200.Voxelizer(local_rank)
    W = args.res
    nvox = W * W * W
    total = args.instances
    mine = list(range(rank, total, world))                  # this rank's instances of the crowd
    I = len(mine)
    flags = {"auto": 0, "packed8": capi.STRATEGY_PACKED8, "count32": capi.STRATEGY_COUNT32, "brick8": capi.STRATEGY_BRICK8,
             "brick8-split": capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT}[args.strategy]
    if args.ring_mib:
        vox.set_scratch_ring_bytes(args.ring_mib << 20)

    # ---- inputs: this rank's instances, pinned on the host, resident on the device -------------
    v0, n_strands, segs, lo0, size0 = make_instance(mine[0], args.seg_len)
    V = v0.shape[0]
    n_seg = n_strands * segs
    host_v = torch.empty((I, V * 3), dtype=torch.float32).pin_memory()
    host_out = torch.empty((I, nvox), dtype=torch.uint8).pin_memory()
    aabbs = []
    for k, g in enumerate(mine):
        v, _, _, lo, size = (v0, n_strands, segs, lo0, size0) if k == 0 else make_instance(g, args.seg_len)
        host_v[k].numpy()[:] = v.reshape(-1)
        aabbs.append((lo, size))
    dev_v = host_v.to(dev)
    dev_out = torch.empty((I, nvox), dtype=torch.uint8, device=dev)
    batch = vox.make_batch([{"vertices": dev_v[k], "segs_per_strand": segs, "aabb_origin": aabbs[k][0],
                             "aabb_size": aabbs[k][1], "out": dev_out[k]} for k in range(I)])

    def frame():
        vox.voxelize_segments_batch_dev(batch, W, W, W, flags=flags)

    sampler = ClockSampler(local_rank)
    # ---- device-resident throughput -----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    sampler.start()
    l0 = vox.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        frame()
    e1.record()
    barrier()
    launches = vox.launch_count - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    step_ms = ms / args.steps
    value = n_seg * total * args.steps / (ms * 1e-3) / 1e6

    # ---- the same K steps with per-phase CUDA events on the launching stream (roofline of the dominant kernel) -----
    vox.profile_enable(True)
    vox.profile_read()
    for _ in range(args.steps):
        frame()
    torch.cuda.synchronize()
    prof = vox.profile_read()
    vox.profile_enable(False)
    phases_ms = {k: prof[k]["ms"] / args.steps for k in prof}
    brick = vox.last_strategy == capi.STRATEGY_BRICK8
    fused = brick and not (flags & capi.BRICK8_SPLIT)
    if fused:
        kernel, kernel_key = "k_frame<3,3> (BRICK8 strand walk + copy-out of the rank's instances, one persistent launch)", f"k_frame@{I}x{W}^3"
        kernel_ms = phases_ms["walk"]
    elif brick:
        kernel, kernel_key = ("k_walk_uniform<3,3> + k_untile_batch (BRICK8 as separate kernels; kernel_ms and traffic are the pair's)",
                              f"k_walk_uniform<3>+k_untile_batch@{I}x{W}^3")
        kernel_ms = phases_ms["walk"] + phases_ms["finish"]
    else:
        kernel, kernel_key = "k_walk_uniform (strand walk)", f"k_walk_uniform<1>@{I}x{W}^3"
        kernel_ms = phases_ms["walk"]
    alg_bytes = I * (12 * V + nvox)                                          # 12*V + W*H*D per instance
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    sname = {capi.STRATEGY_COUNT32: "count32", capi.STRATEGY_PACKED8: "packed8", capi.STRATEGY_BRICK8: "brick8"}.get(vox.last_strategy)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic(kernel_key), "kernel": kernel, "strategy": sname + ("" if fused or not brick else "-split"),
        "kernel_ms_per_launch": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
        "phase_ms_per_step": phases_ms, "kernel_share_of_step": kernel_ms / sum(phases_ms.values()) if sum(phases_ms.values()) else None,
        "whole_path_frac": (alg_bytes / (step_ms * 1e-3) / 1e9) / peak,
        "note": "per-rank figures of rank 0 (its instances of the crowd); event-timed on the launching stream in a separate pass of K steps",
    }

    # ---- end to end through the host-pointer C ABI (pinned host buffers) ------------------------
    e2e = None
    if not args.no_e2e:
        hv = [host_v[k].numpy() for k in range(I)]
        ho = [host_out[k].numpy() for k in range(I)]
        hbatch = vox.make_host_batch([{"vertices": hv[k], "segs_per_strand": segs, "aabb_origin": aabbs[k][0],
                                       "aabb_size": aabbs[k][1], "out": ho[k]} for k in range(I)])

        def frame_e2e():
            # one call for the rank's instances: upload of instance k+1, kernels of k and download of k-1 overlap
            vox.voxelize_segments_batch(hbatch, W, W, W, flags=flags)

        ke = args.e2e_steps or min(args.steps, 5)
        frame_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            frame_e2e()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": n_seg * total * ke / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": total * V * 12,
               "d2h_bytes_per_step": total * nvox, "steps": ke, "ms_per_step": dt / ke * 1e3,
               "api": "vkhr_b200_voxelize_segments_batch (host pointers, pinned; H2D / kernels / D2H of consecutive "
                      "instances pipelined on three streams), one call per rank per step", "host_affinity": numa}
        if rank == 0 and world == 1:
            # like for like with the reference call, which always builds the tangent volume too (hair_style.cc:331-339):
            # densities AND tangents of a few instances, one vkhr_b200_voxelize_segments call each, pinned host buffers
            try:
                nt = min(I, 8)
                td = torch.empty(nvox, dtype=torch.uint8).pin_memory()
                tt = torch.empty((nvox, 4), dtype=torch.int8).pin_memory()
                call = lambda k: vox.voxelize_segments(hv[k], None, aabbs[k][0], aabbs[k][1], W, W, W, segs_per_strand=segs,  # noqa: E731
                                                       flags=flags, want_tangents=True, out=td.numpy(), tangents_out=tt.numpy())
                call(0)
                t0 = time.perf_counter()
                for k in range(nt):
                    call(k)
                dtt = time.perf_counter() - t0
                e2e["densities_and_tangents"] = {
                    "value": n_seg * nt / dtt / 1e6, "unit": UNIT, "instances": nt, "ms_per_instance": dtt / nt * 1e3,
                    "h2d_bytes_per_instance": V * 12, "d2h_bytes_per_instance": nvox * 5,
                    "api": "vkhr_b200_voxelize_segments (host pointers, pinned), one call per instance, densities + int8x4 tangent "
                           "volume -- what the reference's HairStyle::voxelize_segments returns; not pipelined across instances"}
                del td, tt
            except Exception as e:  # noqa: BLE001
                e2e["densities_and_tangents"] = {"error": str(e)[:200]}
        # the frame that came back over PCIe must equal the device-resident one
        frame()
        torch.cuda.synchronize()
        for k in sorted({0, I // 2, I - 1}):
            assert torch.equal(host_out[k], dev_out[k].cpu()), "e2e result differs from the device-resident result"
    clocks = sampler.stop()

    # ---- the reference CPU voxeliser on this box's host cores (rank 0, N == 1 only) --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        def gpu_volume():
            frame()
            torch.cuda.synchronize()
            return dev_out[0].cpu().numpy()
        cpu = cpu_reference_figures(args, v0, n_strands, segs, aabbs[0][0], aabbs[0][1], args.cpu_seconds, check_against=gpu_volume)

    sharded = None
    if not args.no_sharded:
        del batch, dev_out, dev_v
        torch.cuda.empty_cache()
        try:
            sharded = strand_sharded(vox, dev, rank, world)
        except AssertionError:
            raise
        except Exception as e:  # noqa: BLE001
            sharded = {"error": str(e)[:300]}
    others = None
    if rank == 0 and world == 1 and not args.no_others:
        others = other_configs(vox, dev, args, flags)
    if rank == 0:
        sigma = None
        try:
            import oracle
            idx0 = oracle.port().generate_indices(n_strands, segs)
            sigma = oracle.port().count_samples(v0, idx0, aabbs[0][0], aabbs[0][1], W, W, W) / n_seg
        except Exception:  # noqa: BLE001
            pass
        if sigma and brick:
            # secondary bound: the SM -> L2 request path carries one 32-byte packet per distinct sector of a
            # `red` warp instruction, 220 G packets/s on this part (profiles/r01_microbench.json); the brick layout puts
            # 1 / 0.61 samples into one packet (ncu, profiles/traffic.json)
            sps = ncu_traffic(f"red_sectors_per_sample@{W}^3") or 0.61
            g_sectors = sigma * n_seg * I * sps / (kernel_ms * 1e-3) / 1e9
            roofline["atomic"] = {"samples_per_launch": sigma * n_seg * I, "achieved_Gsamples_s": sigma * n_seg * I / (kernel_ms * 1e-3) / 1e9,
                                  "request_sectors_per_sample": sps, "achieved_Gsectors_s": g_sectors,
                                  "peak_Gsectors_s": 220.0, "frac": g_sectors / 220.0,
                                  "peak_source": "tools/microbench.cu on this pool (profiles/r01_microbench.json)"}
        cfg = workload_config(args)
        cfg.update({"instances_per_gpu": I, "sharding": "by instance (rank r takes instances r, r + N, ...), no collective",
                    "samples_per_segment": sigma, "strategy": args.strategy,
                    "cache": f"inputs {I * V * 12 / 1e6:.0f} MB + outputs {I * nvox / 1e6:.0f} MB per rank per step"
                             + (", larger than the 126 MB L2; no flush needed" if I * (V * 12 + nvox) > 200e6 else
                                ": not far above the 126 MB L2 -- part of a rank's input may be served from L2 between steps")})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 walk -> u8 counts", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "strand_sharded": sharded, "other_configs": others,
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bind_to_gpu_numa_node(index: int) -> str:
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPUs next to GPU `index`."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {index}"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({str(e)[:60]})"
    return "unbound"


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.instances % max(world, 1):
        sys.exit(f"--instances {args.instances} is not a multiple of the {world} ranks")
    if args.impl == "reference":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)       # before libgomp initialises (torchrun exports 1)
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch ourselves with one process per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    # the host-side generator is OpenMP: give each rank its share of the cores (torchrun exports OMP_NUM_THREADS=1)
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(world, 1)))
    # Libraries (NCCL prints its version banner) may write to stdout; the contract is ONE JSON line there.  Everything
    # but the final line goes to stderr: fd 1 is pointed at fd 2 until the result is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line, flush=True)
        os.dup2(2, 1)
    globals()["_emit"] = emit
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
