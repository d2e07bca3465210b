#!/usr/bin/env python
"""bench.py -- strand segments voxelised per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default): the multi-character crowd of BASELINE.json configs[3] -- per-frame
re-voxelisation of ponytail-shaped instances (136,320 strands x 12 segments = 1,635,840 segments
each, the shape of configs[0]) into one 256^3 u8 density volume per instance.  A "step" is one frame:
every instance owned by the rank is voxelised once.  Instances are independent objects, so ranks shard
them with NO data-path collective ("scaling": "weak": `--instances` per GPU, 64 by default, i.e. at
N=1 exactly the 64-instance / ~104.7 M-segment crowd).  Inputs are synthetic (the reference's .hair
assets are Git-LFS pointers) and 1.36 GB per rank, far larger than the 126 MB L2, so consecutive
steps cannot be served from cache.

Output: ONE JSON line on rank 0 (see the keys at the bottom).  `value` is device-resident throughput
(CUDA events, max over ranks); `e2e` goes through the host-pointer C-ABI call with pinned host buffers
(H2D + kernels + D2H inside the timed region); `roofline` is the walk kernel's algorithmic bytes over
its event-timed duration; `cpu_baseline` is the unmodified reference CPU voxeliser timed on this box.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified
hair_style.cc; else the C port) on the same per-instance workload, one instance per step.
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

import numpy as np

METRIC = "strand segments voxelised per second"


def _emit(line: str):          # replaced in main() by a writer that keeps library chatter off stdout
    print(line, flush=True)

UNIT = "M seg/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--instances", type=int, default=64, help="crowd instances per GPU")
    p.add_argument("--res", type=int, default=256)
    p.add_argument("--seg-len", type=float, default=0.5, help="synthetic segment length (0.5: ~2 samples/segment at 256^3)")
    p.add_argument("--strategy", default="auto", choices=["auto", "packed8", "count32", "brick8", "brick8-split"])
    p.add_argument("--ring-mib", type=int, default=0, help="BRICK8 scratch ring of the frame kernel in MiB (0 = library default)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-others", action="store_true", help="skip the short device-resident timings of the other BASELINE configs")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


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


def make_instance(seed: int, seg_len: float):
    from harness import synth
    v, n, s = synth.shape("ponytail", seed=seed, seg_len=seg_len)
    lo, hi = synth.host_bounding_box(v)
    return v, n, s, lo, (hi - lo).astype(np.float32)


def other_configs(vox, dev, args, flags=0):
    """Short device-resident timings (CUDA events, 20 reps after 3 warm-ups) of the other BASELINE.json configs that
    fit one GPU; inputs are rotated over 8 copies (> L2) for the small sets.  Parity for these lives in tests/."""
    import torch
    from vkhr_b200 import capi
    from harness import synth
    out = {}
    cases = [("configs[0] ponytail 256^3, one instance", "ponytail", 0.5, 256, 8),
             ("configs[1] Yuksel-straight-shaped 50,000 x 65 at 512^3", "straight", 0.5, 512, 2),
             ("configs[2] 1M strands x 32 segments at 512^3 (one GPU)", "big", 0.5, 512, 1),
             ("configs[4] animated ponytail frame at 1024^3 (voxelise only)", "ponytail", 0.5, 1024, 1),
             ("configs[4] animated ponytail frame at 1024^3 (voxelise + AO/opacity prefilter)", "ponytail", 0.5, 1024, 1)]
    for name, shape, seg_len, W, copies in cases:
        try:
            v, n, s = synth.shape(shape, seed=0x5EED, seg_len=seg_len)
            lo, hi = synth.host_bounding_box(v)
            size = (hi - lo).astype(np.float32)
            vt = [torch.from_numpy(v).to(dev).reshape(-1).clone() for _ in range(copies)]
            o = [torch.empty(W ** 3, dtype=torch.uint8, device=dev) for _ in range(copies)]
            prefilter = "prefilter" in name
            reps = 5 if prefilter else 20
            pf = [torch.empty(W ** 3, dtype=torch.float32, device=dev) for _ in range(2)] if prefilter else None

            def once(r):
                vox.voxelize_segments_dev(vt[r % copies], None, lo, size, W, W, W, segs_per_strand=s, out=o[r % copies], flags=flags)
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
            alg = 12 * v.shape[0] + W ** 3 + (W ** 3 * (1 + 4 * 2) if prefilter else 0)      # prefilter N^3 (1 + 4k)
            sname = {capi.STRATEGY_COUNT32: "count32", capi.STRATEGY_PACKED8: "packed8", capi.STRATEGY_BRICK8: "brick8"}.get(vox.last_strategy)
            out[name] = {"segments": n * s, "ms": ms, "value": n * s / ms / 1e3, "unit": UNIT, "strategy": sname,
                         "hbm_frac_whole_path": alg / (ms * 1e-3) / 1e9 / hbm_peak()[0]}
            del vt, o, pf
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    return out


# --------------------------------------------------------------------------------------
def run_reference(args, rank: int):
    """The reference's own CPU voxeliser on this box's host cores: one ponytail instance per step."""
    if rank != 0:
        return
    import oracle
    v, n, s, lo, size = make_instance(0x5EED, args.seg_len)
    W = args.res
    cores = os.cpu_count() or 1
    if oracle.ref_available():
        kind = "reference"
        hs = oracle.ref().create(v, n, s)          # generate_bounding_box: the same AABB as make_instance

        def step():
            hs.voxelize("segments", W, W, W)
        threads = int(os.environ.get("OMP_NUM_THREADS", cores))
        sample = (f"unmodified reference HairStyle::voxelize_segments({W}^3) whole call (serial walk + OpenMP tangent "
                  f"quantise), 1 ponytail instance ({n * s} segments) per step")
    else:
        kind = "port"
        P = oracle.port()
        idx = P.generate_indices(n, s)

        def step():
            P.voxelize_segments(v, idx, lo, size, W, W, W)
        threads = 1
        sample = f"C port, density-only walk, 1 ponytail instance ({n * s} segments) per step"
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 walk -> u8 counts",
        "data": "synthetic",
        "config": {"workload": f"crowd of ponytail-shaped instances (136,320 strands x 12 segments) at {W}^3; "
                               "one instance per step on the host CPU", "resolution": [W, W, W], "seg_len": args.seg_len},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
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
    I = args.instances
    flags = {"auto": 0, "packed8": capi.STRATEGY_PACKED8, "count32": capi.STRATEGY_COUNT32, "brick8": capi.STRATEGY_BRICK8,
             "brick8-split": capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT}[args.strategy]
    if args.ring_mib:
        vox.set_scratch_ring_bytes(args.ring_mib << 20)

    # ---- inputs: I instances per rank, pinned on the host, resident on the device -------------
    v0, n_strands, segs, lo0, size0 = make_instance(0x5EED + rank * I, args.seg_len)
    V = v0.shape[0]
    n_seg = n_strands * segs
    host_v = torch.empty((I, V * 3), dtype=torch.float32).pin_memory()
    host_out = torch.empty((I, nvox), dtype=torch.uint8).pin_memory()
    aabbs = []
    for k in range(I):
        v, _, _, lo, size = (v0, n_strands, segs, lo0, size0) if k == 0 else make_instance(0x5EED + rank * I + k, args.seg_len)
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
    value = n_seg * I * world * args.steps / (ms * 1e-3) / 1e6

    # ---- the same K steps with per-phase CUDA events (for the roofline of the walk kernel) -----
    vox.profile_enable(True)
    vox.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        frame()
    p1.record()
    torch.cuda.synchronize()
    prof = vox.profile_read()
    vox.profile_enable(False)
    ms_instr = p0.elapsed_time(p1)
    walk_ms = prof["walk"]["ms"] / max(prof["walk"]["spans"], 1)             # per launch
    if vox.last_strategy == capi.STRATEGY_BRICK8:
        # the volume exists only after the copy-out: the pair (walk, copy-out) is what the algorithmic bytes are held against
        walk_ms = (prof["walk"]["ms"] + prof["finish"]["ms"]) / args.steps      # one walk + one copy-out (+ the repair's look at the flags) per step
    alg_bytes = I * (12 * V + nvox)                                          # 12*V + W*H*D per instance
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else 0.0
    phases_ms = {k: prof[k]["ms"] / args.steps for k in prof}
    step_ms = ms / args.steps
    strat = {capi.STRATEGY_COUNT32: (0, "count32", "u32 counts"), capi.STRATEGY_PACKED8: (1, "packed8", "packed u8 atomics in the output volume"),
             capi.STRATEGY_BRICK8: (3, "brick8", "packed u8 atomics in the brick-ordered scratch volume, + k_untile_batch, the copy-out: kernel_ms is the pair")
             }.get(vox.last_strategy, (1, "packed8", "packed u8 atomics"))
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic(f"k_walk_uniform<{strat[0]}>@{I}x{W}^3"),
        "kernel": f"k_walk_uniform<{strat[0]}> (strand walk, {strat[2]})", "strategy": strat[1], "kernel_ms_per_launch": walk_ms,
        "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
        "phase_ms_per_step": phases_ms, "kernel_share_of_step": (phases_ms["walk"] / (ms_instr / args.steps)) if ms_instr else None,
        "whole_path_frac": (alg_bytes / (step_ms * 1e-3) / 1e9) / peak,
    }

    # ---- end to end through the host-pointer C ABI (pinned host buffers) ------------------------
    e2e = None
    if not args.no_e2e:
        hv = [host_v[k].numpy() for k in range(I)]
        ho = [host_out[k].numpy() for k in range(I)]
        hbatch = vox.make_host_batch([{"vertices": hv[k], "segs_per_strand": segs, "aabb_origin": aabbs[k][0],
                                       "aabb_size": aabbs[k][1], "out": ho[k]} for k in range(I)])

        def frame_e2e():
            # one call for the crowd: upload of instance k+1, kernels of k and download of k-1 overlap
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
        e2e = {"value": n_seg * I * world * ke / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": I * V * 12,
               "d2h_bytes_per_step": I * nvox, "steps": ke, "ms_per_step": dt / ke * 1e3,
               "api": "vkhr_b200_voxelize_segments_batch (host pointers, pinned; H2D / kernels / D2H of consecutive "
                      "instances pipelined on three streams)", "host_affinity": numa}
        # the frame that came back over PCIe must equal the device-resident one
        frame()
        torch.cuda.synchronize()
        for k in (0, I // 2, I - 1):
            assert torch.equal(host_out[k], dev_out[k].cpu()), "e2e result differs from the device-resident result"
    clocks = sampler.stop()

    # ---- the reference CPU voxeliser on this box's host cores (rank 0, N == 1 only) --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        cores = os.cpu_count() or 1
        if oracle.ref_available():
            hs = oracle.ref().create(v0, n_strands, segs)
            times = []
            t_end = time.perf_counter() + args.cpu_seconds
            while time.perf_counter() < t_end or len(times) < 2:
                t0 = time.perf_counter()
                d_ref, _, _ = hs.voxelize("segments", W, W, W)
                times.append(time.perf_counter() - t0)
            frame()
            torch.cuda.synchronize()
            assert np.array_equal(dev_out[0].cpu().numpy(), d_ref), "GPU volume differs from the reference's"
            med = sorted(times)[len(times) // 2]
            cpu = {"value": n_seg / med / 1e6, "unit": UNIT, "cores": int(os.environ.get("OMP_NUM_THREADS", cores)),
                   "kind": "reference",
                   "sample": f"instance 0 of the crowd ({n_seg} segments, {W}^3), unmodified reference "
                             f"HairStyle::voxelize_segments whole call, median of {len(times)} runs; bit-exact vs GPU: yes"}
        else:
            P = oracle.port()
            idx = P.generate_indices(n_strands, segs)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < args.cpu_seconds:
                d_ref = P.voxelize_segments(v0, idx, aabbs[0][0], aabbs[0][1], W, W, W)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": n_seg * reps / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"instance 0 of the crowd ({n_seg} segments, {W}^3), C port density-only walk, {reps} runs"}

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
        if sigma:
            # secondary bound: one 32-byte atomic request packet per sample; the measured ceiling of
            # packed atomics with one lane per sector on this part is 220 G/s (profiles/r01_microbench.json)
            # per SECTOR-request, whatever the number of lanes in it; BRICK8 puts 1 / 0.6 samples into one (ncu, profiles/traffic.json)
            sps = ncu_traffic(f"atom_sectors_per_sample<{strat[0]}>@{I}x{W}^3") or 1.0
            g_sectors = sigma * n_seg * I * sps / (walk_ms * 1e-3) / 1e9
            roofline["atomic"] = {"samples_per_launch": sigma * n_seg * I, "achieved_Gsamples_s": sigma * n_seg * I / (walk_ms * 1e-3) / 1e9,
                                  "request_sectors_per_sample": sps, "achieved_Gsectors_s": g_sectors,
                                  "peak_Gsectors_s": 220.0, "frac": g_sectors / 220.0,
                                  "peak_source": "tools/microbench.cu on this pool (profiles/r01_microbench.json)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 walk -> u8 counts", "data": "synthetic",
            "config": {
                "workload": f"multi-character crowd (BASELINE configs[3]): {I} ponytail-shaped instances per GPU "
                            f"(136,320 strands x 12 segments = {n_seg} segments each, {I * n_seg} segments per GPU per "
                            f"step), each re-voxelised into its own {W}^3 u8 volume; instances sharded across GPUs, "
                            "no collective",
                "instances_per_gpu": I, "segments_per_instance": n_seg, "resolution": [W, W, W],
                "seg_len": args.seg_len, "samples_per_segment": sigma, "strategy": args.strategy,
                "cache": f"inputs {I * V * 12 / 1e6:.0f} MB + outputs {I * nvox / 1e6:.0f} MB per rank per step, larger than the 126 MB L2; no flush needed",
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "other_configs": others,
        }
        _emit(json.dumps(line))
    if world > 1:
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
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch ourselves with one process per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
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
