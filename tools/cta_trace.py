"""tools/cta_trace.py -- per-CTA trace of the walk kernel: duration by SM, concurrency, tail (measurement tool)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import vkhr_b200
from vkhr_b200 import capi
from harness import synth

inst = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
vox = vkhr_b200.Voxelizer(0)
W = 256
items = []
for k in range(inst):
    v, n, s = synth.shape("ponytail", seed=100 + k, seg_len=0.5)
    lo, hi = synth.host_bounding_box(v)
    items.append({"vertices": torch.from_numpy(v).to(dev).reshape(-1), "segs_per_strand": s, "aabb_origin": lo,
                  "aabb_size": (hi - lo).astype(np.float32), "out": torch.empty(W ** 3, dtype=torch.uint8, device=dev)})
batch = vox.make_batch(items)
for _ in range(3):
    vox.voxelize_segments_batch_dev(batch, W, W, W)
torch.cuda.synchronize()
capi.check(vox.handle, capi.lib.vkhr_b200_debug_trace(vox.handle, 1, None, 0, None))
vox.voxelize_segments_batch_dev(batch, W, W, W)
torch.cuda.synchronize()
buf = np.zeros((1 << 20, 4), dtype=np.uint64)
n = C.c_uint32(0)
capi.check(vox.handle, capi.lib.vkhr_b200_debug_trace(vox.handle, 0, C.c_void_p(buf.ctypes.data), 1 << 20, C.byref(n)))
r = buf[: n.value]
sm, t0, t1 = r[:, 0].astype(int), r[:, 1].astype(np.int64), r[:, 2].astype(np.int64)
base = t0.min()
dur = (t1 - t0) / 1e3
print("records", n.value, "span us", (t1.max() - base) / 1e3)
print("CTA(warp0) duration us: mean %.2f p10 %.2f p50 %.2f p90 %.2f max %.2f" % (dur.mean(), *np.percentile(dur, [10, 50, 90]), dur.max()))
per_sm_cnt = np.bincount(sm, minlength=148)
per_sm_dur = np.bincount(sm, weights=dur, minlength=148) / np.maximum(per_sm_cnt, 1)
print("CTAs per SM: min %d max %d ; mean duration per SM: min %.2f max %.2f" % (per_sm_cnt.min(), per_sm_cnt.max(), per_sm_dur.min(), per_sm_dur.max()))
order = np.argsort(per_sm_dur)
print("fastest SMs", [(int(i), int(per_sm_cnt[i]), round(float(per_sm_dur[i]), 1)) for i in order[:8]])
print("slowest SMs", [(int(i), int(per_sm_cnt[i]), round(float(per_sm_dur[i]), 1)) for i in order[-8:]])
# concurrency over time (CTAs running), 20 buckets
T = t1.max() - base
edges = np.linspace(0, T, 21)
for a, b in zip(edges[:-1], edges[1:]):
    running = ((t0 - base < b) & (t1 - base > a)).sum()
    pass
cls = {}
for i in range(148):
    cls.setdefault(int(round(per_sm_dur[i])), []).append(i)
print("duration classes (us: #SMs):", {k: len(v) for k, v in sorted(cls.items())})
json.dump({"n": int(n.value), "sm": sm.tolist(), "t0": (t0 - base).tolist(), "t1": (t1 - base).tolist()}, open("gpurun_out/cta_trace.json", "w"))
