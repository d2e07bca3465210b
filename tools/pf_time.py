"""Device-resident timings of the prefilter / ADSM kernels (CUDA events), for profiles/."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vkhr_b200
from harness import synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vox = vkhr_b200.Voxelizer(0)
v, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
lo, hi = synth.host_bounding_box(v)
size = (hi - lo).astype(np.float32)
vt = torch.from_numpy(v).cuda().reshape(-1)
d = vox.voxelize_segments_dev(vt, None, lo, size, W, W, W, segs_per_strand=s)
out = [torch.empty(W ** 3, dtype=torch.float32, device="cuda") for _ in range(3)]
res = {"W": W, "lib": os.path.basename(os.environ.get("VKHR_B200_LIB", "libvkhr_b200.so")), "nonzero_fraction": float((d != 0).float().mean().item())}


def timed(name, fn, reps=5, alg_bytes=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    res[name] = {"ms": ms}
    if alg_bytes:
        res[name]["alg_GBs"] = alg_bytes / ms / 1e6
    print(name, res[name], flush=True)


n3 = W ** 3
timed("ao", lambda: vox.prefilter_dev(d, W, W, W, ao=out[0]), alg_bytes=5 * n3)
from vkhr_b200 import capi
timed("ao_rowwise", lambda: vox.prefilter_dev(d, W, W, W, ao=out[0], flags=capi.PREFILTER_ROWWISE), alg_bytes=5 * n3)
timed("opacity", lambda: vox.prefilter_dev(d, W, W, W, opacity=out[1]), alg_bytes=5 * n3)
timed("gauss3", lambda: vox.prefilter_dev(d, W, W, W, gauss=out[2]), alg_bytes=5 * n3)
timed("ao+opacity", lambda: vox.prefilter_dev(d, W, W, W, ao=out[0], opacity=out[1]), alg_bytes=9 * n3)
timed("ao+opacity+gauss3", lambda: vox.prefilter_dev(d, W, W, W, ao=out[0], opacity=out[1], gauss=out[2]), alg_bytes=13 * n3)
if W <= 512 and not os.environ.get("PF_NO_ADSM"):
    light = lo + size * np.array([0.5, 3.0, 0.5], np.float32)
    timed("adsm_1024steps", lambda: vox.adsm_dev(d, W, W, W, lo, size, light, out=out[0]), reps=2, alg_bytes=5 * n3)
print(json.dumps(res))
