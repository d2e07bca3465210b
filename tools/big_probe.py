"""tools/big_probe.py -- configs[2] (32 M segments at 512^3) and configs[1] on one GPU: frame kernel vs separate kernels."""
import json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import vkhr_b200
from vkhr_b200 import capi
from harness import synth
dev = torch.device("cuda", 0)
vox = vkhr_b200.Voxelizer(0)
res = {}
for shape in ("big", "straight"):
    v, n, s = synth.shape(shape, seed=0x5EED, seg_len=0.5)
    lo, hi = synth.host_bounding_box(v); size = (hi - lo).astype(np.float32)
    vt = torch.from_numpy(v).to(dev).reshape(-1)
    out = torch.empty(512 ** 3, dtype=torch.uint8, device=dev)
    for name, flags, ring in (("split", capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT, 64), ("frame", capi.STRATEGY_BRICK8, 512), ("packed8", capi.STRATEGY_PACKED8, 64)):
        vox.set_scratch_ring_bytes(ring << 20)
        for _ in range(3):
            vox.voxelize_segments_dev(vt, None, lo, size, 512, 512, 512, segs_per_strand=s, out=out, flags=flags)
        torch.cuda.synchronize()
        vox.profile_enable(True); vox.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            vox.voxelize_segments_dev(vt, None, lo, size, 512, 512, 512, segs_per_strand=s, out=out, flags=flags)
        e1.record(); torch.cuda.synchronize()
        pr = vox.profile_read(); vox.profile_enable(False)
        res[f"{shape}/{name}"] = {"ms": e0.elapsed_time(e1) / 10, "phases": {k: round(x["ms"] / 10, 4) for k, x in pr.items() if x["ms"]}}
print(json.dumps(res, indent=1))
