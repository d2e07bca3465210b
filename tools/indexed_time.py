"""Uniform strands (no index buffer) vs the same strands through an explicit index buffer: device-resident timings."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vkhr_b200
from harness import synth

vox = vkhr_b200.Voxelizer(0)
res = {}
for shape, W in (("ponytail", 256), ("straight", 512), ("big", 512)):
    v, n, s = synth.shape(shape, seed=0x5EED, seg_len=0.5)
    lo, hi = synth.host_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    k = np.arange(n * s, dtype=np.int64)
    first = (k + k // s).astype(np.int32)
    idx = torch.from_numpy(np.stack([first, first + 1], axis=1).reshape(-1).copy()).cuda()
    vt = torch.from_numpy(v).cuda().reshape(-1)
    out = torch.empty(W ** 3, dtype=torch.uint8, device="cuda")
    ref = vox.voxelize_segments_dev(vt, None, lo, size, W, W, W, segs_per_strand=s).clone()
    for name, kw in (("uniform", dict(indices=None, segs_per_strand=s)), ("indexed", dict(indices=idx))):
        fn = lambda: vox.voxelize_segments_dev(vt, kw.get("indices"), lo, size, W, W, W, segs_per_strand=kw.get("segs_per_strand", 0), out=out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
        res[f"{shape}@{W} {name}"] = {"ms": a.elapsed_time(b) / 20, "segments": n * s}
print(json.dumps(res))
