"""Address statistics of the strand walk on the bench workload (CPU, numpy) -- the data behind DESIGN.md section 6.7.

For one crowd instance (ponytail shape, 256^3) this reproduces the kernel's lane mapping (a warp instruction = sample
iteration i of 31 consecutive segments of the vertex stream) and counts, per warp-level atomic instruction:
  * distinct 32-bit words   (what a match_any warp aggregation would leave)
  * distinct 32-byte sectors (what the SM -> L2 request path already merges), in the x-fastest layout and in the
    4x4x2 brick layout of the BRICK8 scratch volume
and, per strand, the share of its samples that fall into the brick (B^3 voxels) holding its root -- what a per-CTA
shared-memory brick histogram over strands binned by root could absorb.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from harness import synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
v, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
lo, hi = synth.host_bounding_box(v)
size = (hi - lo).astype(np.float32)
res = np.array([W, W, W], np.float32)
vs = size / res
p = ((v.reshape(-1, 3) - lo) / vs).astype(np.float32).reshape(n, s + 1, 3)
root, tip = p[:, :-1].reshape(-1, 3), p[:, 1:].reshape(-1, 3)
d = tip - root
steps = np.abs(d).max(axis=1)
nsamp = np.ceil(steps).astype(np.int64)
dirn = d / np.maximum(steps, 1e-30)[:, None]
nseg = root.shape[0]
out = {"W": W, "segments": int(nseg), "samples": int(nsamp.sum()), "samples_per_segment": float(nsamp.mean())}

# warp instruction = iteration i of a tile of 31 consecutive segments (strand ends are idle lanes in the kernel; ignored here)
T = 31
ntile = nseg // T
words_per_instr, sectors_per_instr, bricks_per_instr, lanes_per_instr = [], [], [], []
pos = root.copy()
for i in range(int(nsamp.max())):
    act = nsamp > i
    vox = np.minimum(np.floor(pos), res - 1).astype(np.int64)
    idx = vox[:, 0] + vox[:, 1] * W + vox[:, 2] * W * W
    bidx = ((vox[:, 2] >> 1) * (W // 4) + (vox[:, 1] >> 2)) * (W // 4) + (vox[:, 0] >> 2)      # 4x4x2 brick = one sector
    for name, shift, acc in (("w", 2, words_per_instr), ("s", 5, sectors_per_instr), ("b", 0, bricks_per_instr)):
        key = np.where(act, (bidx if name == "b" else idx >> shift), -1 - np.arange(nseg))[: ntile * T].reshape(ntile, T)
        a = act[: ntile * T].reshape(ntile, T)
        ks = np.sort(key, axis=1)
        distinct = 1 + (np.diff(ks, axis=1) != 0).sum(axis=1)
        inactive = T - a.sum(axis=1)
        acc.append((distinct - inactive)[a.any(axis=1)])
    lanes_per_instr.append(act[: ntile * T].reshape(ntile, T).sum(axis=1)[act[: ntile * T].reshape(ntile, T).any(axis=1)])
    pos = pos + dirn
    if i >= 7:
        break
lanes = np.concatenate(lanes_per_instr).sum()
out["atomic_lanes_counted"] = int(lanes)
out["distinct_words_per_lane"] = float(np.concatenate(words_per_instr).sum() / lanes)
out["distinct_sectors_per_lane"] = float(np.concatenate(sectors_per_instr).sum() / lanes)
out["distinct_brick_sectors_per_lane_4x4x2"] = float(np.concatenate(bricks_per_instr).sum() / lanes)

# brick absorption: samples of a strand inside the brick of the strand's root
strand_of = np.repeat(np.arange(n), s)
for B in (16, 32, 64):
    rb = np.floor(p[:, 0, :] / B).astype(np.int64)
    inside = total = 0
    pos = root.copy()
    for i in range(int(min(nsamp.max(), 8))):
        act = nsamp > i
        vb = (np.minimum(np.floor(pos), res - 1) // B).astype(np.int64)
        same = (vb == rb[strand_of]).all(axis=1)
        inside += int((same & act).sum()); total += int(act.sum())
        pos = pos + dirn
    out[f"share_of_samples_in_root_brick_{B}"] = inside / total
print(json.dumps(out, indent=1))
