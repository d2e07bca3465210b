"""The packet rule (tools/packet_model.py) for configs[2]: `big` strands (1/8 of the 1 M, the rest behaves alike) at 512^3.

    python tools/packet_model.py        # the crowd's strands at 256^3
    python tools/packet_model_big.py    # this file

Prints samples per segment (3.0 here: two voxels per segment at this resolution), packets per sample of the walk as it
is (one segment per lane; ncu: 78.2 M red sectors for 96.2 M samples = 0.813, profiles/r02_an_cfg3_one_gpu.json) and
what absorbing the LAST sample of lane t into the first sample of lane t + 1 would leave.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from harness import synth
from packet_model import count

W = 512
v, n, s = synth.shape("big", seed=0x5EED, seg_len=0.5)
lo, hi = synth.host_bounding_box(v)
res = np.array([W, W, W], np.float32)
vs = (hi - lo).astype(np.float32) / res
keep = 125_000 * (s + 1)
P = ((v.reshape(-1, 3)[:keep] - lo) / vs).astype(np.float32)
V, vps = P.shape[0], s + 1
nt = (V - 1) // 31
idx = np.arange(nt * 31).reshape(nt, 31)
root, tip = P[idx], P[idx + 1]
active = (idx % vps) != vps - 1
d = tip - root
steps = np.abs(d).max(axis=2)
ns = np.ceil(steps).astype(np.int64) * active
dirn = d / np.maximum(steps, 1e-30)[..., None]


def word(q):
    x = np.minimum(np.floor(q), res - 1).astype(np.int64)
    brick = ((x[..., 2] >> 1) * (W // 4) + (x[..., 1] >> 2)) * (W // 4) + (x[..., 0] >> 2)
    return brick * 8 + (x[..., 2] & 1) * 4 + (x[..., 1] & 3)


K = min(int(ns.max()), 6)
p = [root]
for k in range(1, K):
    p.append(p[-1] + dirn)
w = [word(q) for q in p]
total = int(ns.clip(max=K).sum())
cur = [count(w[k], ns > k) for k in range(K)]
last_w = np.zeros_like(w[0])
for k in range(K):
    last_w = np.where(ns == k + 1, w[k], last_w)
v0 = ns > 0
nxt_w = np.concatenate([w[0][:, 1:], np.full((nt, 1), -5)], axis=1)
nxt_v = np.concatenate([v0[:, 1:], np.zeros((nt, 1), bool)], axis=1)
absorbed = (ns > 1) & nxt_v & (last_w == nxt_w)
left = sum(count(w[k], (ns > k) & ~(absorbed & (ns == k + 1)))[0] for k in range(K))
print(json.dumps({"W": W, "segments": int(active.sum()), "samples_per_segment": total / int(active.sum()),
                  "samples_per_segment_histogram": np.bincount(ns[active].clip(max=8)).tolist(),
                  "one_segment_per_lane": {"packets_per_sample": sum(c[0] for c in cur) / total,
                                           "distinct_sectors_per_sample": sum(c[1] for c in cur) / total},
                  "last_sample_absorbed_by_the_next_lane": {"packets_per_sample": left / total,
                                                            "share_of_segments": float(absorbed.sum() / (ns > 1).sum())}}, indent=1))
