"""The packet rule of warp-level `red`s, replayed on the CPU (numpy) -- the data behind DESIGN.md section 6.15.

    python tools/packet_model.py [W]

One warp-level `red.global.add.u32` leaves the SM as one request packet per distinct 32-byte sector its lanes touch --
EXCEPT that lanes adding to the same 32-bit word are not merged: each needs a packet of its own.  So the packets of an
instruction are  sum over its sectors of (the largest number of lanes on one word of that sector).  This script counts
that for the bench's strands (one crowd instance, ponytail shape, 256^3, the 4 x 4 x 2-voxel brick layout) and three
lane mappings of the interior walk, and prints packets per sample next to the plain distinct-sector count; ncu's
`l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red` of the same builds (profiles/r02_aa / r02_ac / r02_ad) is

    one segment per lane              120.03 M / 209.63 M samples = 0.5726   (model 0.5714)
    samples dealt out again           114.07 M                    = 0.5441   (model 0.5427)
    neighbour absorption               98.36 M                    = 0.4692   (model 0.4656)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from harness import synth

def count(keys, mask):
    """(packets, distinct sectors) summed over the rows (= warp instructions) of keys[mask]."""
    rows, L = keys.shape
    k = np.sort(np.where(mask, keys, -1 - np.arange(L)[None, :]), axis=1)          # idle lanes: unique negatives
    col = np.arange(L)[None, :].repeat(rows, 0)
    new = np.ones_like(k, bool); new[:, 1:] = k[:, 1:] != k[:, :-1]
    start = np.maximum.accumulate(np.where(new, col, 0), axis=1)
    last = np.ones_like(k, bool); last[:, :-1] = k[:, 1:] != k[:, :-1]
    mult = np.where(last, col - start + 1, 0)                                       # lanes per word, at the end of its run
    sec = np.where(k >= 0, k >> 3, k)
    snew = np.ones_like(sec, bool); snew[:, 1:] = sec[:, 1:] != sec[:, :-1]
    sid = np.cumsum(snew, axis=1) - 1
    r = np.arange(rows)[:, None].repeat(L, 1)
    mx = np.zeros((rows, L), np.int64)
    np.maximum.at(mx, (r, sid), mult)
    ok = np.zeros((rows, L), bool); ok[r, sid] = k >= 0
    return int((mx * ok).sum()), int(ok.sum())



def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    v, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
    lo, hi = synth.host_bounding_box(v)
    res = np.array([W, W, W], np.float32)
    vs = (hi - lo).astype(np.float32) / res
    P = ((v.reshape(-1, 3) - lo) / vs).astype(np.float32)             # vertices in voxel space, strand-major
    V, vps = P.shape[0], s + 1
    # the kernel's tiles: 31 vertices per warp-tile; lane t = vertex 31 * tile + t, idle when it ends a strand
    nt = (V - 1) // 31
    idx = np.arange(nt * 31).reshape(nt, 31)
    root, tip = P[idx], P[idx + 1]
    active = (idx % vps) != vps - 1
    d = tip - root
    steps = np.abs(d).max(axis=2)
    ns = np.ceil(steps).astype(np.int64) * active                      # samples of the lane's segment
    dirn = d / np.maximum(steps, 1e-30)[..., None]
    p = [root, root + dirn, root + dirn + dirn]                        # the first three samples (more are rare here)


    def word(q):
        """Word index of a sample in the brick-ordered scratch (brick_word in walk.cuh); word >> 3 = sector."""
        x = np.minimum(np.floor(q), res - 1).astype(np.int64)
        brick = ((x[..., 2] >> 1) * (W // 4) + (x[..., 1] >> 2)) * (W // 4) + (x[..., 0] >> 2)
        return brick * 8 + (x[..., 2] & 1) * 4 + (x[..., 1] & 3)


    w = [word(q) for q in p]
    total = int(ns.clip(max=3).sum())


    def add(*parts):
        return tuple(sum(x) for x in zip(*parts))


    out = {"W": W, "samples": total, "samples_per_segment": total / int(active.sum())}
    # (0) one segment per lane: instruction k = sample k of the tile's 31 lanes
    out["one_segment_per_lane"] = add(*[count(w[k], ns > k) for k in range(3)])


    # (1) dealt out again: an instruction carries both samples of 16 consecutive segments (lane j: sample j & 1 of segment j / 2)
    def dealt(a, b):
        L = b - a
        keys = np.empty((nt, 2 * L), np.int64); mask = np.empty((nt, 2 * L), bool)
        keys[:, 0::2], keys[:, 1::2] = w[0][:, a:b], w[1][:, a:b]
        mask[:, 0::2], mask[:, 1::2] = ns[:, a:b] > 0, ns[:, a:b] > 1
        return keys, mask


    out["dealt"] = add(count(*dealt(0, 16)), count(*dealt(16, 31)), count(w[2], ns > 2))
    # (1b) the same with runs of adjacent lanes on one word merged in software (what a match_any-style aggregation would add)
    merged = []
    for a, b in ((0, 16), (16, 31)):
        keys, mask = dealt(a, b)
        dup = np.zeros_like(mask); dup[:, 1:] = mask[:, 1:] & mask[:, :-1] & (keys[:, 1:] == keys[:, :-1])
        merged.append(count(keys, mask & ~dup))
    out["dealt_and_adjacent_lanes_merged"] = add(*merged, count(w[2], ns > 2))
    # (2) neighbour absorption: the second sample of lane t joins the first sample of lane t + 1 when the words agree
    v0, v1 = ns > 0, ns > 1
    nxt_w = np.concatenate([w[0][:, 1:], np.full((nt, 1), -5)], axis=1)
    nxt_v = np.concatenate([v0[:, 1:], np.zeros((nt, 1), bool)], axis=1)
    absorbed = v1 & nxt_v & (w[1] == nxt_w)
    out["neighbour_absorption"] = add(count(w[0], v0), count(w[1], v1 & ~absorbed), count(w[2], ns > 2))
    out["share_of_second_samples_absorbed"] = float(absorbed.sum() / v1.sum())
    for k in ("one_segment_per_lane", "dealt", "dealt_and_adjacent_lanes_merged", "neighbour_absorption"):
        pk, sec = out[k]
        out[k] = {"packets_per_sample": pk / total, "distinct_sectors_per_sample": sec / total}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
