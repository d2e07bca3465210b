"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libvkhr_ref.so).

Run in the build container (needs /root/reference to build the reference
driver):   python tests/golden/make_golden.py

The reference itself has no tests or golden vectors for this path
(SURVEY.md F11), so the pins are outputs of the reference's own
``HairStyle::voxelize_segments`` / ``voxelize_vertices`` / ``Volume::normalize``
/ ``Volume::downsample`` (src/vkhr/scene_graph/hair_style.cc:257-357) on
  * the hand-verifiable 4^3 known-answer input of SURVEY.md Appendix B,
  * small seeded strand sets (inputs stored next to the outputs),
  * full-size seeded sets, stored as fingerprints (FNV-1a-64, sum, non-zero,
    saturated) together with the FNV of the generated input.
tests/test_oracle.py checks the C restatement (oracle/voxel_oracle.c) against
all of them; the GPU parity tests check the CUDA path against the same files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import oracle  # noqa: E402
from harness import synth# noqa: E402  (host-side generator only)

KAT_VERTICES = [[0, 0, 0], [4, 4, 4], [0.5, 3.5, 0.5], [3, 3.5, 0.5], [4, 4, 4], [3, 4, 2.5], [1, 1, 1], [1, 1, 1]]


def fnv(a) -> str:
    return f"{oracle.port().fnv1a64(a):016x}"


def sparse(d):
    nz = np.nonzero(d)[0]
    return {str(int(i)): int(d[i]) for i in nz}


def stats(d):
    return {"fnv": fnv(d), "sum": int(d.astype(np.int64).sum()), "nonzero": int(np.count_nonzero(d)),
            "saturated": int(np.count_nonzero(d == 255)), "max": int(d.max())}


SWAY_FRAMES = (0, 59, 119)          # config 5: first, middle and last frame of the 120-frame sequence


def at_size(R):
    """BASELINE.json configs 1, 2, 4 at their FULL sizes (fingerprints only; run with --sizes, merged into golden.json):
    Yuksel-shaped straight / curly 50,000 x 65 at 512^3, 1 M strands x 32 at 512^3, and swayed ponytail frames in the
    union AABB of the 120-frame sequence at 1024^3 (the reference needs ~18 GB and ~2 minutes per frame there)."""
    fp = {}

    def one(name, vb, nb, sb, res, extra, aabb=None):
        h = R.create(vb, nb, sb) if aabb is None else R.create(vb, nb, sb, aabb_min=aabb[0], aabb_max=aabb[1])
        d, _, sec = h.voxelize("segments", *res)
        e = dict(extra)
        e.update({"strands": nb, "segments_per_strand": sb, "resolution": list(res), "input_fnv": fnv(vb),
                  "aabb": [float(x) for x in h.aabb], "segments": stats(d)})
        e["segments"]["reference_seconds"] = round(sec, 3)
        e["normalize_segments"] = stats(R.normalize(d))
        fp[name] = e
        print(name, json.dumps(e)[:400], flush=True)
        h.close()

    for name, shape in (("straight_512_full", "straight"), ("curly_512_full", "curly"), ("big_512_full", "big")):
        vb, nb, sb = synth.shape(shape, seed=0x5EED, seg_len=0.5)
        one(name, vb, nb, sb, (512, 512, 512), {"shape": shape, "seed": 0x5EED, "seg_len": 0.5, "scale": 1.0})
    v0, n0, s0 = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
    lo, hi = synth.sway_union_bounding_box(v0, n0, s0, range(120))
    for t in SWAY_FRAMES:
        vt = synth.sway(v0, n0, s0, float(t))
        one(f"ponytail_sway_t{t}_1024", vt, n0, s0, (1024, 1024, 1024),
            {"shape": "ponytail", "seed": 0x5EED, "seg_len": 0.5, "scale": 1.0, "sway_t": t,
             "union_aabb_min": [float(x) for x in lo], "union_aabb_max": [float(x) for x in hi]}, aabb=(lo, hi))
    return fp


def main():
    R = oracle.ref()
    if "--sizes" in sys.argv:
        path = os.path.join(HERE, "golden.json")
        with open(path) as f:
            out_json = json.load(f)
        out_json["fingerprints_at_size"] = at_size(R)
        with open(path, "w") as f:
            json.dump(out_json, f, indent=1, sort_keys=True)
        print("merged fingerprints_at_size into", path)
        return
    out_json = {}
    try:                                    # keep the at-size fingerprints of an earlier --sizes run
        with open(os.path.join(HERE, "golden.json")) as f:
            out_json["fingerprints_at_size"] = json.load(f)["fingerprints_at_size"]
    except Exception:  # noqa: BLE001
        pass

    # ---- 1. Appendix B known answers -------------------------------------------------
    v = np.array(KAT_VERTICES, dtype=np.float32)
    hs = R.create(v, 4, 1)
    seg, tang, _ = hs.voxelize("segments", 4, 4, 4, want_tangents=True)
    ver, vtang, _ = hs.voxelize("vertices", 4, 4, 4, want_tangents=True)
    out_json["kat4"] = {
        "vertices": KAT_VERTICES, "strands": 4, "segments_per_strand": 1,
        "aabb": [float(x) for x in hs.aabb], "indices": hs.indices.tolist(),
        "tangents_in": hs.tangents.tolist(),
        "voxelize_segments": sparse(seg),
        "voxelize_segments_tangents": {str(int(i)): tang[i].tolist() for i in np.nonzero(seg)[0]},
        "voxelize_vertices": sparse(ver),
        "voxelize_vertices_tangents": {str(int(i)): vtang[i].tolist() for i in np.nonzero(ver)[0]},
        "normalize_segments": sparse(R.normalize(seg)),
        "downsample_sum_segments": R.downsample(seg, 4, 4, 4, 2).tolist(),
        "downsample_max_segments": R.downsample(seg, 4, 4, 4, 0).tolist(),
    }

    # ---- 2. small seeded sets, inputs stored -------------------------------------------
    small = {}
    vs, n, s = synth.shape("ponytail", seed=0xA11CE, seg_len=0.9, scale=2000 / 136320)
    small["in_vertices"] = vs
    small["in_meta"] = np.array([n, s], dtype=np.int64)
    hs = R.create(vs, n, s)                                  # generated AABB (contains the origin, F7)
    small["aabb_generated"] = hs.aabb
    for (W, H, D) in [(64, 64, 64), (64, 32, 16), (16, 16, 16), (30, 20, 10)]:
        tag = f"{W}x{H}x{D}"
        dseg, tseg, _ = hs.voxelize("segments", W, H, D, want_tangents=True)
        dver, _, _ = hs.voxelize("vertices", W, H, D)
        small[f"seg_{tag}"] = dseg
        small[f"segtan_{tag}"] = tseg
        small[f"ver_{tag}"] = dver
        small[f"segnorm_{tag}"] = R.normalize(dseg)
        if W % 2 == 0 and H % 2 == 0 and D % 2 == 0:
            for f in range(4):
                small[f"segdown{f}_{tag}"] = R.downsample(dseg, W, H, D, f)
    # header AABB tight around the data (the has_bounding_box route of real assets)
    lo, hi = vs.min(axis=0), vs.max(axis=0)
    hs2 = R.create(vs, n, s, aabb_min=lo, aabb_max=hi)
    small["aabb_header"] = hs2.aabb
    small["seg_header_64x64x64"] = hs2.voxelize("segments", 64, 64, 64)[0]
    small["ver_header_64x64x64"] = hs2.voxelize("vertices", 64, 64, 64)[0]
    # variable segment counts per strand (has_segments route)
    rng = np.random.default_rng(7)
    segs = rng.integers(1, 9, size=300).astype(np.uint16)
    pieces = []
    for k, c in enumerate(segs):
        pieces.append(synth.strands(1, int(c), seed=1000 + k, seg_len=1.3))
    vv = np.concatenate(pieces, axis=0)
    hs3 = R.create(vv, len(segs), 0, segments=segs)
    small["var_vertices"] = vv
    small["var_segments"] = segs
    small["var_aabb"] = hs3.aabb
    small["var_indices"] = hs3.indices
    small["var_seg_32x32x32"] = hs3.voxelize("segments", 32, 32, 32)[0]
    np.savez_compressed(os.path.join(HERE, "small_sets.npz"), **small)

    # ---- 3. full-size fingerprints ---------------------------------------------------------
    fp = {}

    def big(name, shape, res, seed=0x5EED, seg_len=0.5, scale=1.0, modes=("segments", "vertices")):
        vb, nb, sb = synth.shape(shape, seed=seed, seg_len=seg_len, scale=scale)
        h = R.create(vb, nb, sb)
        e = {"shape": shape, "seed": seed, "seg_len": seg_len, "scale": scale, "strands": nb,
             "segments_per_strand": sb, "resolution": list(res), "input_fnv": fnv(vb),
             "aabb": [float(x) for x in h.aabb]}
        for m in modes:
            d, _, sec = h.voxelize(m, *res)
            e[m] = stats(d)
            e[m]["reference_seconds"] = round(sec, 3)
            if m == "segments":
                e["normalize_segments"] = stats(R.normalize(d))
        fp[name] = e
        print(name, json.dumps(e)[:300], flush=True)

    big("ponytail_256", "ponytail", (256, 256, 256))
    big("ponytail_long_256", "ponytail", (256, 256, 256), seg_len=2.5)
    big("ponytail_sat_32", "ponytail", (32, 32, 32), seg_len=1.0)                    # heavy 255-saturation
    big("ponytail_noncubic", "ponytail", (256, 128, 64))
    big("straight_512", "straight", (512, 512, 512), scale=0.25)                     # fp32-index rounding (F2)
    out_json["fingerprints"] = fp

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out_json, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"), "and small_sets.npz")


if __name__ == "__main__":
    main()
