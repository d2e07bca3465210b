"""Generate tests/golden/prefilter_fixtures.npz: outputs of the prefilter / ADSM restatement (oracle/prefilter_oracle.c) on
fixed inputs, committed so that the restatement cannot drift together with the CUDA kernels it checks.

    python tests/golden/make_prefilter_fixtures.py

These fixtures are ORACLE-generated, not reference-held: the reference's consumers are GLSL fragment/compute shaders
(share/shaders/volumes/local_ambient_occlusion.glsl:9-30, sample_volume.glsl:12-35, approximate_deep_shadows.glsl:24-36)
and there is no GLSL toolchain or Vulkan device in this container -- rows f1 / f3 stay "parity unpinned by the reference"
(DESIGN.md section 3).  What IS independent of the restatement: tests/test_oracle.py holds it to a float64 numpy reading
of the shader text (AO), to the closed form of the Gaussian with its `* sigma2` quirk, to (1 - alpha)^(tau * thickness)
and to hand-computed ADSM marches; this file freezes its float32 outputs bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

OUT = os.path.join(HERE, "prefilter_fixtures.npz")

# (name, W, H, D, fill, seed)
VOLUMES = [("noise_sparse", 20, 12, 9, 0.08, 11), ("noise_dense", 16, 16, 16, 0.7, 12), ("slab", 24, 8, 10, None, 13),
           ("single_voxel", 9, 9, 9, None, 14)]
AO_PARAMS = [(2.5, 10.0, 0.16), (1.25, 4.0, 0.5), (3.0, 10.0, 0.16), (0.0, 10.0, 0.16)]      # radius, exponent, ao_max (volume.frag UI ranges)
GAUSS_WIDTHS = [1.0, 3.0, 5.0]
OPACITY_PARAMS = [(0.3, 11.0), (0.05, 2.0)]
ADSM_CASES = [((-1.0, 0.5, -2.0), (6.0, 5.0, 4.0), (30.0, 80.0, 20.0), 1024.0, 0.3, 11.0),
              ((-1.0, 0.5, -2.0), (6.0, 5.0, 4.0), (2.0, 2.5, 1.0), 100.0, 0.3, 11.0),     # light inside the volume, 101 samples
              ((0.0, 0.0, 0.0), (3.0, 2.0, 2.5), (4.0, 9.0, -3.0), 333.0, 0.05, 2.0)]


def volume(name, W, H, D, fill, seed):
    rng = np.random.default_rng(seed)
    if name == "slab":
        d = np.zeros((D, H, W), np.uint8)
        d[3:6, :, 5:17] = rng.integers(1, 256, (3, H, 12))
        return d.reshape(-1)
    if name == "single_voxel":
        d = np.zeros((D, H, W), np.uint8)
        d[4, 4, 4] = 255
        return d.reshape(-1)
    return ((rng.random(W * H * D) < fill) * rng.integers(1, 256, W * H * D)).astype(np.uint8)


def compute(port):
    out = {}
    for name, W, H, D, fill, seed in VOLUMES:
        d = volume(name, W, H, D, fill, seed)
        out[f"{name}/densities"] = d
        out[f"{name}/res"] = np.array([W, H, D], np.int32)
        for k, (r, e, m) in enumerate(AO_PARAMS):
            out[f"{name}/ao{k}"] = port.prefilter_ao(d, W, H, D, radius=r, exponent=e, ao_max=m)
        for k, w in enumerate(GAUSS_WIDTHS):
            out[f"{name}/gauss{k}"] = port.prefilter_gauss(d, W, H, D, w)
        for k, (a, t) in enumerate(OPACITY_PARAMS):
            out[f"{name}/opacity{k}"] = port.prefilter_opacity(d, a, t)
        for k, (o, s, l, steps, a, t) in enumerate(ADSM_CASES):
            out[f"{name}/adsm{k}"] = port.prefilter_adsm(d, W, H, D, o, s, l, steps=steps, strand_alpha=a, thickness=t)
    return out


if __name__ == "__main__":
    np.savez_compressed(OUT, **compute(oracle.port()))
    print(OUT, os.path.getsize(OUT), "bytes")
