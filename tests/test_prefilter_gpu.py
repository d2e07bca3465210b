"""GPU parity of the density -> AO / opacity / Gaussian prefilter (TMA-staged 3-D stencil) against the CPU restatement
of the reference's GLSL (oracle/prefilter_oracle.c).  Tolerance: 1e-6 relative (BASELINE.json north_star: "derived
float volumes must match within 1e-6 relative"); everything before the final powf is the same fp32 operation sequence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from vkhr_b200 import capi
from harness import synth

REL_TOL = 1e-6


def _close(got, want, what):
    assert got.shape == want.shape
    assert np.all(np.isfinite(got)), what
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    rel = err / np.maximum(np.abs(want.astype(np.float64)), 1e-30)
    bad = (rel > REL_TOL) & (err > 1e-37)
    assert not bad.any(), f"{what}: {bad.sum()} voxels off, worst rel {rel[bad].max():.3e}"
    return float(rel[err > 0].max()) if (err > 0).any() else 0.0


def _noise(rng, n, fill):
    d = (rng.random(n) < fill) * rng.integers(1, 256, n)
    return d.astype(np.uint8)


def _hair(vox, port, W, H, D, seed=3):
    v, n, s = synth.shape("ponytail", seed=seed, seg_len=1.0, scale=0.03)
    lo, hi = port.generate_bounding_box(v)
    return vox.voxelize_segments(v, None, lo, (hi - lo).astype(np.float32), W, H, D, segs_per_strand=s)


@pytest.mark.parametrize("res", [(64, 32, 16), (32, 32, 32), (48, 40, 24), (16, 8, 8), (30, 20, 10), (4, 4, 4), (33, 7, 5)])
@pytest.mark.parametrize("radius", [2.5, 2.0, 1.3, 0.0, 0.5])
def test_ao_against_oracle(vox, port, res, radius):
    W, H, D = res
    rng = np.random.default_rng(W * 131 + int(radius * 10))
    for d in (_noise(rng, W * H * D, 0.15), _noise(rng, W * H * D, 0.9), _hair(vox, port, W, H, D)):
        want = port.prefilter_ao(d, W, H, D, radius=radius)
        got = vox.prefilter(d, W, H, D, ao=True, ao_radius=radius)["ao"]
        _close(got, want, f"ao {res} r={radius}")
        gen = vox.prefilter(d, W, H, D, ao=True, ao_radius=radius, flags=capi.PREFILTER_GENERIC)["ao"]
        assert np.array_equal(gen, got), "tiled and generic kernels must agree bit for bit"


@pytest.mark.parametrize("res", [(64, 32, 16), (48, 40, 24), (32, 24, 11)])
@pytest.mark.parametrize("radius", [0.0, 0.5, 1.0, 1.3, 2.0, 2.5, 3.0, 3.75, 4.0, 4.5, 6.0])
@pytest.mark.parametrize("gauss_width", [None, 9.0])
def test_ao_column_form_equals_row_wise_and_generic(vox, port, res, radius, gauss_width):
    """The register-tiled z-column AO (one instantiation per tap-offset pair up to radius 4) against the row-wise tiled
    kernel and the generic kernel, bit for bit -- also when a wider Gaussian sets the tile's halo (9^3: halo 4)."""
    W, H, D = res
    rng = np.random.default_rng(W * 17 + int(radius * 100))
    for d in (_noise(rng, W * H * D, 0.02), _noise(rng, W * H * D, 0.5), _hair(vox, port, W, H, D)):
        kw = dict(ao=True, ao_radius=radius)
        if gauss_width:
            kw.update(gauss=True, gauss_width=gauss_width, opacity=True)
        col = vox.prefilter(d, W, H, D, **kw)
        row = vox.prefilter(d, W, H, D, flags=capi.PREFILTER_ROWWISE, **kw)
        gen = vox.prefilter(d, W, H, D, flags=capi.PREFILTER_GENERIC, **kw)
        for k in col:
            assert np.array_equal(col[k], row[k]), f"{k}: column vs row-wise, r={radius}"
            assert np.array_equal(col[k], gen[k]), f"{k}: column vs generic, r={radius}"
    _close(col["ao"], port.prefilter_ao(d, W, H, D, radius=radius), f"ao r={radius}")


@pytest.mark.parametrize("radius,exponent,ao_max", [(5.75, 10.0, 0.16), (7.25, 3.0, 0.4), (8.0, 32.0, 0.05), (8.5, 1.0, 0.2), (12.0, 10.0, 0.16)])
def test_ao_large_radii_and_ui_ranges(vox, port, radius, exponent, ao_max):
    """The UI ranges (interface.cc:291-293): radius 0..8, exponent 0..32, clamp 0..0.4; halo > 8 takes the generic kernel."""
    W, H, D = 48, 24, 24
    rng = np.random.default_rng(int(radius * 100))
    d = _noise(rng, W * H * D, 0.3)
    want = port.prefilter_ao(d, W, H, D, radius=radius, exponent=exponent, ao_max=ao_max)
    got = vox.prefilter(d, W, H, D, ao=True, ao_radius=radius, ao_exponent=exponent, ao_max=ao_max)["ao"]
    _close(got, want, f"ao r={radius}")


@pytest.mark.parametrize("res", [(64, 32, 16), (30, 20, 10), (4, 4, 4)])
@pytest.mark.parametrize("width", [1, 3, 5, 7, 9])
def test_gauss_and_opacity_against_oracle(vox, port, res, width):
    W, H, D = res
    rng = np.random.default_rng(W + width)
    d = _noise(rng, W * H * D, 0.3)
    out = vox.prefilter(d, W, H, D, ao=True, opacity=True, gauss=True, gauss_width=float(width), strand_alpha=0.35, thickness=11.0)
    _close(out["gauss"], port.prefilter_gauss(d, W, H, D, float(width)), f"gauss {res} N={width}")
    _close(out["opacity"], port.prefilter_opacity(d, 0.35, 11.0), f"opacity {res}")
    _close(out["ao"], port.prefilter_ao(d, W, H, D), f"ao {res} (all three outputs in one launch)")
    # opacity is a per-density table: exactly 256 distinct values at most, 1.0 where empty
    assert np.all(out["opacity"][d == 0] == 1.0)


def test_prefilter_errors(vox):
    d = np.zeros(64, dtype=np.uint8)
    with pytest.raises(capi.VkhrB200Error) as e:
        vox.prefilter(d, 4, 4, 4, gauss=True, gauss_width=4.0)
    assert e.value.code == capi.ERR_INVALID_ARGUMENT
    with pytest.raises(capi.VkhrB200Error):
        vox.prefilter(d, 4, 4, 4, ao=True, ao_radius=-1.0)
    with pytest.raises(capi.VkhrB200Error):
        vox.prefilter(d, 4, 4, 4, ao=True, ao_radius=float("nan"))
    # empty volume: AO is exactly 1, Gaussian exactly 0
    out = vox.prefilter(d, 4, 4, 4, ao=True, gauss=True)
    assert np.all(out["ao"] == 1.0) and np.all(out["gauss"] == 0.0)


def test_prefilter_full_size_device_path(vox, port):
    """256^3 (the reference's resolution) on the device API: a slab is checked against the oracle, the whole volume
    against the generic kernel bit for bit, and empty space must come out as exactly 1."""
    import torch
    dev = torch.device("cuda", 0)
    W = H = D = 256
    v, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
    lo, hi = synth.host_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    dens = vox.voxelize_segments_dev(torch.from_numpy(v).to(dev).reshape(-1), None, lo, size, W, H, D, segs_per_strand=s)
    ao = torch.empty(W * H * D, dtype=torch.float32, device=dev)
    op = torch.empty_like(ao)
    ao2 = torch.empty_like(ao)
    vox.prefilter_dev(dens, W, H, D, ao=ao, opacity=op)
    vox.prefilter_dev(dens, W, H, D, ao=ao2, flags=capi.PREFILTER_GENERIC)
    torch.cuda.synchronize()
    assert torch.equal(ao, ao2)
    dh = dens.cpu().numpy()
    # oracle on a z-slab with its halo (the slab's interior does not see the cut)
    z0, z1, hz = 120, 136, 3
    sub = dh.reshape(D, H, W)[z0 - hz:z1 + hz].reshape(-1)
    want = port.prefilter_ao(sub, W, H, z1 - z0 + 2 * hz).reshape(-1, H, W)[hz:-hz]
    got = ao.cpu().numpy().reshape(D, H, W)[z0:z1]
    _close(got, want, "ao 256^3 slab")
    _close(op.cpu().numpy(), port.prefilter_opacity(dh), "opacity 256^3")
    a = ao.cpu().numpy()
    assert a.max() == 1.0 and a.min() > 0.17            # (1 - 0.16)^10 is the floor
    # a voxel whose whole footprint is empty is exactly 1
    occ = torch.nn.functional.max_pool3d((dens.reshape(1, 1, D, H, W) > 0).float(), 7, 1, 3).reshape(-1) > 0
    assert torch.all(ao[~occ] == 1.0)


@pytest.mark.parametrize("res", [(128, 64, 48), (64, 64, 64), (96, 40, 24)])
def test_sparse_volumes_skip_empty_tiles_bit_identically(vox, res):
    """The occupancy pre-pass of the tiled kernel (cells of 32 x 8 x 8 voxels -> one activity byte per tile): tiles whose box
    holds no hair are never loaded and get the constants of empty space.  Mostly empty volumes with hair in a few places
    (a corner, a face, one voxel, a thin sheet) must come out bit-identical to the dense form (every tile loaded) and to
    the generic kernel, for AO, opacity and a Gaussian whose halo is wider than the AO's."""
    W, H, D = res
    rng = np.random.default_rng(W + 3 * H)
    vols = []
    d = np.zeros((D, H, W), np.uint8); d[0, 0, 0] = 200; d[D - 1, H - 1, W - 1] = 7; vols.append(d)
    d = np.zeros((D, H, W), np.uint8); d[D // 2, H // 3, 5:W - 9] = rng.integers(1, 255, W - 14); vols.append(d)
    d = np.zeros((D, H, W), np.uint8); d[7:9, :, 31:33] = 255; d[15:17, 8, :] = 3; vols.append(d)           # straddles cell / tile faces
    d = np.zeros((D, H, W), np.uint8); vols.append(d)                                                        # nothing at all
    d = (rng.random((D, H, W)) < 0.0005).astype(np.uint8) * 90; vols.append(d)
    for k, d in enumerate(vols):
        d = d.reshape(-1)
        kw = dict(ao=True, opacity=True, gauss=True, gauss_width=9.0)
        got = vox.prefilter(d, W, H, D, **kw)
        dense = vox.prefilter(d, W, H, D, flags=capi.PREFILTER_DENSE, **kw)
        gen = vox.prefilter(d, W, H, D, flags=capi.PREFILTER_GENERIC, **kw)
        for name in ("ao", "opacity", "gauss"):
            assert np.array_equal(got[name], dense[name]), (k, name, "sparse vs dense")
            assert np.array_equal(got[name], gen[name]), (k, name, "tiled vs generic")


# ---- volumetric ADSM transmittance volume (approximate_deep_shadows.glsl:24-36 at every voxel centre) ----------------
@pytest.mark.parametrize("res", [(32, 32, 32), (48, 20, 12), (7, 5, 3), (4, 4, 4)])
@pytest.mark.parametrize("light", [(30.0, 80.0, 20.0), (-15.0, 3.0, 2.0), (2.0, 2.5, 1.0), (1e4, -2e4, 5e3)])
def test_adsm_against_oracle(vox, port, res, light):
    """Light far outside, close by, INSIDE the volume, and very far: the march is clipped differently each time."""
    W, H, D = res
    rng = np.random.default_rng(W * 7 + H)
    origin, size = np.array([-1.0, 0.5, -2.0], np.float32), np.array([6.0, 5.0, 4.0], np.float32)
    for d in (_noise(rng, W * H * D, 0.1), _noise(rng, W * H * D, 0.6) // 8):
        want = port.prefilter_adsm(d, W, H, D, origin, size, light)
        got = vox.adsm(d, W, H, D, origin, size, light)
        _close(got, want, f"adsm {res} light={light}")


@pytest.mark.parametrize("steps,alpha,thickness", [(1024.0, 0.3, 11.0), (100.0, 0.3, 11.0), (333.0, 0.05, 2.0), (64.0, 0.9, 0.5), (1.0, 0.3, 11.0)])
def test_adsm_parameters(vox, port, steps, alpha, thickness):
    """steps = 100 takes 101 samples (the fp32 accumulation of t in the shader's loop): the t table reproduces it."""
    W, H, D = 24, 16, 20
    rng = np.random.default_rng(int(steps))
    d = _noise(rng, W * H * D, 0.2) // 4
    origin, size, light = [0.0, 0.0, 0.0], [3.0, 2.0, 2.5], [4.0, 9.0, -3.0]
    want = port.prefilter_adsm(d, W, H, D, origin, size, light, steps=steps, strand_alpha=alpha, thickness=thickness)
    got = vox.adsm(d, W, H, D, origin, size, light, steps=steps, strand_alpha=alpha, thickness=thickness)
    _close(got, want, f"adsm steps={steps}")


def test_adsm_of_a_voxelised_style_and_properties(vox, port):
    """On real (synthetic-hair) densities: parity, an empty volume is fully lit, and more hair never lets more light through."""
    W = H = D = 40
    v, n, s = synth.shape("ponytail", seed=5, seg_len=1.0, scale=0.03)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    d = vox.voxelize_segments(v, None, lo, size, W, H, D, segs_per_strand=s)
    light = (lo + size * np.array([0.5, 3.0, 0.5], np.float32))
    got = vox.adsm(d, W, H, D, lo, size, light)
    _close(got, port.prefilter_adsm(d, W, H, D, lo, size, light), "adsm hair")
    assert np.all(vox.adsm(np.zeros_like(d), W, H, D, lo, size, light) == 1.0)
    denser = vox.adsm(np.minimum(d.astype(np.int32) * 2, 255).astype(np.uint8), W, H, D, lo, size, light)
    assert np.all(denser <= got)


def test_adsm_device_resident_matches_host_api(vox):
    import torch
    W, H, D = 32, 16, 8
    rng = np.random.default_rng(9)
    d = _noise(rng, W * H * D, 0.3)
    host = vox.adsm(d, W, H, D, [0, 0, 0], [4, 2, 1], [5, 6, 7])
    dev = vox.adsm_dev(torch.from_numpy(d).cuda(), W, H, D, [0, 0, 0], [4, 2, 1], [5, 6, 7])
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)


def test_cuda_against_the_committed_fixtures(vox):
    """The CUDA prefilter / ADSM against tests/golden/prefilter_fixtures.npz (frozen outputs of the restatement, see
    tests/golden/make_prefilter_fixtures.py) -- no live oracle call on this path: 1e-6 relative."""
    import importlib.util, os
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_prefilter_fixtures", os.path.join(g, "make_prefilter_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fx = np.load(os.path.join(g, "prefilter_fixtures.npz"))
    for name, W, H, D, _, _ in mod.VOLUMES:
        d = fx[f"{name}/densities"]
        for k, (r, e, m) in enumerate(mod.AO_PARAMS):
            _close(vox.prefilter(d, W, H, D, ao=True, ao_radius=r, ao_exponent=e, ao_max=m)["ao"], fx[f"{name}/ao{k}"], f"{name} ao{k}")
        for k, w in enumerate(mod.GAUSS_WIDTHS):
            _close(vox.prefilter(d, W, H, D, ao=False, gauss=True, gauss_width=w)["gauss"], fx[f"{name}/gauss{k}"], f"{name} gauss{k}")
        for k, (a, t) in enumerate(mod.OPACITY_PARAMS):
            _close(vox.prefilter(d, W, H, D, ao=False, opacity=True, strand_alpha=a, thickness=t)["opacity"], fx[f"{name}/opacity{k}"], f"{name} opacity{k}")
        for k, (o, s, l, steps, a, t) in enumerate(mod.ADSM_CASES):
            _close(vox.adsm(d, W, H, D, o, s, l, steps=steps, strand_alpha=a, thickness=t), fx[f"{name}/adsm{k}"], f"{name} adsm{k}")
