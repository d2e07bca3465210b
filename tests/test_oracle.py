"""The CPU restatement (oracle/voxel_oracle.c) against the pins: golden vectors generated from the
unmodified reference (tests/golden/make_golden.py) and, where the reference driver is available,
the reference itself on fresh inputs.  No GPU involved."""
import os
import numpy as np
import pytest

from conftest import dense

from harness import synth


def _fnv(port, a):
    return f"{port.fnv1a64(a):016x}"


def test_kat4_appendix_b(port, golden):
    k = golden["kat4"]
    v = np.array(k["vertices"], dtype=np.float32)
    lo, hi = port.generate_bounding_box(v)
    bb = port.get_bounding_box(lo, hi)
    assert np.allclose(bb, k["aabb"], rtol=0, atol=0)
    idx = port.generate_indices(k["strands"], k["segments_per_strand"])
    assert idx.tolist() == k["indices"]
    tg = port.generate_tangents(v, k["strands"], k["segments_per_strand"])
    assert np.array_equal(tg, np.array(k["tangents_in"], dtype=np.float32), equal_nan=True)
    d, t = port.voxelize_segments(v, idx, bb[:3], bb[4:7], 4, 4, 4, tangents=tg)
    assert np.array_equal(d, dense(k["voxelize_segments"], 64))
    # the hand-verified table of SURVEY.md Appendix B
    assert {i: int(d[i]) for i in np.nonzero(d)[0]} == {0: 1, 12: 1, 13: 1, 14: 1, 21: 1, 42: 1, 63: 3}
    for i, tv in k["voxelize_segments_tangents"].items():
        assert t[int(i)].tolist() == tv
    dv, tv_ = port.voxelize_vertices(v, bb[:3], bb[4:7], 4, 4, 4, tangents=tg)
    assert np.array_equal(dv, dense(k["voxelize_vertices"], 64))
    assert {i: int(dv[i]) for i in np.nonzero(dv)[0]} == {0: 1, 12: 1, 15: 1, 21: 2, 47: 1, 63: 2}
    for i, tv in k["voxelize_vertices_tangents"].items():
        assert tv_[int(i)].tolist() == tv
    assert np.array_equal(port.normalize(d), dense(k["normalize_segments"], 64))
    assert port.downsample(d, 4, 4, 4, 2).tolist() == k["downsample_sum_segments"] == [2, 0, 2, 1, 0, 0, 0, 4]
    assert port.downsample(d, 4, 4, 4, 0).tolist() == k["downsample_max_segments"]


@pytest.mark.parametrize("res", [(64, 64, 64), (64, 32, 16), (16, 16, 16), (30, 20, 10)])
def test_small_sets(port, small_sets, res):
    W, H, D = res
    tag = f"{W}x{H}x{D}"
    v = small_sets["in_vertices"]
    n, s = [int(x) for x in small_sets["in_meta"]]
    bb = small_sets["aabb_generated"]
    lo, hi = port.generate_bounding_box(v)
    assert np.array_equal(port.get_bounding_box(lo, hi), bb)
    idx = port.generate_indices(n, s)
    tg = port.generate_tangents(v, n, s)
    d, t = port.voxelize_segments(v, idx, bb[:3], bb[4:7], W, H, D, tangents=tg)
    assert np.array_equal(d, small_sets[f"seg_{tag}"])
    # tangents: the reference's fp32 sum is reproduced op for op on one thread => exact here
    assert np.array_equal(t, small_sets[f"segtan_{tag}"])
    assert np.array_equal(port.voxelize_segments(v, idx, bb[:3], bb[4:7], W, H, D), d)   # density-only walk
    assert np.array_equal(port.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D), small_sets[f"ver_{tag}"])
    assert np.array_equal(port.normalize(d), small_sets[f"segnorm_{tag}"])
    if f"segdown0_{tag}" in small_sets:
        for f in range(4):
            assert np.array_equal(port.downsample(d, W, H, D, f), small_sets[f"segdown{f}_{tag}"])
    # counts before the clamp: min(count, 255) == density
    c = port.count_segments(v, idx, bb[:3], bb[4:7], W, H, D)
    assert np.array_equal(np.minimum(c, 255).astype(np.uint8), d)
    assert int(c.sum()) <= port.count_samples(v, idx, bb[:3], bb[4:7], W, H, D)


def test_small_header_aabb_and_variable_strands(port, small_sets):
    v = small_sets["in_vertices"]
    n, s = [int(x) for x in small_sets["in_meta"]]
    bb = small_sets["aabb_header"]
    idx = port.generate_indices(n, s)
    assert np.array_equal(port.voxelize_segments(v, idx, bb[:3], bb[4:7], 64, 64, 64), small_sets["seg_header_64x64x64"])
    assert np.array_equal(port.voxelize_vertices(v, bb[:3], bb[4:7], 64, 64, 64), small_sets["ver_header_64x64x64"])
    vv, segs, vbb = small_sets["var_vertices"], small_sets["var_segments"], small_sets["var_aabb"]
    vidx = port.generate_indices(len(segs), 0, segments=segs)
    assert np.array_equal(vidx, small_sets["var_indices"])
    assert np.array_equal(port.voxelize_segments(vv, vidx, vbb[:3], vbb[4:7], 32, 32, 32), small_sets["var_seg_32x32x32"])


@pytest.mark.parametrize("name", ["ponytail_256", "ponytail_long_256", "ponytail_sat_32", "ponytail_noncubic", "straight_512"])
def test_full_size_fingerprints(port, golden, name):
    e = golden["fingerprints"][name]
    v, n, s = synth.shape(e["shape"], seed=e["seed"], seg_len=e["seg_len"], scale=e["scale"])
    assert _fnv(port, v) == e["input_fnv"], "synthetic generator output changed: regenerate tests/golden"
    W, H, D = e["resolution"]
    bb = np.array(e["aabb"], dtype=np.float32)
    idx = port.generate_indices(n, s)
    d = port.voxelize_segments(v, idx, bb[:3], bb[4:7], W, H, D)
    st = e["segments"]
    assert (_fnv(port, d), int(d.astype(np.int64).sum()), int(np.count_nonzero(d)), int((d == 255).sum())) == \
           (st["fnv"], st["sum"], st["nonzero"], st["saturated"])
    assert _fnv(port, port.normalize(d)) == e["normalize_segments"]["fnv"]
    dv = port.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D)
    assert _fnv(port, dv) == e["vertices"]["fnv"]
    if name == "ponytail_sat_32":
        assert st["saturated"] > 0 and st["max"] == 255
    if name == "straight_512":
        # an exact integer index differs from the reference's fp32 index above 2^24 voxels
        de = port.voxelize_segments(v, idx, bb[:3], bb[4:7], W, H, D, flags=1)
        assert not np.array_equal(de, d)
        assert int(np.minimum(port.count_segments(v, idx, bb[:3], bb[4:7], W, H, D, flags=1), 255).sum()) == int(de.astype(np.int64).sum())


def _random_case(rng, n_strands, segs, spread):
    v = synth.strands(n_strands, segs, seed=int(rng.integers(1, 2**31)), seg_len=float(spread),
                      curl=float(rng.uniform(0.1, 2.0)), gravity=float(rng.uniform(0, 0.6)))
    return v


@pytest.mark.parametrize("seed", range(6))
def test_port_matches_live_reference(port, ref, seed):
    """Fresh seeded inputs through the unmodified reference and through the restatement."""
    rng = np.random.default_rng(100 + seed)
    n, s = int(rng.integers(50, 600)), int(rng.integers(1, 20))
    v = _random_case(rng, n, s, rng.uniform(0.2, 6.0))
    W, H, D = [int(x) for x in rng.choice([4, 7, 16, 33, 64, 100], size=3)]
    if seed % 2:
        hs = ref.create(v, n, s)
    else:                                   # header AABB slightly larger than the data
        lo, hi = v.min(axis=0) - 0.5, v.max(axis=0) + 0.25
        hs = ref.create(v, n, s, aabb_min=lo, aabb_max=hi)
    bb = hs.aabb
    idx, tg = hs.indices, hs.tangents
    assert np.array_equal(port.generate_indices(n, s), idx)
    assert np.array_equal(port.generate_tangents(v, n, s), tg, equal_nan=True)
    d_ref, t_ref, _ = hs.voxelize("segments", W, H, D, want_tangents=True)
    d, t = port.voxelize_segments(v, idx, bb[:3], bb[4:7], W, H, D, tangents=tg)
    assert np.array_equal(d, d_ref)
    assert np.array_equal(t, t_ref)
    dv_ref, tv_ref, _ = hs.voxelize("vertices", W, H, D, want_tangents=True)
    dv, tv = port.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D, tangents=tg)
    assert np.array_equal(dv, dv_ref) and np.array_equal(tv, tv_ref)
    assert np.array_equal(port.normalize(d), ref.normalize(d_ref))
    if W % 2 == 0 and H % 2 == 0 and D % 2 == 0:
        for f in range(4):
            assert np.array_equal(port.downsample(d, W, H, D, f), ref.downsample(d_ref, W, H, D, f))


def test_edge_cases_against_reference(port, ref):
    """Max-corner clamp, zero-length segments, vertices on voxel faces, a saturating clump."""
    rng = np.random.default_rng(5)
    clump = np.tile(np.array([[1.25, 1.25, 1.25], [1.75, 2.6, 1.3]], dtype=np.float32), (400, 1))   # 400 x same segment
    faces = np.array([[0, 0, 0], [8, 8, 8], [8, 8, 8], [8, 0, 8], [2, 2, 2], [2, 2, 2], [3, 3, 3], [3, 4, 3],
                      [7.999999, 7.999999, 7.999999], [8, 8, 8], [0, 8, 0], [8, 8, 0]], dtype=np.float32)
    jitter = rng.uniform(0, 8, size=(200, 3)).astype(np.float32)
    v = np.concatenate([clump, faces, jitter], axis=0)
    n = v.shape[0] // 2
    hs = ref.create(v, n, 1)
    bb = hs.aabb
    for res in [(8, 8, 8), (4, 4, 4), (5, 3, 2)]:
        d_ref, t_ref, _ = hs.voxelize("segments", *res, want_tangents=True)
        d, t = port.voxelize_segments(v, hs.indices, bb[:3], bb[4:7], *res, tangents=hs.tangents)
        assert np.array_equal(d, d_ref) and d.max() == 255
        sat = d == 255
        assert np.array_equal(t[~sat], t_ref[~sat])
        assert np.array_equal(port.voxelize_vertices(v, bb[:3], bb[4:7], *res), hs.voxelize("vertices", *res)[0])


def test_defined_extensions(port):
    """Inputs on which the reference is undefined: the rules both the oracle and the CUDA path implement."""
    origin, size = np.zeros(3, np.float32), np.full(3, 4.0, np.float32)
    # a vertex outside the box: index out of range => dropped; in-range wrap-around is kept like the reference
    v = np.array([[9, 9, 9], [9.5, 9, 9], [1, 1, 1], [1, 1, 1]], dtype=np.float32)
    idx = np.array([0, 1, 2, 3], dtype=np.uint32)
    assert port.voxelize_segments(v, idx, origin, size, 4, 4, 4).sum() == 1      # (9,9,9) clamps to voxel 63
    v = np.array([[-9, -9, -9], [-9.5, -9, -9]], dtype=np.float32)
    assert port.voxelize_segments(v, idx[:2], origin, size, 4, 4, 4).sum() == 0  # negative index: dropped
    v = np.array([[np.nan, 0, 0], [1, 1, 1]], dtype=np.float32)
    assert port.voxelize_segments(v, idx[:2], origin, size, 4, 4, 4).sum() == 0
    assert port.voxelize_vertices(v, origin, size, 4, 4, 4).sum() == 1
    # fewer than two indices => empty volume; odd trailing index ignored
    v = np.array([[0, 0, 0], [3, 0, 0], [0, 3, 0]], dtype=np.float32)
    assert port.voxelize_segments(v, np.array([0], np.uint32), origin, size, 4, 4, 4).sum() == 0
    assert port.voxelize_segments(v, np.array([0, 1, 2], np.uint32), origin, size, 4, 4, 4).sum() == 3
    # normalize with max == min leaves the grid unchanged
    flat = np.full(64, 7, dtype=np.uint8)
    assert np.array_equal(port.normalize(flat), flat)


# ---- prefilter oracle (oracle/prefilter_oracle.c) -------------------------------------------------------------
def _shader_lao_f64(d, W, H, D, origin, size, radius, exponent, ao_max):
    """local_ambient_occlusion.glsl:9-30 + sample_volume.glsl:7-9 written out literally in float64 THROUGH WORLD
    SPACE (fragment_position = voxel centre), with LINEAR / CLAMP_TO_BORDER(0) / R8_UNORM sampling per the Vulkan
    spec (unnormalised coordinate u*size - 0.5, floor + fraction).  Independent of the oracle's texel-space shortcut."""
    vol = np.zeros((D + 2, H + 2, W + 2))
    vol[1:-1, 1:-1, 1:-1] = d.reshape(D, H, W) / 255.0                      # one texel of black border
    res = np.array([W, H, D], dtype=np.float64)
    origin, size = np.asarray(origin, np.float64), np.asarray(size, np.float64)
    k, j, i = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    centre = origin + (np.stack([i, j, k], -1) + 0.5) * (size / res)
    kernel_size = 2.0
    kernel_radius = (kernel_size - 1.0) / 2.0
    voxel_sample_scaling = (radius / kernel_radius) * (size / res)

    def texture(pos):
        t = (pos - origin) / size * res - 0.5
        t0 = np.floor(t)
        f = t - t0
        out = np.zeros(pos.shape[:-1])
        for cz in (0, 1):
            for cy in (0, 1):
                for cx in (0, 1):
                    x = np.clip(t0[..., 0] + cx, -1, W).astype(int) + 1
                    y = np.clip(t0[..., 1] + cy, -1, H).astype(int) + 1
                    z = np.clip(t0[..., 2] + cz, -1, D).astype(int) + 1
                    w = (f[..., 0] if cx else 1 - f[..., 0]) * (f[..., 1] if cy else 1 - f[..., 1]) * (f[..., 2] if cz else 1 - f[..., 2])
                    out += w * vol[z, y, x]
        return out

    density = np.zeros((D, H, W))
    for z in (-kernel_radius, kernel_radius):
        for y in (-kernel_radius, kernel_radius):
            for x in (-kernel_radius, kernel_radius):
                density += np.minimum(texture(centre + np.array([x, y, z]) * voxel_sample_scaling), ao_max)
    return (1.0 - density / kernel_size ** 3) ** exponent


@pytest.mark.parametrize("radius", [2.5, 1.25, 3.0, 0.75])
def test_prefilter_ao_oracle_matches_the_shader_text(port, radius):
    rng = np.random.default_rng(int(radius * 8))
    W, H, D = 12, 10, 9
    d = ((rng.random(W * H * D) < 0.3) * rng.integers(1, 256, W * H * D)).astype(np.uint8)
    got = port.prefilter_ao(d, W, H, D, radius=radius, exponent=10.0, ao_max=0.16).reshape(D, H, W)
    # power-of-two voxel size and origin: the world-space detour is exact in float64 and even integer radii
    # (a tap exactly on a texel centre) land on the same side as the oracle's texel-space rule
    want = _shader_lao_f64(d, W, H, D, origin=(-4.0, 8.0, 16.0), size=(W * 0.25, H * 0.5, D * 0.125), radius=radius,
                           exponent=10.0, ao_max=0.16)
    assert np.allclose(got, want, rtol=2e-5, atol=0)


def test_prefilter_gauss_and_opacity_oracle(port):
    rng = np.random.default_rng(4)
    W, H, D = 9, 8, 7
    d = ((rng.random(W * H * D) < 0.4) * rng.integers(1, 256, W * H * D)).astype(np.uint8)
    for N in (1, 3, 5):
        got = port.prefilter_gauss(d, W, H, D, float(N)).reshape(D, H, W)
        R = (N - 1) // 2
        sigma2 = ((N / 2.0) / 2.4) ** 2
        pad = np.zeros((D + 2 * R, H + 2 * R, W + 2 * R))
        pad[R:R + D, R:R + H, R:R + W] = d.reshape(D, H, W) / 255.0
        acc, tot = np.zeros((D, H, W)), 0.0
        for z in range(-R, R + 1):
            for y in range(-R, R + 1):
                for x in range(-R, R + 1):
                    w = 1.0 / (2.0 * np.pi * sigma2) * np.e ** (-1.0 * (x * x + y * y + z * z) / 2.0 * sigma2)   # the quirk: * sigma2
                    acc += w * pad[R + z:R + z + D, R + y:R + y + H, R + x:R + x + W]
                    tot += w
        assert np.allclose(got, acc / tot, rtol=2e-5, atol=1e-7)
    op = port.prefilter_opacity(d, 0.3, 11.0)
    assert np.allclose(op, (1 - 0.3) ** (d / 255.0 * 11.0), rtol=2e-6)


# ---- the ADSM restatement (oracle/prefilter_oracle.c): hand-checkable cases -------------------------------------------
def test_adsm_oracle_known_answers(port):
    """One texel of density 255 (tau = 1), light straight above at ten voxel sizes: the march samples tau * (1 - 10 t) while
    10 t < 1, i.e. for t_k = k / 1024, k = 0..102, so strands = thickness * (103 - 10 * (102 * 103 / 2) / 1024)."""
    d = np.array([255], dtype=np.uint8)
    got = port.prefilter_adsm(d, 1, 1, 1, [0, 0, 0], [1, 1, 1], [0.5, 0.5, 10.5], steps=1024.0, strand_alpha=0.3, thickness=0.1)
    strands = 0.1 * (103 - 10 * (102 * 103 / 2) / 1024)
    assert abs(float(got[0]) - 0.7 ** strands) <= 1e-4 * 0.7 ** strands
    # an empty volume lets everything through; the t sequence is the shader's fp32 accumulation (100 steps -> 101 samples)
    assert np.all(port.prefilter_adsm(np.zeros(27, np.uint8), 3, 3, 3, [0, 0, 0], [1, 1, 1], [5, 5, 5]) == 1.0)
    t = port.adsm_t_table(100.0)
    assert len(t) == 101 and t[0] == 0.0 and np.all(np.diff(t) > 0) and t[-1] < 1.0
    assert len(port.adsm_t_table(1024.0)) == 1024 and port.adsm_t_table(1024.0)[3] == np.float32(3 / 1024)
    # more hair between a voxel and the light never brightens it
    rng = np.random.default_rng(1)
    dens = (rng.random(6 * 5 * 4) < 0.3) * rng.integers(1, 100, 6 * 5 * 4)
    a = port.prefilter_adsm(dens.astype(np.uint8), 6, 5, 4, [0, 0, 0], [3, 2.5, 2], [1, 9, 1])
    b = port.prefilter_adsm((dens * 2).astype(np.uint8), 6, 5, 4, [0, 0, 0], [3, 2.5, 2], [1, 9, 1])
    assert np.all(b <= a) and a.min() >= 0 and a.max() <= 1.0 and (a < 1.0).any()


AT_SIZE = ["straight_512_full", "curly_512_full", "big_512_full", "ponytail_sway_t0_1024", "ponytail_sway_t59_1024", "ponytail_sway_t119_1024"]


def at_size_input(e):
    """The strands of a `fingerprints_at_size` entry (tests/golden/make_golden.py --sizes)."""
    v, n, s = synth.shape(e["shape"], seed=e["seed"], seg_len=e["seg_len"], scale=e["scale"])
    if "sway_t" in e:
        v = synth.sway(v, n, s, float(e["sway_t"]))
    return v, n, s


@pytest.mark.parametrize("name", AT_SIZE)
def test_port_at_the_baseline_sizes(port, golden, name):
    """BASELINE.json configs 1 (straight AND curly, 50,000 x 65 at 512^3), 2 (1 M x 32 at 512^3) and 4 (swayed ponytail
    frames in the union AABB of the sequence at 1024^3): the C restatement against fingerprints of the UNMODIFIED
    reference at those sizes -- the fp32 index rounds in groups of up to 8 (512^3) and 64 (1024^3) voxels there
    (hair_style.cc:321)."""
    e = golden["fingerprints_at_size"][name]
    v, n, s = at_size_input(e)
    assert _fnv(port, v) == e["input_fnv"], "synthetic generator output changed: regenerate tests/golden (--sizes)"
    W, H, D = e["resolution"]
    bb = np.array(e["aabb"], dtype=np.float32)
    if "sway_t" in e:
        lo, hi = np.array(e["union_aabb_min"], np.float32), np.array(e["union_aabb_max"], np.float32)
        assert np.array_equal(bb[:3], lo) and np.array_equal(bb[4:7], hi - lo)
    d = port.voxelize_segments(v, port.generate_indices(n, s), bb[:3], bb[4:7], W, H, D)
    st = e["segments"]
    assert (_fnv(port, d), int(d.astype(np.int64).sum()), int(np.count_nonzero(d)), int((d == 255).sum())) == \
           (st["fnv"], st["sum"], st["nonzero"], st["saturated"])
    assert _fnv(port, port.normalize(d)) == e["normalize_segments"]["fnv"]


# ---- the prefilter / ADSM restatement frozen against committed fixtures ------------------------------------------------
def _prefilter_fixtures():
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_prefilter_fixtures", os.path.join(sys_path, "make_prefilter_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(sys_path, "prefilter_fixtures.npz"))


def test_prefilter_oracle_is_frozen_by_committed_fixtures(port):
    """oracle/prefilter_oracle.c recomputes tests/golden/prefilter_fixtures.npz BIT FOR BIT (float32): the checker of the
    CUDA prefilter / ADSM kernels cannot drift with them.  (Oracle-generated, not reference-held: rows f1 / f3 stay
    "parity unpinned by the reference")"""
    mod, fx = _prefilter_fixtures()
    now = mod.compute(port)
    assert sorted(now) == sorted(fx.files)
    for k in fx.files:
        assert now[k].dtype == fx[k].dtype and np.array_equal(now[k], fx[k]), k


def test_prefilter_fixtures_hold_hand_computed_values():
    """A single texel of density 255 in an empty 9^3 volume, by hand from the shader text:
    AO, radius 0 (local_ambient_occlusion.glsl:9-30): all eight taps sit on the voxel's own centre, each min(1.0, 0.16), so
        ao = (1 - 8 * 0.16 / 8)^10 = 0.84^10 at the texel, 1 everywhere else;
    Gaussian of width 1: the identity, tau = 1 at the texel; opacity = (1 - 0.3)^(1 * 11) = 0.7^11 at the texel."""
    _, fx = _prefilter_fixtures()
    c = (4 * 9 + 4) * 9 + 4
    ao = fx["single_voxel/ao3"]
    assert abs(float(ao[c]) - 0.84 ** 10) <= 2e-6 * 0.84 ** 10 and np.all(np.delete(ao, c) == 1.0)
    g = fx["single_voxel/gauss0"]
    assert g[c] == 1.0 and np.all(np.delete(g, c) == 0.0)
    op = fx["single_voxel/opacity0"]
    assert abs(float(op[c]) - 0.7 ** 11) <= 2e-6 * 0.7 ** 11 and np.all(np.delete(op, c) == 1.0)
    # AO, radius 2.5 (kernel_radius = 0.5, so the eight taps sit at +-2.5 texels on every axis): the centre voxel's taps
    # fall between texels 1|2 and 6|7 -- weight 0 of texel 4 -- so it is unoccluded; voxel (6, 6, 6) has ONE tap at texel
    # coordinate 3.5 on every axis: trilinear weight 0.5^3 = 0.125 of the texel, below the clamp of 0.16, so
    # ao = (1 - 0.125 / 8)^10 = 0.984375^10
    ao25 = fx["single_voxel/ao0"]
    assert ao25[c] == 1.0
    n = (6 * 9 + 6) * 9 + 6
    assert abs(float(ao25[n]) - 0.984375 ** 10) <= 2e-6
