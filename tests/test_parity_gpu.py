"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the golden vectors generated
from the unmodified reference, against the CPU oracle on seeded inputs, and -- at BASELINE.json's
full sizes -- through size-independent properties.  Densities are compared BIT-EXACT."""
import numpy as np
import pytest

from conftest import dense

pytestmark = pytest.mark.gpu

import vkhr_b200
from vkhr_b200 import HairStyle, capi
from harness import synth

# BRICK8 falls back to PACKED8 where it cannot run; BRICK8 | BRICK8_SPLIT = the separate kernels of round 1 instead of the frame kernel
STRATEGIES = [0, capi.STRATEGY_COUNT32, capi.STRATEGY_PACKED8, capi.STRATEGY_BRICK8, capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT]


def _fnv(port, a):
    return f"{port.fnv1a64(a):016x}"


def test_native_library_is_loaded():
    # the tests below would be meaningless on a fallback: assert the native .so is the one in use
    with open("/proc/self/maps") as f:
        assert "libvkhr_b200.so" in f.read()
    assert "sm_100a" in vkhr_b200.Voxelizer.version()


def test_fast_division_is_exact():
    """The kernels replace `a / d` by an FMA sequence with a precomputed reciprocal (walk.cuh div_exact);
    it must equal the IEEE division bit for bit, including operands on both sides of its range guard.
    (The self-test kernels live in the harness library, which includes the product's walk.cuh.)"""
    from harness import selftest
    rng = np.random.default_rng(3)
    divisors = [1.0, 0.21861044, 0.39062497, 1.9999999, 1.0000001, 3.0, 0.1, 7.0, 1.5, 0.75, 255.0, 1e-3, 1e4,
                float(np.float32(100.0) / np.float32(256.0)), float(np.float32(56.558083) / np.float32(256.0))]
    divisors += [float(x) for x in np.exp(rng.uniform(np.log(1e-4), np.log(1e4), 25)).astype(np.float32)]
    divisors += [float(np.frombuffer(np.uint32(0x3F7FFFFF + k).tobytes(), dtype=np.float32)[0]) for k in (0, 1, 2)]
    divisors += [1e-13, 1e13]            # outside the fast range: must still be exact (plain division)
    total = 0
    for i, d in enumerate(divisors):
        assert selftest.division(d, n_trials=1 << 26, seed=1000 + i) == 0, d
        total += 1 << 26
    assert total > 2_000_000_000


def test_walk_reciprocal_is_correctly_rounded_for_every_step_count():
    """`direction /= steps` (hair_style.cc:317): the walk's reciprocal (walk.cuh rcp_steps: MUFU.RCP + one Newton step,
    no range test) equals __frcp_rn, and the FMA division built on it equals the IEEE division, for EVERY float
    steps in [2^-40, 2^24] -- the whole interval the fast path admits (about 5.4e8 divisors, 3 numerators each)."""
    from harness import selftest
    assert selftest.rcp(9.094947e-13, 16777216.0) == 0


def test_unaligned_and_offset_device_buffers(vox, port):
    """Device pointers that are only 4-byte aligned (a view into a larger buffer)."""
    import torch
    dev = torch.device("cuda", 0)
    v, n, s = synth.shape("ponytail", seed=21, seg_len=1.0, scale=0.02)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    want = port.voxelize_segments(v, port.generate_indices(n, s), lo, size, 64, 64, 64)
    big = torch.zeros(v.size + 8, dtype=torch.float32, device=dev)
    for off in (0, 1, 2, 3):
        big[off:off + v.size] = torch.from_numpy(v.reshape(-1)).to(dev)
        view = big[off:off + v.size]
        got = vox.voxelize_segments_dev(view, None, lo, size, 64, 64, 64, segs_per_strand=s)
        assert np.array_equal(got.cpu().numpy(), want), off
        outbuf = torch.zeros(64 ** 3 + 16, dtype=torch.uint8, device=dev)
        got2 = vox.voxelize_segments_dev(view, None, lo, size, 64, 64, 64, segs_per_strand=s, out=outbuf[off + 1:off + 1 + 64 ** 3])
        assert np.array_equal(got2.cpu().numpy(), want), off


def test_kat4_through_hairstyle_mirror(vox, golden):
    """The reference's own call sequence (SceneGraph::add_style + voxelize_segments + normalize)."""
    k = golden["kat4"]
    hs = HairStyle(voxelizer=vox)
    hs.vertices = np.array(k["vertices"], dtype=np.float32)
    hs.set_strand_count(k["strands"])
    hs.set_default_segment_count(k["segments_per_strand"])
    hs.generate_tangents()
    hs.generate_indices()
    hs.generate_bounding_box()
    b = hs.get_bounding_box()
    assert np.array_equal(np.concatenate([b.origin, [b.radius], b.size, [b.volume]]).astype(np.float32),
                          np.array(k["aabb"], dtype=np.float32))
    assert hs.indices.tolist() == k["indices"]
    assert np.array_equal(hs.tangents, np.array(k["tangents_in"], dtype=np.float32), equal_nan=True)
    vol = hs.voxelize_segments(4, 4, 4)
    assert np.array_equal(vol.densities, dense(k["voxelize_segments"], 64))
    assert np.array_equal(hs.voxelize_vertices(4, 4, 4).densities, dense(k["voxelize_vertices"], 64))
    assert vol.downsample(capi.DOWNSAMPLE_SUM).densities.tolist() == k["downsample_sum_segments"]
    assert vol.downsample(capi.DOWNSAMPLE_MAX).densities.tolist() == k["downsample_max_segments"]
    vol.normalize()
    assert np.array_equal(vol.densities, dense(k["normalize_segments"], 64))


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("res", [(64, 64, 64), (64, 32, 16), (16, 16, 16), (30, 20, 10)])
def test_small_sets_golden(vox, small_sets, res, strategy):
    W, H, D = res
    if strategy == capi.STRATEGY_PACKED8 and (W * H * D) % 16:
        pytest.skip("PACKED8 needs W*H*D % 16 == 0")
    tag = f"{W}x{H}x{D}"
    v = small_sets["in_vertices"]
    n, s = [int(x) for x in small_sets["in_meta"]]
    bb = small_sets["aabb_generated"]
    lo, hi = vox.generate_bounding_box(v)
    assert np.array_equal(lo, bb[:3]) and np.array_equal(hi - lo, bb[4:7])
    hs = HairStyle(voxelizer=vox)
    hs.vertices = v
    hs.set_strand_count(n)
    hs.set_default_segment_count(s)
    hs.generate_indices()
    # explicit index buffer and implicit uniform strands must agree with the reference
    d = vox.voxelize_segments(v, hs.indices, bb[:3], bb[4:7], W, H, D, flags=strategy)
    assert np.array_equal(d, small_sets[f"seg_{tag}"])
    d2 = vox.voxelize_segments(v, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, flags=strategy)
    assert np.array_equal(d2, small_sets[f"seg_{tag}"])
    assert np.array_equal(vox.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D, flags=strategy), small_sets[f"ver_{tag}"])
    dn = vox.voxelize_segments(v, hs.indices, bb[:3], bb[4:7], W, H, D, flags=strategy | capi.NORMALIZE)
    assert np.array_equal(dn, small_sets[f"segnorm_{tag}"])
    assert np.array_equal(vox.normalize(d), small_sets[f"segnorm_{tag}"])
    if f"segdown0_{tag}" in small_sets:
        for f in range(4):
            assert np.array_equal(vox.downsample(d, W, H, D, f), small_sets[f"segdown{f}_{tag}"])


def test_small_header_aabb_and_variable_strands(vox, small_sets):
    v = small_sets["in_vertices"]
    n, s = [int(x) for x in small_sets["in_meta"]]
    bb = small_sets["aabb_header"]
    d = vox.voxelize_segments(v, None, bb[:3], bb[4:7], 64, 64, 64, segs_per_strand=s)
    assert np.array_equal(d, small_sets["seg_header_64x64x64"])
    assert np.array_equal(vox.voxelize_vertices(v, bb[:3], bb[4:7], 64, 64, 64), small_sets["ver_header_64x64x64"])
    hs = HairStyle(voxelizer=vox)
    hs.vertices = small_sets["var_vertices"]
    hs.segments = small_sets["var_segments"]
    hs.generate_indices()
    assert np.array_equal(hs.indices, small_sets["var_indices"])
    vbb = small_sets["var_aabb"]
    hs.set_bounding_box(vbb[:3], vbb[:3] + vbb[4:7])
    assert np.array_equal(hs.voxelize_segments(32, 32, 32).densities, small_sets["var_seg_32x32x32"])


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("name", ["ponytail_256", "ponytail_long_256", "ponytail_sat_32", "ponytail_noncubic", "straight_512"])
def test_full_size_fingerprints(vox, port, golden, name, strategy):
    """BASELINE.json configs 1/2 sizes: fingerprints of the unmodified reference's output."""
    e = golden["fingerprints"][name]
    v, n, s = synth.shape(e["shape"], seed=e["seed"], seg_len=e["seg_len"], scale=e["scale"])
    assert _fnv(port, v) == e["input_fnv"], "synthetic generator output changed: regenerate tests/golden"
    W, H, D = e["resolution"]
    bb = np.array(e["aabb"], dtype=np.float32)
    d = vox.voxelize_segments(v, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, flags=strategy)
    st = e["segments"]
    assert (_fnv(port, d), int(d.astype(np.int64).sum()), int(np.count_nonzero(d)), int((d == 255).sum())) == \
           (st["fnv"], st["sum"], st["nonzero"], st["saturated"])
    dv = vox.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D, flags=strategy)
    assert _fnv(port, dv) == e["vertices"]["fnv"]
    assert _fnv(port, vox.normalize(d)) == e["normalize_segments"]["fnv"]


@pytest.mark.parametrize("seed", range(8))
def test_random_sets_against_oracle(vox, port, seed):
    rng = np.random.default_rng(1000 + seed)
    n, s = int(rng.integers(50, 3000)), int(rng.integers(1, 24))
    v = synth.strands(n, s, seed=int(rng.integers(1, 2**31)), seg_len=float(rng.uniform(0.2, 6.0)),
                      curl=float(rng.uniform(0.1, 2.0)), gravity=float(rng.uniform(0, 0.6)))
    W, H, D = [int(x) for x in rng.choice([4, 8, 16, 33, 64, 100, 128], size=3)]
    lo, hi = port.generate_bounding_box(v)
    if seed % 2:
        lo, hi = v.min(axis=0) - np.float32(0.5), v.max(axis=0) + np.float32(0.25)
    size = (hi - lo).astype(np.float32)
    idx = port.generate_indices(n, s)
    for flags in (0, capi.INDEX_EXACT):
        want = port.voxelize_segments(v, idx, lo, size, W, H, D, flags=flags)
        for strat in STRATEGIES:
            if strat == capi.STRATEGY_PACKED8 and (W * H * D) % 16:
                continue
            got = vox.voxelize_segments(v, idx, lo, size, W, H, D, flags=flags | strat)
            assert np.array_equal(got, want), (seed, flags, strat)
        assert np.array_equal(vox.voxelize_vertices(v, lo, size, W, H, D, flags=flags),
                              port.voxelize_vertices(v, lo, size, W, H, D, flags=flags))


def test_edge_cases(vox, port):
    """Saturating clump, max-corner clamp, zero-length segments, out-of-box and NaN vertices, tiny inputs."""
    rng = np.random.default_rng(5)
    clump = np.tile(np.array([[1.25, 1.25, 1.25], [1.75, 2.6, 1.3]], dtype=np.float32), (400, 1))
    faces = np.array([[0, 0, 0], [8, 8, 8], [8, 8, 8], [8, 0, 8], [2, 2, 2], [2, 2, 2], [3, 3, 3], [3, 4, 3],
                      [7.999999, 7.999999, 7.999999], [8, 8, 8], [0, 8, 0], [8, 8, 0],
                      [9, 9, 9], [12, 9, 9], [-3, 1, 1], [2, 1, 1], [np.nan, 1, 1], [1, 1, 1]], dtype=np.float32)
    jitter = rng.uniform(0, 8, size=(200, 3)).astype(np.float32)
    v = np.concatenate([clump, faces, jitter], axis=0)
    idx = np.arange(v.shape[0], dtype=np.uint32)
    origin, size = np.zeros(3, np.float32), np.full(3, 8.0, np.float32)
    for res in [(8, 8, 8), (4, 4, 4), (5, 3, 2), (16, 4, 2)]:
        want = port.voxelize_segments(v, idx, origin, size, *res)
        assert want.max() == 255
        for strat in STRATEGIES:
            if strat == capi.STRATEGY_PACKED8 and np.prod(res) % 16:
                continue
            assert np.array_equal(vox.voxelize_segments(v, idx, origin, size, *res, flags=strat), want), (res, strat)
            assert np.array_equal(vox.voxelize_vertices(v, origin, size, *res, flags=strat),
                                  port.voxelize_vertices(v, origin, size, *res))
    # empty / ragged inputs
    z = np.zeros(64, dtype=np.uint8)
    assert np.array_equal(vox.voxelize_segments(v[:3], np.array([0], np.uint32), origin, size, 4, 4, 4), z)
    assert np.array_equal(vox.voxelize_segments(np.zeros((0, 3), np.float32), None, origin, size, 4, 4, 4,
                                                segs_per_strand=3), z)
    assert np.array_equal(vox.voxelize_vertices(np.zeros((0, 3), np.float32), origin, size, 4, 4, 4), z)
    odd = np.array([0, 1, 2], dtype=np.uint32)          # trailing lone index is ignored (reference loop bound)
    assert np.array_equal(vox.voxelize_segments(v, odd, origin, size, 4, 4, 4),
                          port.voxelize_segments(v, odd, origin, size, 4, 4, 4))
    flat = np.full(64, 7, dtype=np.uint8)
    assert np.array_equal(vox.normalize(flat), flat)     # max == min: unchanged


def test_error_behaviour(vox):
    v = np.zeros((4, 3), dtype=np.float32)
    o, s = np.zeros(3, np.float32), np.ones(3, np.float32)
    with pytest.raises(vkhr_b200.VkhrB200Error) as e:
        vox.voxelize_segments(v, None, o, s, 4, 4, 4, segs_per_strand=0)
    assert e.value.code == capi.ERR_INVALID_ARGUMENT
    with pytest.raises(vkhr_b200.VkhrB200Error):
        vox.voxelize_segments(v, None, o, s, 4, 4, 4, segs_per_strand=2)       # 4 % 3 != 0
    with pytest.raises(vkhr_b200.VkhrB200Error):
        vox.voxelize_segments(v, None, o, np.zeros(3, np.float32), 4, 4, 4, segs_per_strand=1)
    with pytest.raises(vkhr_b200.VkhrB200Error):
        vox.voxelize_vertices(v, o, s, 0, 4, 4)
    with pytest.raises(vkhr_b200.VkhrB200Error) as e:
        vox.voxelize_vertices(v, o, s, 2048, 2048, 2048)
    assert e.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(vkhr_b200.VkhrB200Error):
        vox.voxelize_vertices(v, o, s, 5, 3, 2, flags=capi.STRATEGY_PACKED8)


def test_device_api_batch_and_shards(vox, port):
    """Device-pointer API: the crowd batch, and fake ranks -- k partial u32 grids summed then clamped
    must be byte-identical to the single-rank volume."""
    import torch
    dev = torch.device("cuda", 0)
    W = H = D = 64
    insts, wants = [], []
    for k in range(5):
        v, n, s = synth.shape("ponytail", seed=77 + k, seg_len=0.6 + 0.3 * k, scale=0.01 + 0.002 * k)
        lo, hi = port.generate_bounding_box(v)
        size = (hi - lo).astype(np.float32)
        wants.append(port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D))
        insts.append({"vertices": torch.from_numpy(v).to(dev).reshape(-1), "segs_per_strand": s,
                      "aabb_origin": lo, "aabb_size": size,
                      "out": torch.full((W * H * D,), 9, dtype=torch.uint8, device=dev)})
    for strat in (0, capi.STRATEGY_COUNT32, capi.STRATEGY_BRICK8):
        for ins in insts:
            ins["out"].fill_(9)
        vox.voxelize_segments_batch_dev(insts, W, H, D, flags=strat)
        torch.cuda.synchronize()
        for ins, want in zip(insts, wants):
            assert np.array_equal(ins["out"].cpu().numpy(), want)
    # strand shards -> partial counts -> sum -> clamp
    v, n, s = synth.shape("ponytail", seed=5, seg_len=1.0, scale=0.03)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    W, H, D = 8, 8, 4                                       # coarse: many saturated voxels
    want = port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D)
    assert (want == 255).sum() > 0
    vt = torch.from_numpy(v).to(dev)
    for world in (1, 2, 4, 8):
        total = torch.zeros(W * H * D, dtype=torch.int32, device=dev)
        for rank in range(world):
            a, b = (n * rank) // world, (n * (rank + 1)) // world
            part = torch.zeros(W * H * D, dtype=torch.int32, device=dev)
            shard = vt[a * (s + 1): b * (s + 1)].contiguous().reshape(-1)
            if shard.numel():
                vox.count_segments_dev(shard, None, lo, size, W, H, D, part, segs_per_strand=s)
            total += part
        got = vox.clamp_counts_dev(total)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), world
    # the same through vertex counts
    part = torch.zeros(W * H * D, dtype=torch.int32, device=dev)
    vox.count_vertices_dev(vt.reshape(-1), lo, size, W, H, D, part)
    assert np.array_equal(part.cpu().numpy().astype(np.uint32), port.count_vertices(v, lo, size, W, H, D))


def test_host_crowd_pipelined(vox, port):
    """vkhr_b200_voxelize_segments_batch: more instances than staging slots, mixed uniform / indexed / empty members,
    pinned and pageable host buffers -- every volume must equal the oracle's."""
    import torch
    rng = np.random.default_rng(5)
    W, H, D = 64, 32, 16
    inst, want = [], []
    for k in range(8):
        v, n, s = synth.shape("ponytail", seed=100 + k, seg_len=0.8 + 0.1 * k, scale=0.01 + 0.002 * k)
        lo, hi = port.generate_bounding_box(v)
        size = (hi - lo).astype(np.float32)
        idx = port.generate_indices(n, s)
        if k % 2 == 0:                                       # pinned through torch
            vbuf = torch.from_numpy(v.reshape(-1).copy()).pin_memory().numpy()
            obuf = torch.empty(W * H * D, dtype=torch.uint8).pin_memory().numpy()
        else:                                                # pageable
            vbuf, obuf = v.reshape(-1).copy(), np.empty(W * H * D, dtype=np.uint8)
        obuf[:] = 0xAB
        d = {"vertices": vbuf, "out": obuf, "aabb_origin": lo, "aabb_size": size}
        if k % 3 == 1:
            perm = rng.permutation(idx.reshape(-1, 2)).reshape(-1).astype(np.uint32)
            d["indices"] = perm
            want.append(port.voxelize_segments(v, perm, lo, size, W, H, D))
        elif k == 5:
            d["indices"] = np.zeros(1, dtype=np.uint32)      # fewer than two indices: empty volume
            want.append(np.zeros(W * H * D, dtype=np.uint8))
        else:
            d["segs_per_strand"] = s
            want.append(port.voxelize_segments(v, idx, lo, size, W, H, D))
        inst.append(d)
    for rep in range(2):                                     # second call reuses slots and events
        vox.voxelize_segments_batch(inst, W, H, D)
        for k in range(8):
            assert np.array_equal(inst[k]["out"], want[k]), (rep, k)
    # host_register on a caller-owned numpy buffer
    reg = np.empty(W * H * D, dtype=np.uint8)
    vox.host_register(reg)
    inst[0]["out"] = reg
    vox.voxelize_segments_batch(inst[:1], W, H, D, flags=capi.NORMALIZE)
    vox.host_unregister(reg)
    assert np.array_equal(reg, port.normalize(want[0]))


def test_full_size_properties(vox, port):
    """Config-1 size (1.64 M segments, 256^3) and 512^3: properties that need no CPU run of the full job."""
    import torch
    dev = torch.device("cuda", 0)
    v, n, s = synth.shape("ponytail", seed=0xBEEF, seg_len=1.5)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    idx = port.generate_indices(n, s)
    vt = torch.from_numpy(v).to(dev).reshape(-1)
    for (W, H, D) in [(256, 256, 256), (512, 512, 512)]:
        nvox = W * H * D
        counts = torch.zeros(nvox, dtype=torch.int32, device=dev)
        vox.count_segments_dev(vt, None, lo, size, W, H, D, counts, segs_per_strand=s)
        # (1) unclamped hit counts == the oracle's (its walk alone takes about a second at these sizes),
        #     and no sample is lost: every sample of the reference walk is in the grid or dropped by rule
        want_counts = port.count_segments(v, idx, lo, size, W, H, D)
        assert np.array_equal(counts.cpu().numpy().astype(np.uint32), want_counts)
        assert int(counts.sum(dtype=torch.int64)) <= port.count_samples(v, idx, lo, size, W, H, D)
        del want_counts
        # (2) clamp of the counts == the direct u8 volume, for both strategies
        clamped = vox.clamp_counts_dev(counts)
        for strat in (capi.STRATEGY_PACKED8, capi.STRATEGY_COUNT32):
            d = vox.voxelize_segments_dev(vt, None, lo, size, W, H, D, segs_per_strand=s, flags=strat)
            assert torch.equal(d, clamped)
        # (3) additivity over a strand partition (what the multi-GPU sum relies on)
        half = (n // 2) * (s + 1) * 3
        c2 = torch.zeros(nvox, dtype=torch.int32, device=dev)
        vox.count_segments_dev(vt[:half].contiguous(), None, lo, size, W, H, D, c2, segs_per_strand=s)
        vox.count_segments_dev(vt[half:].contiguous(), None, lo, size, W, H, D, c2, segs_per_strand=s)
        assert torch.equal(c2, counts)
        # (4) explicit indices == implicit uniform strands
        it = torch.from_numpy(idx.astype(np.int32)).to(dev)
        d_idx = vox.voxelize_segments_dev(vt, it, lo, size, W, H, D)
        assert torch.equal(d_idx, clamped)
        # (5) normalize is idempotent once max == 255 and min == 0, and keeps zeros
        nrm = vox.normalize_dev(d_idx.clone())
        assert int(nrm.max()) in (254, 255) and int(nrm.min()) == 0
        assert torch.equal((nrm == 0), (d_idx == 0))
        # (6) downsample(max) of the volume dominates downsample(min); sum of MEAN*8 <= total
        if W == 256:
            mx = vox.downsample_dev(d_idx, W, H, D, capi.DOWNSAMPLE_MAX)
            mn = vox.downsample_dev(d_idx, W, H, D, capi.DOWNSAMPLE_MIN)
            assert bool((mx >= mn).all())
            # oracle on the full grid is cheap for downsample/normalize (no walk)
            dh = d_idx.cpu().numpy()
            assert np.array_equal(mx.cpu().numpy(), port.downsample(dh, W, H, D, 0))
            assert np.array_equal(nrm.cpu().numpy(), port.normalize(dh))
        torch.cuda.synchronize()


def _assert_tangents_close(got, want, dens, what):
    """Volume::tangents: the reference sums fp32 tangents in strand order, the GPU sums integers (order-free):
    within 1 LSB wherever the reference is defined and order-independent (0 < density < 255), w == 0 everywhere,
    and exactly 0 in empty voxels (0/0 -> NaN -> 0 on x86-64 in the reference)."""
    got, want = got.reshape(-1, 4).astype(np.int32), want.reshape(-1, 4).astype(np.int32)
    assert np.all(got[:, 3] == 0), what
    assert np.all(got[dens == 0] == 0), what
    m = (dens > 0) & (dens < 255)
    diff = np.abs(got[m, :3] - want[m, :3])
    assert diff.max(initial=0) <= 1, f"{what}: tangent off by {diff.max()} LSB"
    assert (diff == 0).mean() > 0.9, f"{what}: only {(diff == 0).mean():.3f} of the components identical"


@pytest.mark.parametrize("res", [(64, 64, 64), (64, 32, 16), (16, 16, 16), (30, 20, 10)])
def test_tangent_volume_golden(vox, port, small_sets, res):
    """Tangent volumes against the ones the unmodified reference produced (tests/golden/small_sets.npz)."""
    W, H, D = res
    tag = f"{W}x{H}x{D}"
    v = small_sets["in_vertices"]
    n, s = [int(x) for x in small_sets["in_meta"]]
    bb = small_sets["aabb_generated"]
    hs = HairStyle(voxelizer=vox)
    hs.vertices = v
    hs.set_strand_count(n)
    hs.set_default_segment_count(s)
    hs.generate_indices()
    hs.generate_tangents()
    hs.generate_bounding_box()
    want_d, want_t = small_sets[f"seg_{tag}"], small_sets[f"segtan_{tag}"]
    # explicit indices + explicit tangents (what HairStyle::voxelize_segments reads)
    d, t = vox.voxelize_segments(v, hs.indices, bb[:3], bb[4:7], W, H, D, tangents=hs.tangents)
    assert np.array_equal(d, want_d)
    _assert_tangents_close(t, want_t, want_d, f"indexed {tag}")
    # uniform strands, tangents derived on the fly from the segment direction
    d2, t2 = vox.voxelize_segments(v, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, want_tangents=True)
    assert np.array_equal(d2, want_d)
    _assert_tangents_close(t2, want_t, want_d, f"uniform {tag}")
    # device-resident call with a tangent output: the same bytes as the host-buffer call
    import torch
    dv = torch.from_numpy(v).cuda().reshape(-1)
    dt = torch.empty(4 * W * H * D, dtype=torch.int8, device="cuda")
    dd = vox.voxelize_segments_dev(dv, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, tangents_out=dt)
    torch.cuda.synchronize()
    assert np.array_equal(dd.cpu().numpy(), d2) and np.array_equal(dt.cpu().numpy().reshape(-1, 4), t2.reshape(-1, 4))
    # order independence: reversing the strand order must not change a single byte
    order = np.arange(n)[::-1]
    vr = v.reshape(n, s + 1, 3)[order].reshape(-1, 3)
    d3, t3 = vox.voxelize_segments(vr, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, want_tangents=True)
    assert np.array_equal(d3, d2) and np.array_equal(t3, t2)
    # vertices, against the CPU oracle
    wd, wt = port.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D, tangents=hs.tangents)
    gd, gt = vox.voxelize_vertices(v, bb[:3], bb[4:7], W, H, D, tangents=hs.tangents)
    assert np.array_equal(gd, wd)
    _assert_tangents_close(gt, wt, wd, f"vertices {tag}")
    # through the HairStyle mirror, normalised densities as the caller uploads them
    vol = hs.voxelize_segments(W, H, D, flags=capi.NORMALIZE)
    assert np.array_equal(vol.densities, small_sets[f"segnorm_{tag}"])
    _assert_tangents_close(vol.tangents, want_t, want_d, f"mirror {tag}")


def test_kat4_tangents(vox, golden):
    k = golden["kat4"]
    hs = HairStyle(voxelizer=vox)
    hs.vertices = np.array(k["vertices"], dtype=np.float32)
    hs.set_strand_count(k["strands"])
    hs.set_default_segment_count(k["segments_per_strand"])
    hs.generate_tangents()
    hs.generate_indices()
    hs.generate_bounding_box()
    vol = hs.voxelize_segments(4, 4, 4)
    for idx, want in k["voxelize_segments_tangents"].items():
        got = vol.tangents[int(idx)].astype(int)
        assert np.abs(got - np.array(want)).max() <= 1, (idx, got, want)
    assert np.all(vol.tangents[vol.densities == 0] == 0)


def test_tangent_volume_full_size(vox, port):
    """Ponytail-shaped set at 256^3: densities bit-exact with the density-only path, tangents within 1 LSB of the oracle."""
    v, n, s = synth.shape("ponytail", seed=0x5EED, seg_len=0.5)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    tin = port.generate_tangents(v, n, s)
    idx = port.generate_indices(n, s)
    wd, wt = port.voxelize_segments(v, idx, lo, size, 256, 256, 256, tangents=tin)
    d, t = vox.voxelize_segments(v, None, lo, size, 256, 256, 256, segs_per_strand=s, tangents=tin)
    assert np.array_equal(d, wd)
    assert np.array_equal(d, vox.voxelize_segments(v, None, lo, size, 256, 256, 256, segs_per_strand=s))
    _assert_tangents_close(t, wt, wd, "ponytail 256^3")


def test_cpp_dropin_against_linked_reference():
    """adapter/_build/dropin_test links the UNMODIFIED reference hair_style.cc and libvkhr_b200.so and runs the
    caller's sequence (voxelize_segments(256^3) + normalize) through both; built by __graft_entry__.build()."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adapter", "_build", "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("adapter/_build/dropin_test not built (needs the reference tree at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "drop-in test ok" in r.stdout
    assert r.stdout.count("bit-exact") == 5 and "MISMATCH" not in r.stdout


def test_cpp_sharded_entry_point_through_the_c_abi_only():
    """adapter/_build/sharded_test: a C++ caller splits one style over 2, 3 and 4 ranks with nothing but include/vkhr_b200.h
    (no CUDA headers, no NCCL, no torch); every rank's volume equals the one-context volume."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adapter", "_build", "sharded_test")
    if not os.path.exists(exe):
        pytest.skip("adapter/_build/sharded_test not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("byte-identical") == 3 and "MISMATCH" not in r.stdout


# ---- .hair file image -> volume (HairStyle::load + SceneGraph::add_style + voxelize_segments) -----------------------
def _hair_file(tmp_path, name, strands, segs, *, segments=None, indices=False, tangents=False, bbox=False, seed=4):
    from vkhr_b200.hair_style import HairStyle
    hs = HairStyle()
    if segments is None:
        hs.vertices = synth.strands(strands, segs, seed=seed)
        hs.set_strand_count(strands)
        hs.set_default_segment_count(segs)
    else:
        hs.vertices = np.concatenate([synth.strands(1, int(c), seed=seed + k) for k, c in enumerate(segments)])
        hs.segments = np.asarray(segments, dtype=np.uint16)
    if indices:
        hs.generate_indices()
    if tangents:
        hs.generate_tangents()
    if bbox:
        lo, hi = synth.host_bounding_box(hs.vertices)
        hs.set_bounding_box(lo - 0.25, hi + 0.5)               # NOT the generated box: the header's must be used
    p = str(tmp_path / name)
    assert hs.save(p)
    return p


@pytest.mark.parametrize("case", ["uniform", "uniform+bbox", "uniform+indices+tangents+bbox", "segments", "segments+bbox",
                                  "segments-equal", "segments+indices"])
@pytest.mark.parametrize("res", [(32, 32, 32), (40, 24, 16)])
def test_voxelize_hair_file_matches_the_reference_loader(vox, ref, tmp_path, case, res):
    """The reference loads the same file, generates what add_style would generate, and voxelises: same bytes."""
    W, H, D = res
    rng = np.random.default_rng(len(case))
    kw = {"indices": "indices" in case, "tangents": "tangents" in case, "bbox": "bbox" in case}
    if case.startswith("segments-equal"):
        p = _hair_file(tmp_path, "s.hair", 0, 0, segments=[5] * 37, **kw)
    elif case.startswith("segments"):
        p = _hair_file(tmp_path, "s.hair", 0, 0, segments=rng.integers(1, 9, 61), **kw)      # odd strand count: vertices at offset % 4 == 2
    else:
        p = _hair_file(tmp_path, "u.hair", 53, 7, **kw)
    r = ref.load(p).prepare()
    want, want_t, _ = r.voxelize("segments", W, H, D, want_tangents=True)
    blob = open(p, "rb").read()
    got, got_t, origin, size = vox.voxelize_hair(blob, W, H, D, want_tangents=True)
    assert np.array_equal(got, want), f"{case}: densities differ from the reference's"
    box = r.aabb
    assert np.array_equal(origin, box[0:3]) and np.array_equal(size, box[4:7])
    unsat = (want > 0) & (want < 255)                           # tangents: +-1 LSB where the reference's own sum is order-stable
    assert np.abs(got_t.astype(np.int16) - want_t.astype(np.int16))[unsat].max(initial=0) <= 1
    assert want.sum() > 0


def test_voxelize_hair_file_rejects_bad_images(vox, tmp_path):
    from vkhr_b200.capi import VkhrB200Error
    p = _hair_file(tmp_path, "ok.hair", 9, 3)
    blob = open(p, "rb").read()
    for bad in (b"", blob[:100], b"NOPE" + blob[4:], blob[:-8]):
        with pytest.raises(VkhrB200Error):
            vox.voxelize_hair(bad, 8, 8, 8)
    # a header whose counts do not add up
    import struct
    broken = bytearray(blob)
    struct.pack_into("<I", broken, 16, 4)                       # default_segment_count 3 -> 4
    with pytest.raises(VkhrB200Error):
        vox.voxelize_hair(bytes(broken), 8, 8, 8)


# ---- multi-GPU combine on one device: k "fake ranks"  ----------------------------------------------------
@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("res", [(64, 64, 64), (12, 8, 8)])
def test_fake_rank_u8_combine_equals_single_voxelisation(vox, port, k, res):
    """Strand shards voxelised into saturated u8 partials, combined by (a) the saturating-sum kernel over stacked slabs and
    (b) the fused peer-memory kernel (every fake rank sums its slab of all partials and stores it into all outputs):
    byte-identical to one voxelisation of the whole set -- also on a coarse grid where voxels pass 255."""
    import torch
    from vkhr_b200 import sharding
    W, H, D = res
    v, n, s = synth.shape("ponytail", seed=21, seg_len=1.0, scale=0.04)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    want = port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D)
    if res == (12, 8, 8):
        assert want.max() == 255
    nv = W * H * D
    nvp = (nv + 512 * k - 1) // (512 * k) * (512 * k)              # whole bitmap words per slab (the p2p schedule's padding)
    partials = torch.zeros((k, nvp), dtype=torch.uint8, device="cuda")
    for r in range(k):
        mine = np.ascontiguousarray(sharding.shard_vertices(v, n, s, k, r))
        if mine.size:
            vox.voxelize_segments_dev(torch.from_numpy(mine.reshape(-1)).cuda(), None, lo, size, W, H, D, segs_per_strand=s,
                                      out=partials[r, :nv])
    got = vox.saturating_sum_u8_dev(partials)
    torch.cuda.synchronize()
    assert np.array_equal(got[:nv].cpu().numpy(), want)
    outs = torch.full((k, nvp), 7, dtype=torch.uint8, device="cuda")
    slab = nvp // k
    pp, op = [partials[r].data_ptr() for r in range(k)], [outs[r].data_ptr() for r in range(k)]
    for r in range(k):
        vox.combine_peer_u8_dev(pp, op, r * slab, slab)
    torch.cuda.synchronize()
    for r in range(k):
        assert np.array_equal(outs[r, :nv].cpu().numpy(), want), f"fake rank {r}"
    # sparse form: chunk bitmaps, zeroed outputs, only non-zero chunks move
    bitmaps = torch.empty((k, nvp // 512), dtype=torch.int32, device="cuda")
    for r in range(k):
        vox.chunk_bitmap_dev(partials[r], bitmaps[r])
    chunks = partials.view(k, -1, 16).ne(0).any(dim=2)                                   # [k, n_chunks]
    bits = (bitmaps.view(k, -1, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1
    assert torch.equal(bits.view(k, -1).bool(), chunks), "chunk bitmap"
    outs.zero_()
    bp = [bitmaps[r].data_ptr() for r in range(k)]
    for r in range(k):
        vox.combine_peer_u8_sparse_dev(pp, bp, op, r * slab, slab)
    torch.cuda.synchronize()
    for r in range(k):
        assert np.array_equal(outs[r, :nv].cpu().numpy(), want), f"fake rank {r} (sparse)"


# ---- BRICK8: the brick-ordered scratch volume + copy-out ------------------------------------------------------------

def test_brick8_runs_where_it_can_and_falls_back_elsewhere(vox, port):
    v, n, s = synth.shape("ponytail", seed=31, seg_len=0.8, scale=0.02)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    idx = port.generate_indices(n, s)
    B = capi.STRATEGY_BRICK8
    cases = [((64, 64, 64), True, B), ((64, 32, 16), True, B), ((4, 4, 2), True, B), ((128, 8, 6), True, B), ((36, 20, 8), True, B),
             ((64, 64, 64), False, B), ((36, 20, 8), False, B),      # explicit index pairs (k_walk_indexed)
             ((30, 20, 9), True, capi.STRATEGY_COUNT32),            # W % 4 != 0 and W*H*D % 16 != 0
             ((30, 20, 10), True, capi.STRATEGY_PACKED8),           # W % 4 != 0
             ((32, 32, 15), True, capi.STRATEGY_PACKED8)]           # D odd
    for flag in (B, 0):                                             # BRICK8 is also what the default picks where it can run
        for (W, H, D), uniform, expect in cases:
            want = port.voxelize_segments(v, idx, lo, size, W, H, D)
            got = vox.voxelize_segments(v, None if uniform else idx, lo, size, W, H, D, segs_per_strand=s if uniform else 0, flags=flag)
            assert vox.last_strategy == expect, ((W, H, D), uniform, vox.last_strategy)
            assert np.array_equal(got, want), ((W, H, D), uniform)
    got = vox.voxelize_segments(v, None, lo, size, 64, 64, 64, segs_per_strand=s, flags=B | capi.INDEX_EXACT)
    assert vox.last_strategy == capi.STRATEGY_PACKED8                                   # exact-index mode keeps the linear layout
    assert np.array_equal(got, port.voxelize_segments(v, idx, lo, size, 64, 64, 64, flags=capi.INDEX_EXACT))


@pytest.mark.parametrize("res", [(8, 8, 4), (16, 8, 2), (4, 4, 2), (12, 12, 6)])
def test_brick8_saturation_is_repaired_exactly(vox, port, res):
    """More than 255 hits per voxel: the byte carries inside the brick word, the word is flagged at its LINEAR
    position and recounted after the copy-out."""
    W, H, D = res
    v, n, s = synth.shape("ponytail", seed=5, seg_len=1.0, scale=0.03)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    want = port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D)
    assert (want == 255).sum() > 0
    for _ in range(2):                                              # the second call sees the scratch the first one left
        got = vox.voxelize_segments(v, None, lo, size, W, H, D, segs_per_strand=s, flags=capi.STRATEGY_BRICK8)
        assert vox.last_strategy == capi.STRATEGY_BRICK8
        assert np.array_equal(got, want)
        assert np.array_equal(vox.voxelize_segments(v, None, lo, size, W, H, D, segs_per_strand=s, flags=capi.STRATEGY_BRICK8 | capi.NORMALIZE),
                              port.normalize(want))


@pytest.mark.parametrize("seed", range(6))
def test_brick8_strands_leaving_the_box_and_scratch_reuse(vox, port, seed):
    """A bounding box smaller than the hair (negative and beyond-the-grid coordinates: the reference's fp32 index
    still lands some of those samples in the grid), then a different style and resolution on the same context:
    the scratch must come back all zero every time."""
    rng = np.random.default_rng(4000 + seed)
    for trial in range(3):
        n, s = int(rng.integers(50, 2500)), int(rng.integers(1, 24))
        v = synth.strands(n, s, seed=int(rng.integers(1, 2**31)), seg_len=float(rng.uniform(0.2, 6.0)),
                          curl=float(rng.uniform(0.1, 2.0)), gravity=float(rng.uniform(0, 0.6)))
        W, H, D = int(rng.choice([4, 8, 16, 36, 64, 128])), int(rng.choice([4, 8, 20, 64, 100])), int(rng.choice([2, 4, 6, 32, 64]))
        lo, hi = v.min(axis=0), v.max(axis=0)
        shrink = np.float32(rng.uniform(0.0, 0.35))
        lo2 = (lo + shrink * (hi - lo)).astype(np.float32)
        size = ((hi - lo) * np.float32(1.0 - 1.7 * shrink)).astype(np.float32)
        want = port.voxelize_segments(v, port.generate_indices(n, s), lo2, size, W, H, D)
        got = vox.voxelize_segments(v, None, lo2, size, W, H, D, segs_per_strand=s, flags=capi.STRATEGY_BRICK8)
        assert vox.last_strategy == capi.STRATEGY_BRICK8
        assert np.array_equal(got, want), (seed, trial, (W, H, D))


@pytest.mark.parametrize("res,expect", [((512, 256, 256), capi.STRATEGY_BRICK8), ((1024, 128, 256), capi.STRATEGY_BRICK8),
                                        ((256, 512, 130), capi.STRATEGY_BRICK8), ((300, 300, 200), capi.STRATEGY_PACKED8)])
def test_brick8_above_2pow24_voxels_keeps_the_fp32_index_rounding(vox, port, res, expect):
    """Grids above 2^24 voxels: the reference's fp32 index rounds (hair_style.cc:321), and the ROUNDED index names the
    voxel.  With W and H powers of two the brick is taken from the bit fields of that index; other sizes stay PACKED8.
    A box smaller than the hair adds negative / clamped coordinates."""
    W, H, D = res
    v, n, s = synth.shape("ponytail", seed=41, seg_len=2.5, scale=0.03)
    lo, hi = port.generate_bounding_box(v)
    idx = port.generate_indices(n, s)
    for shrink in (0.0, 0.2):
        lo2 = (lo + np.float32(shrink) * (hi - lo)).astype(np.float32)
        size = ((hi - lo) * np.float32(1.0 - 1.6 * shrink)).astype(np.float32)
        want = port.voxelize_segments(v, idx, lo2, size, W, H, D)
        got = vox.voxelize_segments(v, None, lo2, size, W, H, D, segs_per_strand=s, flags=capi.STRATEGY_BRICK8)
        assert vox.last_strategy == expect
        assert np.array_equal(got, want), (res, shrink)
        got = vox.voxelize_segments(v, idx, lo2, size, W, H, D, flags=capi.STRATEGY_BRICK8)      # index pairs: k_walk_indexed
        assert vox.last_strategy == expect
        assert np.array_equal(got, want), (res, shrink, "indexed")
        # the rounding is real at these sizes: the exact-index volume differs
        if shrink == 0.0 and expect == capi.STRATEGY_BRICK8:
            assert not np.array_equal(want, port.voxelize_segments(v, idx, lo2, size, W, H, D, flags=capi.INDEX_EXACT))


def test_brick8_default_follows_the_segment_density(vox, port):
    """The default takes BRICK8 only when the copy-out pays: about one segment per 100 voxels or more."""
    v, n, s = synth.shape("ponytail", seed=43, seg_len=1.0, scale=0.01)            # 1363 strands x 12 segments
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    for res, expect in [((64, 64, 64), capi.STRATEGY_BRICK8), ((128, 128, 128), capi.STRATEGY_PACKED8)]:
        W, H, D = res
        got = vox.voxelize_segments(v, None, lo, size, W, H, D, segs_per_strand=s)
        assert vox.last_strategy == expect, res
        assert np.array_equal(got, port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D))


def test_brick8_batch_with_one_saturating_instance(vox, port):
    """BRICK8 decides per instance whether any byte carried (samples added == byte sum of the volume): in a batch only
    the instance that saturates is recounted, the others keep their copied-out volumes; a second run on the same
    context sees clean statistics and a clean scratch."""
    import torch
    dev = torch.device("cuda", 0)
    W, H, D = 16, 8, 4
    insts, wants = [], []
    for k, scale in enumerate((0.0005, 0.03, 0.001)):                     # sparse, dense (saturates), sparse
        v, n, s = synth.shape("ponytail", seed=60 + k, seg_len=1.0, scale=scale)
        lo, hi = port.generate_bounding_box(v)
        size = (hi - lo).astype(np.float32)
        wants.append(port.voxelize_segments(v, port.generate_indices(n, s), lo, size, W, H, D))
        insts.append({"vertices": torch.from_numpy(v).to(dev).reshape(-1), "segs_per_strand": s, "aabb_origin": lo,
                      "aabb_size": size, "out": torch.full((W * H * D,), 7, dtype=torch.uint8, device=dev)})
    assert (wants[1] == 255).sum() > 0 and (wants[0] == 255).sum() == 0 and (wants[2] == 255).sum() == 0
    for _ in range(2):
        for ins in insts:
            ins["out"].fill_(7)
        vox.voxelize_segments_batch_dev(insts, W, H, D, flags=capi.STRATEGY_BRICK8)
        torch.cuda.synchronize()
        assert vox.last_strategy == capi.STRATEGY_BRICK8
        for k, (ins, want) in enumerate(zip(insts, wants)):
            assert np.array_equal(ins["out"].cpu().numpy(), want), k


def test_brick8_crowd_at_256_equals_packed8(vox, port):
    """The bench configuration in small: ponytail-shaped instances at 256^3 through the batch entry point, every
    output byte written (pre-filled with 9), identical to PACKED8; instance 0 against the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    W = H = D = 256
    insts = []
    for k in range(3):
        v, n, s = synth.shape("ponytail", seed=900 + k, seg_len=0.5, scale=0.25 if k else 1.0)
        lo, hi = synth.host_bounding_box(v)
        insts.append({"vertices": torch.from_numpy(v).to(dev).reshape(-1), "segs_per_strand": s, "aabb_origin": lo,
                      "aabb_size": (hi - lo).astype(np.float32), "out": torch.full((W * H * D,), 9, dtype=torch.uint8, device=dev),
                      "host": (v, n, s, lo, hi)})
    outs = {}
    for strat in (capi.STRATEGY_PACKED8, capi.STRATEGY_BRICK8, capi.STRATEGY_BRICK8):
        for ins in insts:
            ins["out"].fill_(9)
        vox.voxelize_segments_batch_dev([{k: x for k, x in ins.items() if k != "host"} for ins in insts], W, H, D, flags=strat)
        torch.cuda.synchronize()
        assert vox.last_strategy == strat
        outs[strat] = [ins["out"].clone() for ins in insts]
    for a, b in zip(outs[capi.STRATEGY_PACKED8], outs[capi.STRATEGY_BRICK8]):
        assert torch.equal(a, b)
    v, n, s, lo, hi = insts[1]["host"]
    want = port.voxelize_segments(v, port.generate_indices(n, s), lo, (hi - lo).astype(np.float32), W, H, D)
    assert np.array_equal(outs[capi.STRATEGY_BRICK8][1].cpu().numpy(), want)


@pytest.mark.parametrize("ring_volumes", [1, 2, 3, 8])
def test_frame_kernel_rings_mixed_batches_and_repeated_frames(port, ring_volumes):
    """The frame kernel (one persistent launch: walk + copy-out through a ring of scratch volumes): batches larger than
    the ring, a ring of ONE volume (copy-out directly behind its walk), uniform / indexed / empty / saturating instances in
    one batch, and several frames in a row (the two control blocks alternate) -- always the oracle's volumes, and
    always identical to the split form."""
    import torch
    dev = torch.device("cuda", 0)
    W, H, D = 64, 32, 16
    nv = W * H * D
    rng = np.random.default_rng(77)
    with vkhr_b200.Voxelizer(0) as vox:
        vox.set_scratch_ring_bytes(ring_volumes * nv)
        insts, wants = [], []
        for k in range(7):
            scale = (0.001, 0.02, 0.004, 0.0, 0.06, 0.002, 0.01)[k]
            if scale == 0.0:                                              # an instance without segments: its volume must come back all zero
                v, n, s = np.zeros((0, 3), np.float32), 0, 5
                lo, size = np.zeros(3, np.float32), np.ones(3, np.float32)
                want = np.zeros(nv, np.uint8)
                idx = None
            else:
                if k == 4:                                                # a clump: thousands of hits in a few voxels (verdict + repair)
                    n, s = 3000, 12
                    v = synth.strands(n, s, seed=304, root_min=(1.0, 1.0, 1.0), root_max=(3.0, 3.0, 3.0), seg_len=0.4)
                    v = np.concatenate([v, synth.strands(50, s, seed=305, seg_len=1.5)])      # + a few strands that span the box
                    n += 50
                else:
                    v, n, s = synth.shape("ponytail", seed=300 + k, seg_len=float(rng.uniform(0.5, 2.0)), scale=scale)
                lo, hi = port.generate_bounding_box(v)
                size = (hi - lo).astype(np.float32)
                idx = port.generate_indices(n, s)
                want = port.voxelize_segments(v, idx, lo, size, W, H, D)
            ins = {"vertices": torch.from_numpy(v).to(dev).reshape(-1) if v.size else torch.zeros(3, dtype=torch.float32, device=dev)[:0],
                   "aabb_origin": lo, "aabb_size": size, "out": torch.full((nv,), 5, dtype=torch.uint8, device=dev)}
            if k % 3 == 2 and idx is not None:                            # explicit index pairs (what the reference's caller passes)
                ins["indices"] = torch.from_numpy(idx.astype(np.int32)).to(dev)
            else:
                ins["segs_per_strand"] = s
            insts.append(ins)
            wants.append(want)
        assert any((w == 255).any() for w in wants), "one instance must saturate (verdict + repair path)"
        for frame in range(3):
            for flags in (capi.STRATEGY_BRICK8, capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT):
                for ins in insts:
                    ins["out"].fill_(5)
                live = [i for i in insts if i["vertices"].numel()] if frame == 1 else insts    # frame 1: a different batch size
                live_w = [w for i, w in zip(insts, wants) if i["vertices"].numel()] if frame == 1 else wants
                vox.voxelize_segments_batch_dev(live, W, H, D, flags=flags)
                torch.cuda.synchronize()
                assert vox.last_strategy == capi.STRATEGY_BRICK8
                for k, (ins, want) in enumerate(zip(live, live_w)):
                    assert np.array_equal(ins["out"].cpu().numpy(), want), (frame, flags, k)


def test_frame_kernel_is_one_launch_per_frame(vox, port):
    """BRICK8 through the frame kernel: the frame kernel + the repair kernel's look at the flags; the split form: clear,
    walk, copy-out, verdict, repair."""
    import torch
    dev = torch.device("cuda", 0)
    v, n, s = synth.shape("ponytail", seed=9, seg_len=0.5, scale=0.05)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    vt = torch.from_numpy(v).to(dev).reshape(-1)
    out = torch.empty(64 ** 3, dtype=torch.uint8, device=dev)
    vox.voxelize_segments_dev(vt, None, lo, size, 64, 64, 64, segs_per_strand=s, out=out, flags=capi.STRATEGY_BRICK8)   # warm-up: allocations, memsets
    l0 = vox.launch_count
    vox.voxelize_segments_dev(vt, None, lo, size, 64, 64, 64, segs_per_strand=s, out=out, flags=capi.STRATEGY_BRICK8)
    assert vox.launch_count - l0 == 2
    l0 = vox.launch_count
    vox.voxelize_segments_dev(vt, None, lo, size, 64, 64, 64, segs_per_strand=s, out=out, flags=capi.STRATEGY_BRICK8 | capi.BRICK8_SPLIT)
    assert vox.launch_count - l0 == 5
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), port.voxelize_segments(v, port.generate_indices(n, s), lo, size, 64, 64, 64))


# ---- BASELINE.json configs at their FULL sizes ---------------------------------------------------------------------
from test_oracle import AT_SIZE, at_size_input  # noqa: E402


@pytest.mark.parametrize("name", AT_SIZE)
def test_baseline_configs_at_full_size(vox, port, golden, name):
    """configs[1] straight AND curly 50,000 x 65 at 512^3, configs[2] 1 M x 32 at 512^3, configs[4] swayed ponytail
    frames t = 0, 59, 119 in the union AABB of the 120-frame sequence at 1024^3: the default strategy and every forced
    one against fingerprints of the UNMODIFIED reference at that size (bit-exact; the fp32 index of hair_style.cc:321
    rounds in groups of up to 8 voxels at 512^3 and 64 at 1024^3)."""
    import torch
    dev = torch.device("cuda", 0)
    e = golden["fingerprints_at_size"][name]
    v, n, s = at_size_input(e)
    assert _fnv(port, v) == e["input_fnv"], "synthetic generator output changed: regenerate tests/golden (--sizes)"
    W, H, D = e["resolution"]
    bb = np.array(e["aabb"], dtype=np.float32)
    st = e["segments"]
    vt = torch.from_numpy(v).to(dev).reshape(-1)
    out = torch.empty(W * H * D, dtype=torch.uint8, device=dev)
    seen = set()
    for strat in STRATEGIES:
        out.fill_(3)
        vox.voxelize_segments_dev(vt, None, bb[:3], bb[4:7], W, H, D, segs_per_strand=s, flags=strat, out=out)
        seen.add(vox.last_strategy)
        d = out.cpu().numpy()
        assert (_fnv(port, d), int(d.astype(np.int64).sum()), int(np.count_nonzero(d)), int((d == 255).sum())) == \
               (st["fnv"], st["sum"], st["nonzero"], st["saturated"]), (name, strat)
    assert seen == {capi.STRATEGY_COUNT32, capi.STRATEGY_PACKED8, capi.STRATEGY_BRICK8}
    assert _fnv(port, vox.normalize_dev(out).cpu().numpy()) == e["normalize_segments"]["fnv"]
    # explicit index pairs (the reference's own caller passes them) through the default strategy
    if W == 512 and n <= 100_000:
        it = torch.from_numpy(port.generate_indices(n, s).astype(np.int32)).to(dev)
        d = vox.voxelize_segments_dev(vt, it, bb[:3], bb[4:7], W, H, D, out=out).cpu().numpy()
        assert _fnv(port, d) == st["fnv"]
    del out, vt
    torch.cuda.empty_cache()


def test_sway_sequence_union_bounding_box(vox, golden):
    """configs[4]: the fixed AABB of the animated sequence is the union of the per-frame generate_bounding_box results
    (the reference keeps the load-time AABB for every frame, rasterizer/hair_style.cc:66); all 120 frames on the GPU."""
    import torch
    dev = torch.device("cuda", 0)
    e = golden["fingerprints_at_size"]["ponytail_sway_t0_1024"]
    v0, n, s = synth.shape(e["shape"], seed=e["seed"], seg_len=e["seg_len"], scale=e["scale"])
    lo = hi = None
    box = torch.empty(6, dtype=torch.float32, device=dev)
    for t in range(120):
        vt = torch.from_numpy(synth.sway(v0, n, s, float(t))).to(dev).reshape(-1)
        b = vox.generate_bounding_box_dev(vt, out=box).cpu().numpy()
        lo = b[:3].copy() if lo is None else np.minimum(lo, b[:3])
        hi = b[3:].copy() if hi is None else np.maximum(hi, b[3:])
    assert np.array_equal(lo, np.array(e["union_aabb_min"], np.float32))
    assert np.array_equal(hi, np.array(e["union_aabb_max"], np.float32))


def test_bounding_box_signed_zeros_follow_the_reference_fold(vox, port, ref):
    """generate_bounding_box folds with glm::min(position, min) = (min < position) ? min : position (hair_style.cc:215-234):
    on a tie the NEW position wins and -0.0f ties with +0.0f, so the sign of a zero minimum / maximum is that of the last
    zero coordinate in vertex order.  Byte-identical to the port and to the unmodified reference."""
    rng = np.random.default_rng(11)
    cases = [np.array([[-0.0, 1.0, 2.0], [3.0, -0.0, 0.0], [1.0, 2.0, -0.0]], dtype=np.float32),
             np.array([[0.0, -0.0, 5.0], [-0.0, 0.0, 1.0], [2.0, 3.0, 4.0]], dtype=np.float32),
             np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=np.float32)]
    big = rng.uniform(0.0, 5.0, size=(5000, 3)).astype(np.float32)
    big[rng.integers(0, 5000, 40), rng.integers(0, 3, 40)] = -0.0
    big[rng.integers(0, 5000, 40), rng.integers(0, 3, 40)] = 0.0
    cases.append(big)
    for v in cases:
        lo, hi = vox.generate_bounding_box(v)
        plo, phi = port.generate_bounding_box(v)
        assert lo.tobytes() == plo.tobytes() and hi.tobytes() == phi.tobytes()
        n = v.shape[0] - (v.shape[0] % 2)
        h = ref.create(v[:n], n // 2, 1)
        if n == v.shape[0]:
            assert h.aabb[:3].tobytes() == lo.tobytes()
    # NaN coordinates: the reference's fold is order-dependent there; defined here as ignored
    v = np.array([[1.0, np.nan, 2.0], [np.nan, 3.0, -1.0]], dtype=np.float32)
    lo, hi = vox.generate_bounding_box(v)
    assert lo.tolist() == [0.0, 0.0, -1.0] and hi.tolist() == [1.0, 3.0, 2.0]


@pytest.mark.parametrize("k", [2, 4])
def test_sharded_entry_point_with_fake_ranks_on_one_device(port, k):
    """vkhr_b200_voxelize_segments_sharded_dev, the multi-GPU form of voxelize_segments in the C ABI: k "ranks" = k contexts
    with a stream each on ONE device, plain device pointers as peer buffers.  Partial volumes, chunk bitmaps, the
    device-side barriers over the signal pads and the fused sparse combine all run as on k GPUs; every rank must end up
    with the volume of the WHOLE strand set, frame after frame (the barriers pair up by call count)."""
    import torch
    from vkhr_b200 import sharding
    dev = torch.device("cuda", 0)
    W, H, D = 64, 48, 32
    v, n, s = synth.shape("ponytail", seed=71, seg_len=1.0, scale=0.02)
    lo, hi = port.generate_bounding_box(v)
    size = (hi - lo).astype(np.float32)
    nv = W * H * D
    ranks = [vkhr_b200.Voxelizer(0) for _ in range(k)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    nvp = ranks[0].sharded_volume_bytes(W, H, D, k)
    assert nvp % (512 * k) == 0 and nvp >= nv
    partials = [torch.zeros(nvp, dtype=torch.uint8, device=dev) for _ in range(k)]
    bitmaps = [torch.zeros(nvp // 512, dtype=torch.int32, device=dev) for _ in range(k)]
    outs = [torch.full((nvp,), 9, dtype=torch.uint8, device=dev) for _ in range(k)]
    signals = [torch.zeros(32, dtype=torch.int32, device=dev) for _ in range(k)]
    ptrs = lambda ts: [t.data_ptr() for t in ts]   # noqa: E731
    # ONE DEVICE ONLY: a context's first voxelisation allocates its scratch, and a device memory allocation serialises the
    # device's streams (CUDA's implicit synchronisation) -- inside a sharded call that would put rank r + 1's kernels behind
    # rank r's waiting barrier kernel.  So every context allocates up front (ranks on different GPUs need none of this).
    warm = torch.from_numpy(v).to(dev).reshape(-1)
    for r in range(k):
        with torch.cuda.stream(streams[r]):
            ranks[r].voxelize_segments_dev(warm, None, lo, size, W, H, D, segs_per_strand=s, out=outs[r][:nv])
    torch.cuda.synchronize()
    try:
        for frame in range(3):
            vf = synth.sway(v, n, s, float(7 * frame))                     # the hair moves: a different volume every frame
            want = port.voxelize_segments(vf, port.generate_indices(n, s), lo, size, W, H, D)
            shards = [torch.from_numpy(np.ascontiguousarray(sharding.shard_vertices(vf, n, s, k, r))).to(dev).reshape(-1) for r in range(k)]
            torch.cuda.synchronize()
            for r in range(k):
                with torch.cuda.stream(streams[r]):
                    ranks[r].voxelize_segments_sharded_dev(shards[r], None, lo, size, W, H, D, r, ptrs(partials), ptrs(bitmaps), ptrs(outs),
                                                           ptrs(signals), segs_per_strand=s)
            torch.cuda.synchronize()
            for r in range(k):
                assert np.array_equal(outs[r][:nv].cpu().numpy(), want), (frame, r)
    finally:
        for vox in ranks:
            vox.close()
