"""Host logic and the C-ABI boundary, without a GPU: symbol export, struct layout, .hair I/O,
the generators of the HairStyle mirror, and loud failure when no device is present."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import vkhr_b200
from vkhr_b200 import HairStyle, capi
from harness import synth


def test_library_exports_every_declared_symbol():
    names = capi.header_symbols()
    assert len(names) >= 28
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/vkhr_b200.h but not exported: {missing}"
    for n in names:
        assert n in capi._PROTOTYPES, f"{n} has no ctypes prototype"
        getattr(capi.lib, n)
    # nothing but the C ABI leaks out of the shared object
    leaked = [e for e in exported if not e.startswith("vkhr_b200_") and not e.startswith("_")]
    assert not leaked, leaked


def test_library_is_sm100a_and_links_no_oracle():
    sass = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if sass.returncode == 0:
        assert "sm_100a" in sass.stdout
    deps = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "oracle" not in deps and "vkhr_ref" not in deps


def test_no_product_file_touches_the_oracle():
    root = os.path.dirname(os.path.abspath(vkhr_b200.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f


def test_instance_struct_layout_matches_header():
    # typedef struct vkhr_b200_instance { ptr, ptr, u64, u32, u32, float[3], float[3], ptr }
    assert C.sizeof(capi.Instance) == 64
    assert capi.Instance.d_densities_out.offset == 56
    assert capi.Instance.aabb_origin.offset == 32


def test_create_fails_loudly_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(vkhr_b200.VkhrB200Error) as e:
        vkhr_b200.Voxelizer(0)
    assert e.value.code == capi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def _style(n=40, s=5, seed=3):
    hs = HairStyle()
    hs.vertices = synth.strands(n, s, seed=seed, seg_len=1.1)
    hs.set_strand_count(n)
    hs.set_default_segment_count(s)
    return hs


def test_generators_match_oracle(port):
    hs = _style()
    hs.generate_indices()
    hs.generate_tangents()
    hs.generate_thickness(0.042)
    assert np.array_equal(hs.indices, port.generate_indices(40, 5))
    assert np.array_equal(hs.tangents, port.generate_tangents(hs.vertices, 40, 5))
    assert hs.get_segment_count() == 200 and hs.get_vertex_count() == 240
    assert hs.thickness[5] == 0 and hs.thickness[4] == np.float32(0.042)
    # variable strand lengths (has_segments route)
    segs = np.array([1, 4, 2, 7, 3], dtype=np.uint16)
    v = np.concatenate([synth.strands(1, int(c), seed=9 + k) for k, c in enumerate(segs)])
    hs2 = HairStyle()
    hs2.vertices, hs2.segments = v, segs
    hs2.generate_indices()
    hs2.generate_tangents()
    assert hs2.get_strand_count() == 5
    assert np.array_equal(hs2.indices, port.generate_indices(5, 0, segments=segs))
    assert np.array_equal(hs2.tangents, port.generate_tangents(v, 5, 0, segments=segs))
    lo, hi = synth.host_bounding_box(v)
    plo, phi = port.generate_bounding_box(v)
    assert np.array_equal(lo, plo) and np.array_equal(hi, phi)
    hs2.set_bounding_box(lo, hi)
    b = hs2.get_bounding_box()
    assert np.array_equal(np.concatenate([b.origin, [b.radius], b.size, [b.volume]]).astype(np.float32),
                          port.get_bounding_box(lo, hi))


def test_hair_file_roundtrip_with_reference_loader(tmp_path, ref):
    """A .hair written by the mirror loads in the unmodified reference and vice versa (Appendix C)."""
    hs = _style(30, 4)
    hs.generate_indices()
    hs.generate_tangents()
    lo, hi = synth.host_bounding_box(hs.vertices)
    hs.set_bounding_box(lo, hi)
    p = str(tmp_path / "a.hair")
    assert hs.save(p)
    assert os.path.getsize(p) == 128 + 150 * 12 * 2 + 120 * 2 * 4
    r = ref.load(p)
    assert r.vertex_count == 150 and r.strand_count == 30 and r.segment_count == 120 and r.has_bounding_box
    assert np.array_equal(r.vertices, hs.vertices) and np.array_equal(r.indices, hs.indices)
    assert np.array_equal(r.tangents, hs.tangents)
    b = hs.get_bounding_box()
    assert np.array_equal(r.aabb, np.concatenate([b.origin, [b.radius], b.size, [b.volume]]).astype(np.float32))
    q = str(tmp_path / "b.hair")
    r.save(q)
    back = HairStyle(q)
    assert np.array_equal(back.vertices, hs.vertices) and np.array_equal(back.indices, hs.indices)
    assert back.has_bounding_box() and back.get_default_segment_count() == 4
    # variable segments survive too
    segs = np.array([2, 3, 1], dtype=np.uint16)
    hv = HairStyle()
    hv.vertices = np.concatenate([synth.strands(1, int(c), seed=k + 1) for k, c in enumerate(segs)])
    hv.segments = segs
    hv.generate_indices()
    assert hv.save(p)
    r2 = ref.load(p)
    assert r2.segments.tolist() == [2, 3, 1] and np.array_equal(r2.indices, hv.indices)
    assert not HairStyle().load(str(tmp_path / "missing.hair"))
    with open(p, "wb") as f:
        f.write(b"NOPE" + bytes(124))
    assert not HairStyle().load(p)


def test_synth_shapes():
    v, n, s = synth.shape("ponytail", scale=0.01)
    assert v.shape == (n * (s + 1), 3) and s == 12 and v.dtype == np.float32
    v2, _, _ = synth.shape("ponytail", scale=0.01)
    assert np.array_equal(v, v2)                               # deterministic
    d = np.linalg.norm(np.diff(v.reshape(n, s + 1, 3), axis=1), axis=2)
    assert np.allclose(d, 0.5, atol=1e-4)                      # fixed segment length
    w = synth.sway(v, n, s, t=3.0)
    assert np.array_equal(w.reshape(n, s + 1, 3)[:, 0], v.reshape(n, s + 1, 3)[:, 0])   # roots stay
    assert not np.array_equal(w, v)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` needs no GPU: rank 0 prints ONE JSON line (metric, unit, cpu_baseline, e2e with zero
    copy bytes), every other rank prints nothing and exits 0."""
    import json
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--res", "64"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "M seg/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the reference arm must not map the product library: it imports oracle/ and harness/ only
    probe = subprocess.run([sys.executable, "-c",
                            "import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--res', '32'];"
                            "sys.path.insert(0, %r); import runpy; runpy.run_path(os.path.join(%r, 'bench.py'), run_name='__main__');"
                            "m = open('/proc/self/maps').read(); assert 'libvkhr_b200' not in m, 'product library mapped'; "
                            "assert 'vkhr_b200' not in sys.modules" % (root, root)],
                           capture_output=True, text=True, env=env, timeout=600)
    assert probe.returncode == 0, probe.stderr[-600:]
    env["RANK"] = "1"
    other = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, env=env, timeout=600)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_ao_column_form_equals_per_voxel_form_in_host_emulation(tmp_path):
    """lao_column (register-tiled z column) against lao_at (one evaluation per voxel), both taken verbatim from
    prefilter.cuh and compiled for the host: every output bit-identical, for every tap-offset pair the kernel is
    instantiated for (tests/host_emulation/lao_column_check.cc).  The GPU test of the same name checks the kernels."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "vkhr_b200", "csrc", "prefilter.cuh")).read()
    body = text[text.index("constexpr int kPfTX"):text.index("// Gaussian weight of tap")]
    assert "lao_column" in body and "lao_column_rolled" in body and "lao_at" in body
    extract = tmp_path / "lao_extract.inc"
    extract.write_text(body.replace("__device__ __forceinline__", "static inline"))
    exe = tmp_path / "lao_column_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", f'-DLAO_EXTRACT="{extract}"',
                           "-o", str(exe), os.path.join(root, "tests", "host_emulation", "lao_column_check.cc")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "TOTAL bad=0" in out.stdout, out.stdout[-2000:]
    assert "nontrivial=0\n" not in out.stdout.split("fill=0.00")[0], "the dense cases must exercise non-empty outputs"


def test_brick8_layout_is_a_permutation_of_the_volume_words():
    """The BRICK8 scratch layout restated in numpy (walk.cuh brick_word / brick_word_pow2, kernels.cuh k_untile_batch):
    every 32-bit word of the x-fastest volume has exactly one brick word, the copy-out's (brick, word) -> linear word
    mapping is its inverse, and the bit-field form used above 2^24 voxels equals the general form on power-of-two grids."""
    for W, H, D in [(4, 4, 2), (8, 4, 2), (36, 20, 8), (64, 32, 16), (128, 8, 6), (256, 256, 4)]:
        z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
        x, y, z = x.ravel(), y.ravel(), z.ravel()
        lin = (z * H + y) * W + x
        brick = ((z >> 1) * (H >> 2) + (y >> 2)) * (W >> 2) + (x >> 2)
        bword = (brick << 3) | ((z & 1) << 2) | (y & 3)
        bbyte = (bword << 2) | (x & 3)                                   # byte address in the scratch volume
        assert np.array_equal(np.sort(bbyte), np.arange(W * H * D)), (W, H, D)
        assert np.array_equal(bbyte & 3, lin & 3)                        # the same byte of the word in both layouts
        # the copy-out: brick b, word w of the brick -> linear word
        wrow, byn = W >> 2, H >> 2
        b, w = bword >> 3, bword & 7
        bx, t = b % wrow, b // wrow
        by, bz = t % byn, t // byn
        lin_word = (2 * bz + (w >> 2)) * (wrow * H) + (4 * by + (w & 3)) * wrow + bx
        assert np.array_equal(lin_word, lin >> 2), (W, H, D)
        if W & (W - 1) == 0 and H & (H - 1) == 0:
            lw, lh = W.bit_length() - 1, H.bit_length() - 1
            x2, y2, z2 = lin & (W - 1), (lin >> lw) & (H - 1), lin >> (lw + lh)
            b2 = ((((z2 >> 1) << (lh - 2)) | (y2 >> 2)) << (lw - 2)) | (x2 >> 2)
            assert np.array_equal((b2 << 3) | ((z2 & 1) << 2) | (y2 & 3), bword), (W, H, D)


def test_byte_sum_equals_sample_count_iff_no_byte_carried():
    """The overflow test of the BRICK8 walk: packed-u8 adds of 1 << 8*byte into 32-bit words; the byte sum of the
    volume equals the number of adds exactly when no voxel received more than 255 of them."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        n_words = int(rng.integers(1, 6))
        hot = rng.random() < 0.5                                          # half of the trials push some voxel past 255
        n_adds = int(rng.integers(200, 1500)) if hot else int(rng.integers(0, 200))
        vox = rng.integers(0, 4 * n_words, n_adds) if not hot else np.where(rng.random(n_adds) < 0.7, rng.integers(0, 4 * n_words), rng.integers(0, 4 * n_words, n_adds))
        words = np.zeros(n_words, dtype=np.uint64)
        for v in vox:
            words[v >> 2] = (words[v >> 2] + (np.uint64(1) << np.uint64(8 * (v & 3)))) & np.uint64(0xFFFFFFFF)
        byte_sum = int(sum(int((w >> np.uint64(8 * b)) & np.uint64(0xFF)) for w in words for b in range(4)))
        counts = np.bincount(vox, minlength=4 * n_words)
        assert (byte_sum == n_adds) == bool((counts <= 255).all()), (trial, byte_sum, n_adds, counts.max())
        assert byte_sum <= n_adds


def test_volume_save_writes_what_the_reference_writes(tmp_path, ref):
    """Volume::save (hair_style.cc:359-369) through the C ABI and the Python mirror: the same bytes on disk as the
    unmodified reference's Volume::save, and False (not an exception) for a path that cannot be written."""
    from vkhr_b200.hair_style import AABB, Volume
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, 24 * 16 * 8, dtype=np.uint8)
    ours, theirs = str(tmp_path / "ours.raw"), str(tmp_path / "ref.raw")
    vol = Volume(resolution=np.array([24.0, 16.0, 8.0], np.float32), bounds=AABB(np.zeros(3, np.float32), 1.0, np.ones(3, np.float32), 1.0),
                 densities=d)
    assert vol.save(ours) is True
    assert ref.volume_save(d, theirs) is True
    assert open(ours, "rb").read() == open(theirs, "rb").read() == d.tobytes()
    assert vol.save(str(tmp_path / "no_such_dir" / "x.raw")) is False
    assert ref.volume_save(d, str(tmp_path / "no_such_dir" / "x.raw")) is False
    assert capi.lib.vkhr_b200_volume_save(None, None, 0) == capi.ERR_INVALID_ARGUMENT


def test_copy_out_brick_coordinates_by_multiply_high():
    """numpy restatement of the frame kernel's copy-out addressing (kernels.cuh, k_frame): brick number b ->
    (bx, by, bz) with m = floor((2^32 - 1) / d) + 1 and one correction step, for every divisor a volume of the frame
    kernel can have (W / 4, H / 4 < 65536) and brick numbers up to 2^23 -- against integer division."""
    rng = np.random.default_rng(7)
    ds = np.unique(np.concatenate([np.arange(1, 300), 2 ** np.arange(0, 16), 2 ** np.arange(1, 16) - 1, 2 ** np.arange(1, 16) + 1,
                                   rng.integers(1, 65536, 500), [65535]])).astype(np.uint64)
    ds = ds[ds < 65536]
    b = np.concatenate([np.arange(0, 4096), rng.integers(0, 1 << 23, 20000), [(1 << 23) - 1, (1 << 20) - 1, 1 << 20]]).astype(np.uint64)
    for d in ds:
        m = np.uint64(0) if d == 1 else np.uint64(0xFFFFFFFF) // d + np.uint64(1)
        assert m < (1 << 32)
        q = b.copy() if d == 1 else (b * m) >> np.uint64(32)                   # __umulhi
        r = (b.astype(np.int64) - (q * d).astype(np.int64))                    # 32-bit wrap == signed difference here
        neg = r < 0
        q = np.where(neg, q - np.uint64(1), q)
        r = np.where(neg, r + np.int64(d), r)
        assert np.array_equal(q, b // d) and np.array_equal(r.astype(np.uint64), b % d), int(d)


def _simulate_frame_schedule(items, slots, ring, copiers, copiers_last, rng):
    """The dependency structure of k_frame (kernels.cuh) as a model: CTAs are dispatched in linear block order
    (instance-major) into `slots` resident places and leave when their last phase is done.  Phases of CTA (i, x):
    [i >= ring: wait until instance i - ring is copied out] -> walk, report -> [one of the last min(items_i, copiers)
    CTAs of i > 0: wait until every CTA of instance i - 1 has reported, copy, count] -> [last instance, one of its
    last min(items, copiers_last) CTAs: wait for every CTA of its own instance, copy, count].  Phases complete in
    random order (any interleaving the hardware could produce).  Returns True when every CTA finishes."""
    n = len(items)
    order = [(i, x) for i in range(n) for x in range(items[i])]
    walk_done, copy_done = [0] * n, [0] * n
    need_copies = [min(items[t + 1], copiers) if t + 1 < n else min(items[t], copiers_last) for t in range(n)]
    resident, nxt, finished = {}, 0, 0            # (i, x) -> phase index
    while finished < len(order):
        while nxt < len(order) and len(resident) < slots:
            resident[order[nxt]] = 0
            nxt += 1
        movable = []
        for (i, x), ph in resident.items():
            if ph == 0:                            # slot wait
                ok = i < ring or copy_done[i - ring] >= need_copies[i - ring]
            elif ph == 1:                          # walk + report: never blocks
                ok = True
            elif ph == 2:                          # role 0
                mine = i > 0 and x + min(items[i], copiers) >= items[i]
                ok = (not mine) or walk_done[i - 1] >= items[i - 1]
            else:                                  # role 1
                mine = i == n - 1 and x + min(items[i], copiers_last) >= items[i]
                ok = (not mine) or walk_done[i] >= items[i]
            if ok:
                movable.append((i, x))
        if not movable:
            return False                           # every resident CTA waits and nothing else can be dispatched
        i, x = movable[rng.integers(len(movable))]
        ph = resident[(i, x)]
        if ph == 1:
            walk_done[i] += 1
        elif ph == 2 and i > 0 and x + min(items[i], copiers) >= items[i]:
            copy_done[i - 1] += 1
        elif ph == 3 and i == n - 1 and x + min(items[i], copiers_last) >= items[i]:
            copy_done[i] += 1
        if ph == 3:
            del resident[(i, x)]
            finished += 1
        else:
            resident[(i, x)] = ph + 1
    return all(copy_done[t] == need_copies[t] for t in range(n)) and all(walk_done[t] == items[t] for t in range(n))


def test_frame_kernel_schedule_cannot_deadlock_in_the_model():
    """Whatever the batch (ragged, instances of one item = no segments), the ring (>= 2 slots for a batch), the copier
    counts and the interleaving: with in-order dispatch every wait of k_frame is on a smaller block index, except the
    last instance's own copiers, which the host clamps to half the device's CTA slots (run_frame)."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        n = int(rng.integers(1, 9))
        items = [int(rng.integers(1, 30)) for _ in range(n)]
        slots = int(rng.integers(2, 40))
        ring = int(rng.integers(2, 5)) if n > 1 else 1
        copiers = int(rng.integers(1, 40))
        copiers_last = max(1, min(int(rng.integers(1, 40)), slots // 2))          # the host's clamp
        assert _simulate_frame_schedule(items, slots, min(ring, n), copiers, copiers_last, rng), (items, slots, ring, copiers, copiers_last)
    # the clamp is what makes it so: more waiting copiers of the last instance than slots cannot finish
    assert not _simulate_frame_schedule([20], 8, 1, 4, 16, np.random.default_rng(0))
