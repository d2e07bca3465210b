// kernels.cuh -- sm_100a kernels of the strand-voxelisation path.
//
// Reference behaviour being replaced (paths relative to the reference tree):
//   segment walk + saturating u8 count   src/vkhr/scene_graph/hair_style.cc:311-329
//   vertex splat                          src/vkhr/scene_graph/hair_style.cc:272-281
//   Volume::normalize                     src/vkhr/scene_graph/hair_style.cc:344-357
//   Volume::downsample                    include/vkhr/scene_graph/hair_style.hh:228-257
//   generate_bounding_box                 src/vkhr/scene_graph/hair_style.cc:215-234
#pragma once
#include "walk.cuh"

namespace vkhr_b200 {

// Measurement builds only (tools/runs/gpu_r2_ag.sh; never the product): the frame kernel without its reds / without its copy-out.
#ifndef VKHR_PROBE_NO_RED
#define VKHR_PROBE_NO_RED 0
#endif
#ifndef VKHR_PROBE_NO_COPY
#define VKHR_PROBE_NO_COPY 0
#endif

constexpr int kWalkThreads = 256;
#ifndef VKHR_WALK_MIN_CTAS
#define VKHR_WALK_MIN_CTAS 5
#endif

// How an instance's work is tiled over CTAs.
enum WalkKind : uint32_t { WK_UNIFORM = 0, WK_INDEXED = 1, WK_SPLAT = 2 };

// Device-side description of one instance of a batch.
struct InstanceDev {
    const float*    vertices;
    const uint32_t* indices;        // nullptr => uniform strands
    uint64_t        n_segments;
    uint32_t        n_vertices;
    uint32_t        segs_per_strand;
    GridParams      grid;
    uint8_t*        densities;      // W*H*D u8 (PACKED8: counted in place)
    uint32_t*       counts;         // W*H*D u32 (COUNT32 / recount scratch) or nullptr
    uint32_t*       ovf_bitmap;     // PACKED8: 1 bit per 32-bit word of `densities`
    uint32_t*       ovf_flag;       // PACKED8: != 0 when any word overflowed
    uint8_t*        brick;          // BRICK8: W*H*D u8 scratch in brick order (zero between calls), or nullptr
    unsigned long long* stats;      // BRICK8: kStatSlots partial counts of the samples the walk added, then kStatSlots
                                    // partial byte sums of the copy-out (k_brick_verdict compares the totals); or nullptr
    uint32_t        n_tiles;        // CTAs this instance needs in its walk kernel
    uint32_t        kind;           // WalkKind
    uint32_t        vps_magic;      // floor(2^32 / (segs_per_strand + 1)) + 1   (strand-end test without a division)
    uint32_t        tile_step;      // 31 mod (segs_per_strand + 1): how far the strand position moves from one warp-tile to the next
};

// A batch travels to the kernels as a __grid_constant__ parameter: every per-instance constant is then
// read through the constant cache, not through L1TEX (which the atomics need), and no table upload
// precedes the launch.  blockIdx.y + first selects the instance.
constexpr uint32_t kMaxBatch = 64;
// BRICK8 statistics are spread over slots: thousands of warps adding to ONE address serialise at L2 (measured: the
// crowd walk went from 0.97 to 1.37 ms with a single counter per instance).
constexpr uint32_t kStatSlots = 256;
struct Batch {
    uint32_t n;
    uint32_t pad[3];
    InstanceDev inst[kMaxBatch];
};
static_assert(sizeof(Batch) <= 16 * 1024, "kernel parameter space");

// ---------------------------------------------------------------------------
// Sinks: what one sample does to the grid.
// ---------------------------------------------------------------------------

// Global-space atomics spelled out: after a pointer has been pinned into registers (asm volatile) the compiler
// no longer knows its address space and would emit the slower generic ATOM.
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_u32(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// COUNT32: plain u32 hit counter, clamped later.  `red.global.add.u32` (no return).
struct SinkCount32 {
    uint32_t* counts;
    template <int SLOT = 0>
    __device__ __forceinline__ void put(uint32_t idx) { red_add_u32(counts + idx, 1u); }
    __device__ __forceinline__ void finish() {}
    __device__ __forceinline__ void words_pin() { asm volatile("" : "+l"(counts)); }
};

// PACKED8: the u8 output grid itself is the counter; four voxels share one
// 32-bit word and a hit adds 1 << (8 * byte).  The returned old word tells
// the adding thread whether ITS add carried out of the byte (old field ==
// 255).  The first carry in a word is always seen on a clean word, so a word
// is flagged in the bitmap if and only if one of its voxels received more
// than 255 hits; flagged words are recounted exactly by k_repair_packed.
// Returned words are examined kDepth samples later (a ring of kDepth pending
// results in registers, slot chosen at compile time by the unrolled walk), so
// up to kDepth atomics of a thread are in flight and their round trips overlap
// the arithmetic of the following samples.
// (A fire-and-forget `red` with a byte-sum verification pass was measured in
// this x-fastest layout as well -- DESIGN.md 6.4: the walk is no faster, because
// the cost of a sample is its 32-byte request packet to L2 either way, and the
// extra pass over the grid is not free.  In the brick layout it is, and BRICK8
// below works that way.)
struct SinkPacked8 {
    static constexpr int kDepth = 4;
    uint32_t* words;
    uint32_t* ovf_bitmap;
    uint32_t* ovf_flag;
    uint32_t pend_old[kDepth] = {0, 0, 0, 0};
    uint32_t pend_idx[kDepth] = {0, 0, 0, 0};
    template <int SLOT>
    __device__ __forceinline__ void check() {
        // byte (idx & 3) of the word as it was before this thread's add
        if (__byte_perm(pend_old[SLOT], 0u, 0x4440u | (pend_idx[SLOT] & 3u)) == 0xFFu) {
            const uint32_t w = pend_idx[SLOT] >> 2;
            atomicOr(ovf_bitmap + (w >> 5), 1u << (w & 31u));
            *ovf_flag = 1u;
        }
    }
    template <int SLOT>
    __device__ __forceinline__ void put(uint32_t idx) {
        check<SLOT>();
        pend_old[SLOT] = atom_add_u32(words + (idx >> 2), 1u << ((idx & 3u) * 8u));
        pend_idx[SLOT] = idx;
    }
    __device__ __forceinline__ void finish() { check<0>(); check<1>(); check<2>(); check<3>(); }
    __device__ __forceinline__ void words_pin() { asm volatile("" : "+l"(words)); }
};

// BRICK8: the same counter words, stored brick by brick in a scratch volume (brick_word in walk.cuh), and added
// with a fire-and-forget `red` -- no returned word, no pending ring, no wait for a round trip.  Whether a byte ever
// carried is decided once per instance instead: the walk counts the samples it added, the copy-out (k_untile_batch)
// sums the bytes it reads (slotted partial sums in `stats`).  Every add raises the byte sum by exactly 1 unless it carries
// out of its byte, and then by less (-254 into the next byte, -255 out of the word), so the two numbers are equal if
// and only if no voxel received more than 255 hits -- in which case the bytes ARE the counts.  Otherwise
// k_brick_verdict flags the instance and k_repair_packed recounts it in u32 and rewrites its whole volume.
// In the x-fastest layout the same `red` was slower than the returning `atom` (a `red` that misses L2
// stalls its slice); in the brick layout it is faster: 1.19 -> 0.97 ms for the crowd walk.
// (Moving rare paths into __noinline__ functions was measured too: the call ABI costs spills in the tile loop.)
// POW2: W and H are powers of two and `wh` holds their logarithms -- a linear index is decomposed by bit fields, which
// is also how the fast path of large grids finds its brick (walk.cuh, EXACT == 4).
template <bool POW2>
struct SinkPacked8Brick {
    uint32_t* words;                // the brick-ordered scratch
    unsigned long long* stats;
    uint32_t wh;                    // (W / 4) | (H / 4) << 16 (only the literal path needs it); POW2: log2 W | log2 H << 8
    uint32_t wb, hb;                // bricks per row (W / 4), brick rows per slab (H / 4): put_xyz
    uint32_t added = 0;             // samples this lane added
    template <int SLOT>
    __device__ __forceinline__ void put_brick(uint32_t lin, uint32_t bword) {
#if VKHR_PROBE_NO_RED          // measurement build (tools/): the walk with its address arithmetic but without the red itself
        if (bword == 0xFFFFFFFFu)
#endif
        red_add_u32(words + bword, 1u << ((lin & 3u) * 8u));
        ++added;
    }
    // a voxel inside the grid by its coordinates (the int32-index walk, EXACT == 3): two IMADs name the brick
    template <int SLOT>
    __device__ __forceinline__ void put_xyz(int ix, int iy, int iz) {
        const uint32_t brick = ((uint32_t)(iz >> 1) * hb + (uint32_t)(iy >> 2)) * wb + (uint32_t)(ix >> 2);
        put_brick<SLOT>((uint32_t)ix, brick * 8u + ((uint32_t)(iz & 1) * 4u + (uint32_t)(iy & 3)));
    }
    // the same voxel as ONE number, (brick word << 2) | byte -- below 2^24 on the grids of the int32-index walk (at most
    // 2^24 voxels) -- so that another lane can add it (walk_interior_tile_dealt)
    __device__ __forceinline__ uint32_t pack_xyz(int ix, int iy, int iz) const {
        const uint32_t brick = ((uint32_t)(iz >> 1) * hb + (uint32_t)(iy >> 2)) * wb + (uint32_t)(ix >> 2);
        return ((brick * 8u + ((uint32_t)(iz & 1) * 4u + (uint32_t)(iy & 3))) << 2) | ((uint32_t)ix & 3u);
    }
    __device__ __forceinline__ void put_packed(uint32_t a) { put_brick<0>(a, a >> 2); }
    // a red that carries several samples of one word (`value` = the sum of their byte increments); `n` of them are this lane's own
    __device__ __forceinline__ void put_packed_value(uint32_t a, uint32_t value, uint32_t n) { red_add_u32(words + (a >> 2), value); added += n; }
    __device__ __forceinline__ void count(uint32_t n) { added += n; }
    template <int SLOT>
    __device__ __forceinline__ void put_linear(uint32_t lin) { put_brick<SLOT>(lin, brick_word_pow2(lin, wh & 0xFFu, wh >> 8)); }
    template <int SLOT>
    __device__ __forceinline__ void put(uint32_t lin) {            // literal path: voxel from the linear index
        if (POW2) { put_linear<SLOT>(lin); return; }
        const uint32_t W = (wh & 0xFFFFu) << 2, H = (wh >> 16) << 2;
        const uint32_t x = lin % W, t = lin / W, y = t % H, z = t / H;
        const uint32_t brick = ((z >> 1) * (H >> 2) + (y >> 2)) * (W >> 2) + (x >> 2);
        put_brick<SLOT>(lin, (brick << 3) | ((z & 1u) << 2) | (y & 3u));
    }
    // all 32 lanes of the warp call finish() together (the kernels return early only as whole warps)
    __device__ __forceinline__ void finish() {
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, added);
        if ((threadIdx.x & 31u) == 0u && total)
            atomicAdd(stats + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (kStatSlots - 1u)), (unsigned long long)total);
        added = 0;
    }
    __device__ __forceinline__ void words_pin() { asm volatile("" : "+l"(words)); asm volatile("" : "+r"(wb)); asm volatile("" : "+r"(hb)); }
};

// Recount pass of PACKED8 / BRICK8: only samples landing in flagged words of the voxel chunk [c0, c0 + span) are
// counted, into the u32 chunk scratch (k_repair_packed).
struct SinkRecount {
    const uint32_t* ovf_bitmap;     // nullptr: every word is recounted (BRICK8: the byte sum did not match)
    uint32_t* counts;               // u32 per voxel of the chunk
    uint32_t c0, span;
    template <int SLOT = 0>
    __device__ __forceinline__ void put(uint32_t idx) {
        const uint32_t rel = idx - c0, w = idx >> 2;
        if (rel < span && (!ovf_bitmap || ((__ldg(ovf_bitmap + (w >> 5)) >> (w & 31u)) & 1u))) atomicAdd(counts + rel, 1u);
    }
    __device__ __forceinline__ void finish() {}
};

template <int MODE> struct SinkOf;
template <> struct SinkOf<0> { using type = SinkCount32;
    __device__ static type make(const InstanceDev& I) { return SinkCount32{I.counts}; } };
template <> struct SinkOf<1> { using type = SinkPacked8;
    __device__ static type make(const InstanceDev& I) { SinkPacked8 k; k.words = reinterpret_cast<uint32_t*>(I.densities); k.ovf_bitmap = I.ovf_bitmap; k.ovf_flag = I.ovf_flag; return k; } };

template <> struct SinkOf<3> { using type = SinkPacked8Brick<false>;
    __device__ static type make(const InstanceDev& I) {
        type k; k.words = reinterpret_cast<uint32_t*>(I.brick); k.stats = I.stats;
        k.wh = (I.grid.W >> 2) | ((I.grid.H >> 2) << 16); k.wb = I.grid.W >> 2; k.hb = I.grid.H >> 2; return k; } };
template <> struct SinkOf<4> { using type = SinkPacked8Brick<true>;
    __device__ static type make(const InstanceDev& I) {
        type k; k.words = reinterpret_cast<uint32_t*>(I.brick); k.stats = I.stats;
        k.wh = (31u - (uint32_t)__clz((int)I.grid.W)) | ((31u - (uint32_t)__clz((int)I.grid.H)) << 8); k.wb = I.grid.W >> 2; k.hb = I.grid.H >> 2; return k; } };

// ---------------------------------------------------------------------------
// Walk kernel for uniform strands (no index buffer): the hot kernel.
//
// Warp-autonomous and barrier-free.  A warp-tile is 32 consecutive vertices
// (96 floats, three fully coalesced 128-byte requests); lane t moves vertex t to
// voxel space ONCE (three exact divisions), receives vertex t+1 from lane t+1
// by shuffle and walks the segment (t, t+1) unless t is the last vertex of its
// strand; lane 31 only supplies the tip of lane 30, so consecutive tiles overlap
// by one vertex (tile stride 31).  Each warp walks kTilesPerWarp tiles and
// fetches the next tile's floats into registers before walking the current one,
// so the HBM latency of the vertex stream hides behind the walk.  Vertices are
// read from HBM once (1/31 twice); per-instance constants come from the constant
// bank (the batch is a __grid_constant__ parameter); the L1TEX path carries only
// the vertex stream and the reds.  blockIdx.y + first = instance.
// MODE 0 = COUNT32, 1 = PACKED8, 3 = BRICK8 on small grids (EXACT = 3), 4 = BRICK8 on power-of-two grids of any size (EXACT = 4).
// ---------------------------------------------------------------------------
// The interior walk of a tile: 0 = one segment per lane (the default); 1 = the tile's samples dealt out again so that one
// red instruction carries both samples of 16 consecutive segments; 2 = neighbour absorption (walk.cuh).  Both
// alternatives were built to cut request packets and do (120 M -> 114 M / 98 M red sectors per crowd frame), both
// are bit-exact on the whole GPU suite, and both are SLOWER (1.00 -> 1.08 / 1.09 ms): the extra shuffles and selects
// cost more issue slots than the packets saved (profiles/r02_ac_*, r02_ad_*).
#ifndef VKHR_WALK_DEALT
#define VKHR_WALK_DEALT 0
#endif
#ifndef VKHR_TILES_PER_WARP
#define VKHR_TILES_PER_WARP 8                                    // a multiple of 4 (whole 16-byte units per range); 16 fits 48 KB of static shared memory with one stage
#endif
constexpr uint32_t kTilesPerWarp = VKHR_TILES_PER_WARP;
constexpr uint32_t kWarpsPerBlock = kWalkThreads / 32;
constexpr uint32_t kTileStride = 31;          // segments (= new vertices) per warp-tile
constexpr uint32_t kRangeFloats = 3u * kTileStride * kTilesPerWarp;   // 8 tiles: 744 floats = 2976 bytes (a multiple of 16)
constexpr uint32_t kNeedFloats = kRangeFloats + 3u;                   // + the tip vertex of the range's last segment
constexpr uint32_t kBulkBytes = ((kNeedFloats * 4u + 15u) / 16u) * 16u;   // 3008: what one bulk copy moves
constexpr uint32_t kStageFloats = kBulkBytes / 4u;   // shared-memory slot of one warp (752 floats for 8 tiles; two per warp must fit 48 KB of static shared memory)
constexpr uint32_t kStagePad = 96u;           // floats behind the last warp's slot: the tile loop fetches one tile ahead, also behind the last tile
static_assert(kRangeFloats * 4u % 16u == 0 && kBulkBytes <= kStageFloats * 4u, "bulk copy geometry");
static_assert(kRangeFloats + 3u * 32u <= kStageFloats + kStagePad, "the look-ahead fetch stays inside the padded stage");

// mbarrier + 1-D bulk copy (TMA unit, `cp.async.bulk`, SASS UBLKCP): global -> shared without passing
// through the LSU queue the reds occupy.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// The vertices are read exactly once: the copy carries an evict_first L2 policy, so that the 1.36 GB of vertices that
// stream through a crowd frame do not push the frame kernel's scratch ring out of L2 (measured: 1.087 -> 1.027 ms per
// frame, DRAM traffic 3.77 -> 3.29 GB; evict_last on the ring's reds / reads and evict_first policies on the output
// stores were measured beside it and add nothing: profiles/r02_st_l2_hints.json).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    unsigned long long once;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(once));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(once) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();                           // a lost copy must fail, not hang the device
    }
}

template <class T> __device__ __forceinline__ T* pin(T* p) { asm volatile("" : "+l"(p)); return p; }

// A warp's range of an instance with uniform strands: 8 tiles of 31 segments (2976 bytes of vertices + one tip vertex).
// Two steps, so that a persistent kernel can request the NEXT range's vertices before it walks the current one.
//
// stage_range: start moving the range's vertices into `stage`.  16-byte aligned vertex buffers: ONE bulk copy issued
// by lane 0, completion on the mbarrier `bar`; anything else (a 4-byte aligned view, the last bytes of the buffer):
// coalesced 32-bit loads.  Returns whether a bulk copy is in flight (the walk then waits for `bar`).
__device__ __forceinline__ bool stage_range(const InstanceDev& I, uint32_t range, float* stage, uint32_t bar) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_floats = 3u * I.n_vertices;
    const float* __restrict__ verts = I.vertices;
    const uint32_t start = kRangeFloats * range;                // first float of the range
    // the common case, decided by a handful of warp-uniform instructions: a whole range and its tip vertex inside a
    // 16-byte aligned buffer -- one bulk copy of constant size, nothing to fill in by hand
    if (start + kStageFloats <= n_floats && (reinterpret_cast<uintptr_t>(verts) & 15u) == 0u) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(smem_u32(stage), verts + start, kBulkBytes, bar);
        }
        return true;
    }
    if (start >= n_floats) return false;
    const uint32_t need = min(kNeedFloats, n_floats - start);   // floats this warp reads from `stage`
    uint32_t bulk = 0;                                          // floats that arrive by bulk copy
    if ((reinterpret_cast<uintptr_t>(verts) & 15u) == 0u)
        bulk = min(kBulkBytes, ((n_floats - start) * 4u) & ~15u) / 4u;
    if (bulk && lane == 0) {
        // (a persistent kernel reuses the slot: earlier generic-proxy accesses are ordered before the async-proxy write)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_load(smem_u32(stage), verts + start, bulk * 4u, bar);
    }
    for (uint32_t j = bulk + lane; j < need; j += 32u) stage[j] = __ldg(verts + start + j);
    // floats past the end of the vertex buffer (last warp of an instance only) are never walked, but their
    // lanes take part in the warp's votes: give them a harmless value instead of stale shared memory
    for (uint32_t j = need + lane; j < kNeedFloats; j += 32u) stage[j] = I.grid.ox;
    return bulk != 0u;
}

// walk_range: the walk of a staged range.  `parity`: the phase of `bar` the bulk copy completes (flipped here); the
// sink is the caller's (it calls finish()).  All 32 lanes call this together.
template <int EXACT, bool FAST, class Sink>
__device__ __forceinline__ void walk_range(const InstanceDev& I, const GridParams& g, uint32_t range, float* stage, uint32_t bar, bool bulk, uint32_t& parity, Sink& sink) {
    // FAST (the caller's g.fast_div, warp-uniform, decided once per CTA instead of twice per tile): the FMA division
    // applies to this instance's voxel sizes, which is what the unguarded interior walk needs
    // g: the caller's pin(I.grid) -- registers, not indexed constant loads, and loaded once per CTA, not once per range
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_vertices = I.n_vertices;
    const uint32_t n_warp_tiles = (n_vertices + kTileStride - 1u) / kTileStride;
    const uint32_t tile0 = range * kTilesPerWarp;
    const uint32_t n_tiles = min(kTilesPerWarp, n_warp_tiles - min(tile0, n_warp_tiles));
    if (bulk) { mbar_wait(bar, parity); parity ^= 1u; }
    __syncwarp();
    if (n_tiles == 0) return;                                      // whole warps only

    // ---- which of this lane's vertices start a segment ---------------------------------------------------------
    // vertex x starts a segment unless it is the last of its strand: r = x mod (segs + 1) (one multiply-high division
    // per range, then += 31 mod (segs + 1) per tile) is segs.  Lane 31 only supplies the tip of lane 30: it is parked on
    // r = segs.  No test against n_vertices: n_vertices is a multiple of segs + 1 (checked on the host), so the
    // instance's last vertex ends a strand, and the lanes past it hold stage_range's filler -- one and the same point,
    // segments of zero steps, which add nothing on either path of the walk.
    // (sp1 = r + 1 in [1, vps] is what is kept, so that the test and the wrap need vps alone; vps and the step are pinned
    // in registers: left to itself the compiler reloads them from the indexed constant bank for every tile)
    const uint32_t vps = pin(I.segs_per_strand + 1u);
    uint32_t sp1, r_step;
    {
        const uint32_t x = kTileStride * tile0 + lane;
        uint32_t r = x - __umulhi(x, I.vps_magic) * vps;            // vps_magic = floor(2^32 / vps) + 1: quotient exact or one too large
        if ((int32_t)r < 0) r += vps;
        const bool tip_only = lane >= kTileStride;
        sp1 = tip_only ? vps : r + 1u;
        r_step = pin(tip_only ? 0u : I.tile_step);                  // 31 mod (segs + 1), from the host
    }

    // ---- software-pipelined tile loop ------------------------------------------------------------------
    // saddr: shared-memory address of this lane's vertex of the tile being fetched (a tile is 93 floats further), kept
    // as a pinned 32-bit shared address (left to itself the compiler rebuilds it from %tid for every tile).  The raw
    // floats of tile k+1 are requested before tile k is walked and transformed right after it: the shared-memory
    // round trip hides behind the walk, and only three registers of look-ahead stay live across it.
    uint32_t saddr = pin(smem_u32(stage) + 12u * lane);
    float r0, r1, r2, px, py, pz, tx, ty, tz;
    auto fetch = [&]() { r0 = lds_f32(saddr); r1 = lds_f32(saddr + 4u); r2 = lds_f32(saddr + 8u); saddr += 12u * kTileStride; };
    auto shuffle_tips = [&]() {                                    // lane t's tip is lane t + 1's vertex
        tx = __shfl_down_sync(kFullWarp, px, 1);
        ty = __shfl_down_sync(kFullWarp, py, 1);
        tz = __shfl_down_sync(kFullWarp, pz, 1);
    };
    constexpr bool kInterior = EXACT == 3;                          // the unguarded interior walk exists for the int32-index brick sink
    constexpr bool fast_div = kInterior && FAST;
    auto transform = [&]() {
        if (fast_div) {
            // unguarded: validated by vertex_is_interior on the RESULT (walk.cuh); a tile that fails the vote is redone below
            px = div_fast(__fsub_rn(r0, g.ox), g.vsx, g.rvx);
            py = div_fast(__fsub_rn(r1, g.oy), g.vsy, g.rvy);
            pz = div_fast(__fsub_rn(r2, g.oz), g.vsz, g.rvz);
        } else {
            to_voxel_space_warp(g, r0, r1, r2, px, py, pz);
        }
        shuffle_tips();
    };
    fetch();
    transform();
    for (uint32_t k = n_tiles;;) {
        fetch();                                                   // (behind the last tile: up to 96 floats past the range -- the stage is padded)
        const bool active = sp1 != vps;
        sp1 += r_step;
        if (sp1 > vps) sp1 -= vps;
        bool walked = false;
        if constexpr (fast_div) {
            const float dx = __fsub_rn(tx, px), dy = __fsub_rn(ty, py), dz = __fsub_rn(tz, pz);
            const float steps = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));   // no NaN here when the vote passes
            const bool ok = vertex_is_interior(g, px, py, pz) && (!active || steps < 2048.0f);
            if (__all_sync(kFullWarp, ok)) {
#if VKHR_WALK_DEALT == 2
                walk_interior_tile_absorb(active && steps > 0.0f, px, py, pz, dx, dy, dz, steps, sink);
#elif VKHR_WALK_DEALT == 1
                walk_interior_tile_dealt(active && steps > 0.0f, px, py, pz, dx, dy, dz, steps, sink);
#else
                if (active && steps > 0.0f) walk_interior_lane(px, py, pz, dx, dy, dz, steps, sink);
#endif
                walked = true;
            }
        }
        if (!walked) {
            if (fast_div) {
                // a vertex on or outside the faces of the box, a NaN, a very long segment: the guarded code, from the
                // tile's raw floats (still in the stage: 93 floats per tile)
                const uint32_t a = smem_u32(stage) + 12u * lane + 12u * kTileStride * (n_tiles - k);
                to_voxel_space_warp(g, lds_f32(a), lds_f32(a + 4u), lds_f32(a + 8u), px, py, pz);
                shuffle_tips();
            }
            walk_voxel_space_warp<EXACT, false>(g, active, px, py, pz, tx, ty, tz, sink);
        }
        // the lanes without a segment skipped the walk: meet them here.  (Left to itself the compiler reconverges
        // behind the transform and runs its 19 instructions once for each half of the warp: 7 % of the frame kernel's
        // instructions, ncu source page of round 2.)
        __syncwarp();
        if (--k == 0u) break;
        transform();
    }
}

template <int MODE, int EXACT>
__global__ void __launch_bounds__(kWalkThreads, VKHR_WALK_MIN_CTAS)
k_walk_uniform(const __grid_constant__ Batch B, uint32_t first) {
    __shared__ __align__(128) float s_stage_flat[kWarpsPerBlock * kStageFloats + kStagePad];
    float (*s_stage)[kStageFloats] = reinterpret_cast<float (*)[kStageFloats]>(s_stage_flat);
    __shared__ __align__(8) unsigned long long s_bar[kWarpsPerBlock];
    const InstanceDev& I = B.inst[first + blockIdx.y];
    if (blockIdx.x >= I.n_tiles || I.kind != WK_UNIFORM) return;          // n_tiles counts CTAs
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t bar = smem_u32(&s_bar[warp]);
    if ((threadIdx.x & 31u) == 0u) mbar_init(bar, 1);
    __syncwarp();
    uint32_t parity = 0;
    auto sink = SinkOf<MODE>::make(I);
    sink.words_pin();
    const uint32_t range = blockIdx.x * kWarpsPerBlock + warp;     // this warp's range of kTilesPerWarp tiles
    const bool bulk = stage_range(I, range, s_stage[warp], bar);
    const GridParams g = pin(I.grid);
    if (g.fast_div) walk_range<EXACT, true>(I, g, range, s_stage[warp], bar, bulk, parity, sink);
    else walk_range<EXACT, false>(I, g, range, s_stage[warp], bar, bulk, parity, sink);
    sink.finish();
}

// ---------------------------------------------------------------------------
// Generic walk: explicit index buffer (arbitrary vertex pairs), one thread per
// segment, gathered vertex loads.
// ---------------------------------------------------------------------------
// Segment `s` of an indexed instance for this lane (all 32 lanes call it together).
template <int EXACT, class Sink>
__device__ __forceinline__ void walk_indexed_lane(const InstanceDev& I, uint64_t s, Sink& sink) {
    const bool active = s < I.n_segments;
    const GridParams& g = I.grid;
    float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    if (active) {
        const uint2 pr = __ldg(reinterpret_cast<const uint2*>(I.indices) + s);
        if (pr.x < I.n_vertices && pr.y < I.n_vertices) {          // an index past the vertex array (a corrupt file) is dropped, not read
            const float* a = I.vertices + 3ull * pr.x;
            const float* b = I.vertices + 3ull * pr.y;
            ax = __ldg(a); ay = __ldg(a + 1); az = __ldg(a + 2);
            bx = __ldg(b); by = __ldg(b + 1); bz = __ldg(b + 2);
        }
    }
    float px, py, pz, tx, ty, tz;
    to_voxel_space_warp(g, ax, ay, az, px, py, pz);
    to_voxel_space_warp(g, bx, by, bz, tx, ty, tz);
    walk_voxel_space_warp<EXACT>(g, active, px, py, pz, tx, ty, tz, sink);
}

template <int MODE, int EXACT>
__global__ void __launch_bounds__(kWalkThreads)
k_walk_indexed(const __grid_constant__ Batch B, uint32_t first) {
    const InstanceDev& I = B.inst[first + blockIdx.y];
    if (blockIdx.x >= I.n_tiles || I.kind != WK_INDEXED) return;
    auto sink = SinkOf<MODE>::make(I);
    walk_indexed_lane<EXACT>(I, (uint64_t)blockIdx.x * kWalkThreads + threadIdx.x, sink);
    sink.finish();
}

// ---------------------------------------------------------------------------
// The frame kernel: BRICK8 walk AND copy-out of a whole batch in ONE launch.
//
// The grid is k_walk_uniform's: blockIdx.y = instance, blockIdx.x = walk item (8 warp-ranges of 8 tiles, kFrameRanges of
// them in a row; 2048 segments of an indexed instance), so instances START in order.  A CTA walks its item, adds its
// sample count to the instance's statistics and reports to the instance's counter (one fence, by one thread, behind the
// CTA's barrier).  The last `copiers` CTAs of instance i + 1 then copy out instance i -- the scratch volume to the
// caller's x-fastest volume, an equal share of the bricks each, zeroing behind themselves -- while every other CTA slot
// of the machine walks on.  Every CTA of instance i was dispatched before them and has almost always reported by the
// time they come off their own walk (each warp looks at the counter as it finishes; the CTA's report barrier is the
// vote), so no CTA slot is held by a copier that waits.  Only the batch's last instance is copied out by its own last
// `copiers_last` CTAs, which wait for its walk.  Instance i counts in scratch slot i mod `ring` (the pointer comes from
// the host): a ring of four slots (64 MiB at 256^3) covers the window of instances in flight.
//
// Waiting: a copier waits for CTAs of the instance before its own; a walk CTA of instance i waits (with its vertex copy
// already in flight) until instance i - ring has been copied out (by CTAs of instance i - ring + 1).  Everything waited
// for has a smaller block index, and CTAs are dispatched in increasing linear block index (the order every
// spin-on-the-previous-block scheme relies on -- serial split-K semaphores, decoupled look-back with block-index
// tickets), so it is resident and running.  Waits are bounded: a lost dependency traps instead of hanging the device.
// The verdict of the fire-and-forget `red` walk (samples added == byte sum, see SinkPacked8Brick) is taken by an
// instance's last copier.  The control block of the NEXT call is zeroed here (two blocks alternate), so a frame is one
// launch (+ the repair kernel's look at the flags).
//
// Forms built and measured before this one (crowd frame, ms; separate kernels: 1.19; this one: 0.98-1.01): persistent
// CTAs drawing CTA-wide items from one ordered ticket queue with the copy-out as queue items 1.27; the same with
// autonomous warps and one item of look-ahead 2.1 (the tickets parked in look-ahead were the dependencies other warps
// span on); ticketed one-item CTAs whose last finishers copy out 1.14-1.5 (a fifth of all warp-time at the barrier behind
// the ticket); this grid with warp-wise reports and copiers 1.78 (a fence per warp: 15 % of all warp-time in the fence);
// this grid with an instance's OWN last 64 CTAs waiting for its walk and copying it out 1.03-1.08 (7 % of all warp-time
// at that wait, and the copy-out paced the ring).  profiles/r02_b_* ... r02_h_*, r02_y_*, r02_aa_*; what bounds this
// form: DESIGN.md 6.16.
// ---------------------------------------------------------------------------
constexpr uint32_t kFrameStatSlots = 32;
constexpr uint32_t kFrameIndexedSegs = 2048;                    // segments per walk item of an indexed instance
#ifndef VKHR_FRAME_RANGES
#define VKHR_FRAME_RANGES 3
#endif
#ifndef VKHR_FRAME_MIN_CTAS
#define VKHR_FRAME_MIN_CTAS 4
#endif
#ifndef VKHR_FRAME_STAGES
#define VKHR_FRAME_STAGES 1
#endif
#ifndef VKHR_FRAME_INFLIGHT
#define VKHR_FRAME_INFLIGHT 4                                   // bricks (2 x 16 bytes) a copier thread has in flight
#endif
constexpr uint32_t kFrameStages = VKHR_FRAME_STAGES;            // vertex stage buffers per warp
constexpr uint32_t kFrameRanges = VKHR_FRAME_RANGES;            // warp-ranges a warp walks per item: the item's fixed costs (fence, report) are paid once
struct FrameCtl {
    uint32_t walk_done[kMaxBatch];                              // CTAs of the instance that have reported
    uint32_t copy_done[kMaxBatch];                              // copiers of the instance that have finished
    unsigned long long added[kMaxBatch][kFrameStatSlots];       // samples the walk added, slotted
    unsigned long long bytes[kMaxBatch][kFrameStatSlots];       // byte sums the copy-out read, slotted
};
struct FramePlan {
    uint32_t ring;                                              // scratch slots
    uint32_t copiers;                                           // CTAs of instance i + 1 (its last ones by block index) that copy out instance i
    uint32_t n_bricks;
    uint32_t copiers_last;                                      // CTAs of the batch's LAST instance that copy it out themselves (nobody comes behind it)
    FrameCtl* ctl;
    FrameCtl* ctl_next;
};

// thread 0 waits until *p >= need (acquire), then the CTA meets at a barrier
__device__ __forceinline__ void frame_wait_ge(const uint32_t* p, uint32_t need) {
    if (threadIdx.x == 0) {
        uint32_t v;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            if (v >= need) break;
            __nanosleep(100);
            if (spin > (1u << 23)) __trap();                       // a lost dependency must fail, not hang the device
        }
    }
    __syncthreads();
}

template <int MODE, int EXACT>
__global__ void __launch_bounds__(kWalkThreads, VKHR_FRAME_MIN_CTAS)
k_frame(const __grid_constant__ Batch B, const __grid_constant__ FramePlan P) {
    // kFrameStages = 2: the vertices of a warp's next range are in flight while it walks the current one (measured: no
    // gain -- 1.12 ms against 1.08 ms per crowd frame with one buffer, which leaves twice the L1; profiles/r02_m_*)
    __shared__ __align__(128) float s_stage_flat[kFrameStages * kWarpsPerBlock * kStageFloats + kStagePad];
    float (*s_stage)[kWarpsPerBlock][kStageFloats] = reinterpret_cast<float (*)[kWarpsPerBlock][kStageFloats]>(s_stage_flat);
    __shared__ __align__(8) unsigned long long s_bar[kFrameStages][kWarpsPerBlock];
    __shared__ unsigned long long s_sum[kWarpsPerBlock];           // the walk's sample counts
    __shared__ unsigned long long s_csum[kWarpsPerBlock];          // the copy-out's byte sums (thread 0 may still be reading s_sum)
    const uint32_t i = blockIdx.y;
    const InstanceDev& I = B.inst[i];
    const uint32_t items = max(I.n_tiles, 1u);                     // CTAs of this instance (one for an instance without segments: its copy-out)
    if (blockIdx.x >= items) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t bar[2] = {smem_u32(&s_bar[0][warp]), smem_u32(&s_bar[kFrameStages - 1u][warp])};
    if (lane == 0) { mbar_init(bar[0], 1); if (kFrameStages > 1u) mbar_init(bar[1], 1); }
    __syncwarp();
    FrameCtl* const ctl = P.ctl;
    if (i == 0 && blockIdx.x == 0) {                               // the next call's control block (nobody uses it during this call)
        uint4* z = reinterpret_cast<uint4*>(P.ctl_next);
        for (uint32_t k = threadIdx.x; k < sizeof(FrameCtl) / 16u; k += blockDim.x) z[k] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t n_inst = gridDim.y;
    const bool uniform_item = blockIdx.x < I.n_tiles && I.kind == WK_UNIFORM;

    // ---- the walk ------------------------------------------------------------------------------------------------
    {
        auto sink = SinkOf<MODE>::make(I);
        sink.words_pin();
        const GridParams g = pin(I.grid);
        uint32_t parity[2] = {0, 0};
        // the first range's vertices are on their way while the slot is checked
        uint32_t range = blockIdx.x * kFrameRanges * kWarpsPerBlock + warp;
        bool bulk = uniform_item && stage_range(I, range, s_stage[0][warp], bar[0]);
        if (i >= P.ring)                                           // the slot's previous tenant, instance i - ring, has been copied out
            frame_wait_ge(&ctl->copy_done[i - P.ring], min(max(B.inst[i - P.ring + 1u].n_tiles, 1u), P.copiers));   // by CTAs of the instance behind it
        if (uniform_item) {
            // (not unrolled: one copy of the walk per FAST instead of kFrameRanges -- the kernel is instruction-cache sized)
#pragma unroll 1
            for (uint32_t rr = 0; rr < kFrameRanges; ++rr) {
                const uint32_t b = kFrameStages > 1u ? (rr & 1u) : 0u, nb = kFrameStages > 1u ? (b ^ 1u) : 0u;
                bool bulk_nxt = false;
                if (kFrameStages > 1u && rr + 1u < kFrameRanges)   // (buffer nb was last read two ranges ago, by this warp)
                    bulk_nxt = stage_range(I, range + kWarpsPerBlock, s_stage[nb][warp], bar[nb]);
                if (g.fast_div) walk_range<EXACT, true>(I, g, range, s_stage[b][warp], bar[b], bulk, parity[b], sink);
                else walk_range<EXACT, false>(I, g, range, s_stage[b][warp], bar[b], bulk, parity[b], sink);
                __syncwarp();
                range += kWarpsPerBlock;
                if (kFrameStages == 1u && rr + 1u < kFrameRanges) bulk_nxt = stage_range(I, range, s_stage[0][warp], bar[0]);
                bulk = bulk_nxt;
            }
        } else if (blockIdx.x < I.n_tiles) {
            for (uint32_t j = 0; j < kFrameIndexedSegs / kWalkThreads; ++j)
                walk_indexed_lane<EXACT>(I, ((uint64_t)blockIdx.x * (kFrameIndexedSegs / kWalkThreads) + j) * kWalkThreads + threadIdx.x, sink);
        }
        const uint32_t added = __reduce_add_sync(kFullWarp, sink.added);
        if (lane == 0) s_sum[warp] = added;
    }
    // A CTA that is going to copy out instance i - 1 (role 0 below) looks at that instance's counter NOW, every warp for
    // itself as it comes off its walk: the round trip of the acquire load hides behind the wait for the CTA's slowest
    // warp, and the barrier below doubles as the vote.  (Before: thread 0 polled after the barrier and seven warps
    // waited for its round trip at a second barrier -- 7 % of all warp-time, profiles/r02_aa_*.)
    bool ready = false;
    if (i > 0u && blockIdx.x + min(items, P.copiers) >= items) {
        uint32_t seen = 0;
        if (lane == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&ctl->walk_done[i - 1u]) : "memory");
        seen = __shfl_sync(kFullWarp, seen, 0);
        ready = seen >= max(B.inst[i - 1u].n_tiles, 1u);
    }
    const bool all_ready = __syncthreads_and(ready) != 0;
    if (threadIdx.x == 0) {
        unsigned long long sum = 0;
        for (uint32_t w = 0; w < kWarpsPerBlock; ++w) sum += s_sum[w];
        if (sum) atomicAdd(&ctl->added[i][blockIdx.x & (kFrameStatSlots - 1u)], sum);
        // release: the fence is cumulative over everything that happens-before it -- the reds of all threads of the CTA
        // are ordered before it by the barrier -- so the observer that acquires the counter sees them
        __threadfence();
        atomicAdd(&ctl->walk_done[i], 1u);
    }
    // ---- the copy-out: brick order -> the x-fastest output volume, zeroing behind itself -------------------------
    // Role 0: the last `copiers` CTAs of instance i copy out instance i - 1, an equal share of its bricks each.  Every
    // CTA of instance i - 1 was dispatched before this one and has (almost always) reported by now: no CTA slot is
    // held by a waiting copier, and the copy-out of an instance is spread over as many CTAs as the walk leaves time for.
    // Role 1: nobody comes behind the batch's last instance -- its own last `copiers_last` CTAs wait for its walk and
    // copy it out.
#pragma unroll 1
    for (uint32_t role = 0; role < 2u; ++role) {
        if (role == 0u ? i == 0u : i + 1u != n_inst) continue;
        const uint32_t t = role == 0u ? i - 1u : i;                // the instance copied out
        const uint32_t copiers = min(items, role == 0u ? P.copiers : P.copiers_last);
        if (blockIdx.x + copiers < items) continue;                // not one of instance i's last `copiers` CTAs
        const uint32_t c = blockIdx.x - (items - copiers);         // (the FIRST CTAs instead were measured: no gain, profiles/r02_ae_*)
        const InstanceDev& T = B.inst[t];
        if (!(role == 0u && all_ready))                            // (role 0: usually seen already, above)
            frame_wait_ge(&ctl->walk_done[t], max(T.n_tiles, 1u)); // every walk CTA of instance t has reported (acquire)
#if VKHR_PROBE_NO_COPY         // measurement build (tools/): the frame kernel without the copy-out's loads and stores
        if (P.n_bricks == 0xFFFFFFFFu)
#endif
        {
            const uint32_t wrow = T.grid.W >> 2, byn = T.grid.H >> 2;  // words (= bricks) per row, brick rows per slab
            const uint32_t wslab = wrow * T.grid.H;
            // brick number -> (bx, by, bz) by multiply-high: m = floor((2^32 - 1) / d) + 1 gives the quotient or one more
            // (then the remainder is negative); computed once per thread here -- in the loop the compiler re-derived both
            // divisions, reciprocals included, for every brick (a quarter of the copy-out's instructions)
            const uint32_t m_row = wrow > 1u ? 0xFFFFFFFFu / wrow + 1u : 0u;   // (a power of two: 2^32 / d itself, exact)
            const uint32_t m_byn = byn > 1u ? 0xFFFFFFFFu / byn + 1u : 0u;
            uint4* __restrict__ src = reinterpret_cast<uint4*>(T.brick);
            uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(T.densities);   // a volume of the frame kernel is at most 2^23 words
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint32_t bytes = 0;                                    // < 2^32: at most 8160 per brick
            const uint32_t per = ((P.n_bricks + copiers - 1u) / copiers + 31u) & ~31u;     // whole warp-rows of 32 bricks
            const uint32_t b_end = min((c + 1u) * per, P.n_bricks);
            // Four bricks per thread and round: ALL eight 16-byte loads are issued before the first of them is used
            // (as a plain unrolled loop each brick's loads stay behind the previous brick's stores to `src`).
            constexpr uint32_t kInFlight = VKHR_FRAME_INFLIGHT;
            for (uint32_t b0 = c * per + threadIdx.x; b0 < b_end; b0 += kInFlight * kWalkThreads) {
                uint4 q[kInFlight][2];
#pragma unroll
                for (uint32_t j = 0; j < kInFlight; ++j) {
                    const uint32_t b = b0 + j * kWalkThreads;
                    if (b < b_end) { q[j][0] = __ldcg(src + 2u * b); q[j][1] = __ldcg(src + 2u * b + 1u); }   // L2 is where the reds landed; L1 may be stale
                    else { q[j][0] = z; q[j][1] = z; }
                }
#pragma unroll
                for (uint32_t j = 0; j < kInFlight; ++j) {
                    const uint32_t b = b0 + j * kWalkThreads;
                    if (b >= b_end) break;
                    const uint4 q0 = q[j][0], q1 = q[j][1];
                    bytes = __dp4a(q0.x, 0x01010101u, bytes); bytes = __dp4a(q0.y, 0x01010101u, bytes);
                    bytes = __dp4a(q0.z, 0x01010101u, bytes); bytes = __dp4a(q0.w, 0x01010101u, bytes);
                    bytes = __dp4a(q1.x, 0x01010101u, bytes); bytes = __dp4a(q1.y, 0x01010101u, bytes);
                    bytes = __dp4a(q1.z, 0x01010101u, bytes); bytes = __dp4a(q1.w, 0x01010101u, bytes);
                    uint32_t tt = m_row ? __umulhi(b, m_row) : b, bx = b - tt * wrow;
                    if ((int32_t)bx < 0) { bx += wrow; --tt; }
                    uint32_t bz = m_byn ? __umulhi(tt, m_byn) : tt, by = tt - bz * byn;
                    if ((int32_t)by < 0) { by += byn; --bz; }
                    const uint32_t o = (2u * bz) * wslab + (4u * by) * wrow + bx;
                    __stcs(dst + o, q0.x); __stcs(dst + (o + wrow), q0.y); __stcs(dst + (o + 2u * wrow), q0.z); __stcs(dst + (o + 3u * wrow), q0.w);
                    const uint32_t o1 = o + wslab;
                    __stcs(dst + o1, q1.x); __stcs(dst + (o1 + wrow), q1.y); __stcs(dst + (o1 + 2u * wrow), q1.z); __stcs(dst + (o1 + 3u * wrow), q1.w);
                    if ((q0.x | q0.y | q0.z | q0.w | q1.x | q1.y | q1.z | q1.w) != 0u) { __stcg(src + 2u * b, z); __stcg(src + 2u * b + 1u, z); }
                }
            }
            const unsigned long long wsum = (unsigned long long)__reduce_add_sync(kFullWarp, bytes & 0xFFFFu) +
                                            ((unsigned long long)__reduce_add_sync(kFullWarp, bytes >> 16) << 16);
            if (lane == 0) s_csum[warp] = wsum;
        }
        __syncthreads();                                           // every thread's stores are issued, s_csum is complete
        if (warp == 0) {                                           // (the other warps are done: nothing below concerns them)
            uint32_t last = 0;
            if (lane == 0) {
                unsigned long long sum = 0;
                for (uint32_t w = 0; w < kWarpsPerBlock; ++w) sum += s_csum[w];
                if (sum) atomicAdd(&ctl->bytes[t][c & (kFrameStatSlots - 1u)], sum);
                __threadfence();                                   // the zeros (of every thread: barrier above) are in place before the slot is released
                last = (atomicAdd(&ctl->copy_done[t], 1u) == copiers - 1u) ? 1u : 0u;
            }
            if (__shfl_sync(kFullWarp, last, 0)) {
                // instance t is complete: samples added != byte sum of the volume means some byte carried (more than 255
                // hits in a voxel) -> flag 2, k_repair_packed recounts the instance in u32
                __threadfence();
                unsigned long long a = *(volatile unsigned long long*)&ctl->added[t][lane], y = *(volatile unsigned long long*)&ctl->bytes[t][lane];
                for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(kFullWarp, a, o); y += __shfl_down_sync(kFullWarp, y, o); }
#if VKHR_PROBE_NO_RED || VKHR_PROBE_NO_COPY
                a = y;                                             // (probe builds produce no volume: never send them to the repair kernel)
#endif
                if (lane == 0) *T.ovf_flag = (a != y) ? 2u : 0u;
            }
        }
        if (role == 0u && i + 1u == n_inst) __syncthreads();       // s_csum is reused by role 1
    }
}

// Vertex splat (voxelize_vertices): one thread per vertex.
template <int MODE, int EXACT>
__global__ void __launch_bounds__(kWalkThreads)
k_splat_batch(const __grid_constant__ Batch B, uint32_t first) {
    const InstanceDev& I = B.inst[first + blockIdx.y];
    if (blockIdx.x >= I.n_tiles || I.kind != WK_SPLAT) return;
    const uint32_t i = blockIdx.x * kWalkThreads + threadIdx.x;
    const bool active = i < I.n_vertices;
    const GridParams& g = I.grid;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    if (active) {
        const float* v = I.vertices + 3ull * i;
        wx = __ldg(v); wy = __ldg(v + 1); wz = __ldg(v + 2);
    }
    float px, py, pz;
    to_voxel_space_warp(g, wx, wy, wz, px, py, pz);
    auto sink = SinkOf<MODE>::make(I);
    uint32_t idx;
    if (active && voxel_index<EXACT>(g, px, py, pz, idx)) sink.template put<0>(idx);
    sink.finish();
}

// ---------------------------------------------------------------------------
// Tangent volume (Volume::tangents, hair_style.cc:309,:323,:331-339): per voxel the
// mean of the tangents of the samples that hit it, times 127, truncated to int8.
//
// The reference adds fp32 tangents in strand order (and shuffles the strands with a
// random seed first, scene_graph.cc:233), so its result is order-dependent; here
// the sum is an INTEGER sum of tangents quantised to 1/8192, which is the same for
// any order and any sharding.  Accumulator: 16 bytes per voxel, two u64 words
//   w0 = count | (sum_x << 32)            w1 = (sum_y + 8192*count) | (sum_z << 32)
// so a sample is two 64-bit reds (the high fields wrap modulo 2^32, which is exact
// for signed sums; the low field of w1 is biased to stay non-negative so that it
// never borrows from sum_z).  Exact up to 262,143 hits per voxel.
// ---------------------------------------------------------------------------
constexpr int kTangentScale = 8192;

// Sparse form (bits != nullptr): accumulators exist only for the voxels the density pass found non-empty -- slot of
// voxel idx = prefix[idx / 32] + popcount of the lower bits of bits[idx / 32] -- so the scratch is 16 bytes per NON-EMPTY
// voxel (35 MB for a ponytail at 256^3, 83 MB at 1024^3) instead of 16 bytes per voxel (256 MiB / 16 GiB).
struct SinkTangent {
    unsigned long long* acc;
    const uint32_t* bits = nullptr;       // one bit per voxel: density != 0
    const uint32_t* prefix = nullptr;     // non-empty voxels before each 32-voxel word
    unsigned long long w0 = 0, w1 = 0;
    __device__ __forceinline__ void set(float tx, float ty, float tz) {
        auto q = [](float t) { return max(-kTangentScale, min(kTangentScale, __float2int_rn(t * (float)kTangentScale))); };   // NaN -> 0
        const int qx = q(tx), qy = q(ty), qz = q(tz);
        w0 = 1ull | ((unsigned long long)(uint32_t)qx << 32);
        w1 = (unsigned long long)(uint32_t)(qy + kTangentScale) | ((unsigned long long)(uint32_t)qz << 32);
    }
    template <int SLOT = 0>
    __device__ __forceinline__ void put(uint32_t idx) {
        unsigned long long slot = idx;
        if (bits) {
            const uint32_t w = idx >> 5, b = __ldg(bits + w);
            if (!((b >> (idx & 31u)) & 1u)) return;                 // (cannot happen: the density pass counted this very sample)
            slot = __ldg(prefix + w) + __popc(b & ((1u << (idx & 31u)) - 1u));
        }
        atomicAdd(acc + 2ull * slot, w0);
        atomicAdd(acc + 2ull * slot + 1ull, w1);
    }
    __device__ __forceinline__ void finish() {}
};

// glm::normalize(v) = v * inversesqrt(dot(v, v)), what HairStyle::generate_tangents stores (hair_style.cc:171-194).
__device__ __forceinline__ void glm_normalize(float& x, float& y, float& z) {
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const float r = __fdiv_rn(1.0f, __fsqrt_rn(d));
    x = __fmul_rn(x, r); y = __fmul_rn(y, r); z = __fmul_rn(z, r);
}

// One thread per segment (KIND uniform / indexed) or per vertex (splat); counts and tangent sums into `acc`.
// tangents == nullptr (segments only): the tangent of a segment's root vertex is normalize(tip - root), which is
// what generate_tangents produces for every vertex that starts a segment.
template <int KIND, int EXACT>
__global__ void __launch_bounds__(kWalkThreads)
k_walk_tangent(const float* __restrict__ vertices, const uint32_t* __restrict__ indices, const float* __restrict__ tangents,
               uint64_t n_items, uint32_t n_vertices, uint32_t segs, const __grid_constant__ GridParams g, unsigned long long* __restrict__ acc,
               const uint32_t* __restrict__ bits, const uint32_t* __restrict__ prefix) {
    const uint64_t s = (uint64_t)blockIdx.x * kWalkThreads + threadIdx.x;
    bool active = s < n_items;
    uint32_t i0 = 0, i1 = 0;
    if (active) {
        if (KIND == WK_SPLAT) i0 = i1 = (uint32_t)s;
        else segment_vertices(KIND == WK_INDEXED ? indices : nullptr, segs, s, i0, i1);
        if (i0 >= n_vertices || i1 >= n_vertices) active = false;    // an index past the vertex array is dropped, not read
    }
    float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    SinkTangent sink{acc, bits, prefix};
    if (active) {
        const float* a = vertices + 3ull * i0;
        ax = __ldg(a); ay = __ldg(a + 1); az = __ldg(a + 2);
        if (KIND != WK_SPLAT) { const float* b = vertices + 3ull * i1; bx = __ldg(b); by = __ldg(b + 1); bz = __ldg(b + 2); }
        float tx, ty, tz;
        if (tangents) { const float* t = tangents + 3ull * i0; tx = __ldg(t); ty = __ldg(t + 1); tz = __ldg(t + 2); }
        else { tx = __fsub_rn(bx, ax); ty = __fsub_rn(by, ay); tz = __fsub_rn(bz, az); glm_normalize(tx, ty, tz); }
        sink.set(tx, ty, tz);
    }
    float px, py, pz;
    to_voxel_space_warp(g, ax, ay, az, px, py, pz);
    if (KIND == WK_SPLAT) {
        uint32_t idx;
        if (active && voxel_index<EXACT>(g, px, py, pz, idx)) sink.put(idx);
    } else {
        float qx, qy, qz;
        to_voxel_space_warp(g, bx, by, bz, qx, qy, qz);
        walk_voxel_space_warp<EXACT>(g, active, px, py, pz, qx, qy, qz, sink);
    }
}

// acc -> densities (min(count, 255)) and tangents (int8 x 4, w = 0); zeroes acc behind itself.
// tangent component = (int8) trunc( (sum / count) * 127 ), the expression of hair_style.cc:336-338 on the exact mean.
// Empty voxels: the reference divides 0/0 and converts NaN (0 on x86-64); here 0.
__global__ void __launch_bounds__(256)
k_finish_tangent(unsigned long long* __restrict__ acc, uint64_t n_voxels, uint8_t* __restrict__ dens, uint32_t* __restrict__ tang) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_voxels; i += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2* a = reinterpret_cast<ulonglong2*>(acc) + i;
        const ulonglong2 w = *a;
        const uint32_t count = (uint32_t)w.x;
        uint32_t t = 0;
        if (count) {
            const float fc = (float)count, inv = 1.0f / (float)kTangentScale;
            const int sx = (int)(uint32_t)(w.x >> 32);
            const int sy = (int)((uint32_t)w.y - (uint32_t)kTangentScale * count);
            const int sz = (int)(uint32_t)(w.y >> 32);
            auto q = [&](int sum) { return (uint32_t)(uint8_t)(int8_t)__float2int_rz(__fmul_rn(__fdiv_rn(__fmul_rn((float)sum, inv), fc), 127.0f)); };
            t = q(sx) | (q(sy) << 8) | (q(sz) << 16);
            *a = make_ulonglong2(0ull, 0ull);
        }
        dens[i] = (uint8_t)min(count, 255u);
        if (tang) tang[i] = t;
    }
}

__device__ __forceinline__ uint32_t tangent_of(ulonglong2 w) {
    const uint32_t count = (uint32_t)w.x;
    if (!count) return 0u;
    const float fc = (float)count, inv = 1.0f / (float)kTangentScale;
    const int sx = (int)(uint32_t)(w.x >> 32);
    const int sy = (int)((uint32_t)w.y - (uint32_t)kTangentScale * count);
    const int sz = (int)(uint32_t)(w.y >> 32);
    auto q = [&](int sum) { return (uint32_t)(uint8_t)(int8_t)__float2int_rz(__fmul_rn(__fdiv_rn(__fmul_rn((float)sum, inv), fc), 127.0f)); };
    return q(sx) | (q(sy) << 8) | (q(sz) << 16);
}

// Sparse tangent pass, step 1: one bit per voxel (density != 0) and the number of set bits per 32-voxel word (the
// exclusive scan of the counts gives every word its first accumulator slot).  One thread per word: two 16-byte loads.
__global__ void __launch_bounds__(256)
k_tangent_bitmap(const uint8_t* __restrict__ dens, uint32_t n_words, uint32_t* __restrict__ bits, uint32_t* __restrict__ counts) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint4* p = reinterpret_cast<const uint4*>(dens) + 2ull * w;
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // bit per non-zero byte: (x | x >> 4 ...) folded to the byte's lowest bit
        uint32_t x = q[k];
        x |= x >> 4; x |= x >> 2; x |= x >> 1; x &= 0x01010101u;
        m |= ((x & 1u) | ((x >> 7) & 2u) | ((x >> 14) & 4u) | ((x >> 21) & 8u)) << (4 * k);
    }
    bits[w] = m;
    counts[w] = (uint32_t)__popc(m);
}
// step 3: every voxel's int8 tangent from its accumulator (0 where the voxel is empty).
__global__ void __launch_bounds__(256)
k_finish_tangent_sparse(const unsigned long long* __restrict__ acc, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ prefix,
                        uint64_t n_voxels, uint32_t* __restrict__ tang) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_voxels; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(i >> 5), b = __ldg(bits + w), bit = (uint32_t)i & 31u;
        uint32_t t = 0;
        if ((b >> bit) & 1u) {
            const unsigned long long slot = __ldg(prefix + w) + __popc(b & ((1u << bit) - 1u));
            t = tangent_of(__ldg(reinterpret_cast<const ulonglong2*>(acc) + slot));
        }
        tang[i] = t;
    }
}

// ---------------------------------------------------------------------------
// Grid-stride helpers: clear, clamp, repair.
// ---------------------------------------------------------------------------

// Zero `n16` 16-byte words (grid-stride, st.global.v4).
__global__ void __launch_bounds__(256) k_zero16(uint4* __restrict__ p, uint64_t n16) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = z;
}

// PACKED8 clear for a batch: densities, overflow bitmap and flag of every instance.
// blockIdx.y + first = instance.  n_voxels % 16 == 0 is guaranteed by the host (else COUNT32).
// volumes == 0 (BRICK8): flag and statistics only -- the copy-out writes every byte of the densities, and the walk
// does not use the bitmap.
__global__ void __launch_bounds__(256) k_clear_packed_batch(const __grid_constant__ Batch B, uint32_t first, uint32_t volumes) {
    const InstanceDev& I = B.inst[first + blockIdx.y];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *I.ovf_flag = 0u;
    if (I.stats) for (uint32_t i = t; i < 2u * kStatSlots; i += stride) I.stats[i] = 0ull;
    if (!volumes) return;
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* d = reinterpret_cast<uint4*>(I.densities);
    const uint32_t n16 = I.grid.n_voxels >> 4;
    for (uint32_t i = t; i < n16; i += stride) d[i] = z;
    const uint32_t n_bm = (I.grid.n_voxels / 4 + 31) / 32;          // bitmap words
    for (uint32_t i = t; i < n_bm; i += stride) I.ovf_bitmap[i] = 0u;
}

// BRICK8 copy-out: brick-ordered scratch -> the x-fastest output volume (Volume::densities, hair_style.hh:89-101),
// one brick (two 16-byte loads) per thread; the 32 bricks of a warp are consecutive in x, so each of its eight
// 4-byte stores covers 128 contiguous bytes of one output row.  A brick that held anything is zeroed behind the
// read, which restores the scratch's all-zero state for the next call without a clear pass.  It also sums the bytes
// it reads (slotted partial sums): equal to the number of samples the walk added iff no byte carried (SinkPacked8Brick).
__global__ void __launch_bounds__(256) k_untile_batch(const __grid_constant__ Batch B, uint32_t first) {
    const InstanceDev& I = B.inst[first + blockIdx.y];
    const uint32_t wrow = I.grid.W >> 2, byn = I.grid.H >> 2;             // words (= bricks) per row, brick rows per slab
    const uint32_t wslab = wrow * I.grid.H;
    const uint32_t n_bricks = wrow * byn * (I.grid.D >> 1);
    uint4* __restrict__ src = reinterpret_cast<uint4*>(I.brick);
    uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(I.densities);
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t bytes = 0;                                                   // < 2^32: at most 8160 per brick, n_bricks / threads bricks
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_bricks; b += stride) {
        const uint4 q0 = __ldcs(src + 2u * b), q1 = __ldcs(src + 2u * b + 1u);   // z even, z odd: rows y & 3 = 0..3
        bytes += __vsadu4(q0.x, 0u) + __vsadu4(q0.y, 0u) + __vsadu4(q0.z, 0u) + __vsadu4(q0.w, 0u) +
                 __vsadu4(q1.x, 0u) + __vsadu4(q1.y, 0u) + __vsadu4(q1.z, 0u) + __vsadu4(q1.w, 0u);
        const uint32_t bx = b % wrow, t = b / wrow, by = t % byn, bz = t / byn;
        uint32_t* o = dst + (size_t)(2u * bz) * wslab + (size_t)(4u * by) * wrow + bx;
        __stcs(o, q0.x); __stcs(o + wrow, q0.y); __stcs(o + 2u * wrow, q0.z); __stcs(o + 3u * wrow, q0.w);
        o += wslab;
        __stcs(o, q1.x); __stcs(o + wrow, q1.y); __stcs(o + 2u * wrow, q1.z); __stcs(o + 3u * wrow, q1.w);
        if ((q0.x | q0.y | q0.z | q0.w | q1.x | q1.y | q1.z | q1.w) != 0u) { src[2u * b] = z; src[2u * b + 1u] = z; }
    }
    // the byte sum of the instance (the loop's trip count differs by at most one between lanes: no lane has left)
    const unsigned long long lo = __reduce_add_sync(0xFFFFFFFFu, bytes & 0xFFFFu), hi = __reduce_add_sync(0xFFFFFFFFu, bytes >> 16);
    if ((threadIdx.x & 31u) == 0u && (lo | hi))
        atomicAdd(I.stats + kStatSlots + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (kStatSlots - 1u)), lo + (hi << 16));
}

// BRICK8 verdict, one CTA of kStatSlots threads per instance: samples added != byte sum of the volume means some byte
// carried; the instance is flagged (2 = recount every word) for k_repair_packed.
__global__ void __launch_bounds__(kStatSlots) k_brick_verdict(const __grid_constant__ Batch B, uint32_t first) {
    const InstanceDev& I = B.inst[first + blockIdx.x];
    __shared__ unsigned long long s_sum[2][kStatSlots / 32];
    unsigned long long a = I.stats[threadIdx.x], b = I.stats[kStatSlots + threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xFFFFFFFFu, a, o); b += __shfl_down_sync(0xFFFFFFFFu, b, o); }
    if ((threadIdx.x & 31u) == 0u) { s_sum[0][threadIdx.x >> 5] = a; s_sum[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0; b = 0;
        for (uint32_t w = 0; w < kStatSlots / 32; ++w) { a += s_sum[0][w]; b += s_sum[1][w]; }
        if (a != b) *I.ovf_flag = 2u;
    }
}

// densities = min(counts, 255)  (hair_style.cc:322: `if (d != 255) d += 1`),
// 16 voxels per thread-iteration: 4 x ld.v4.u32 -> 1 x st.v4.u32.
// ZERO: also write zeros back over the counts (leaves the scratch clean for the next frame).
__device__ __forceinline__ uint32_t clamp4(uint4 c) {
    return min(c.x, 255u) | (min(c.y, 255u) << 8) | (min(c.z, 255u) << 16) | (min(c.w, 255u) << 24);
}

template <bool ZERO>
__global__ void __launch_bounds__(256)
k_clamp_counts(uint32_t* __restrict__ counts, uint64_t n, uint8_t* __restrict__ dens) {
    const uint64_t n16 = n >> 4;
    uint4* c4 = reinterpret_cast<uint4*>(counts);
    uint4* d4 = reinterpret_cast<uint4*>(dens);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = c4[4 * i], b = c4[4 * i + 1], c = c4[4 * i + 2], d = c4[4 * i + 3];
        d4[i] = make_uint4(clamp4(a), clamp4(b), clamp4(c), clamp4(d));
        if (ZERO) { c4[4 * i] = z; c4[4 * i + 1] = z; c4[4 * i + 2] = z; c4[4 * i + 3] = z; }
    }
    // tail (n % 16 voxels), first threads of block 0
    if (blockIdx.x == 0) {
        for (uint64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) {
            dens[i] = (uint8_t)min(counts[i], 255u);
            if (ZERO) counts[i] = 0u;
        }
    }
}

// Multi-GPU combine on saturated u8 partials: out[i] = min(sum_r slab_r[i], 255).  Every partial is already
// min(count_r, 255), and min(sum_r min(c_r, 255), 255) == min(sum_r c_r, 255), so the byte-wise saturating add
// (`__vaddus4`, four voxels per instruction) is exact -- and the partials are 4x smaller than u32 counts on the wire.
// slabs: n_slabs consecutive arrays of slab_bytes (a multiple of 16) each.
__global__ void __launch_bounds__(256)
k_saturating_sum_u8(const uint4* __restrict__ slabs, uint32_t n_slabs, uint64_t slab16, uint4* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slab16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 acc = __ldcs(slabs + i);
        for (uint32_t r = 1; r < n_slabs; ++r) {
            const uint4 v = __ldcs(slabs + (uint64_t)r * slab16 + i);
            acc.x = __vaddus4(acc.x, v.x); acc.y = __vaddus4(acc.y, v.y); acc.z = __vaddus4(acc.z, v.z); acc.w = __vaddus4(acc.w, v.w);
        }
        out[i] = acc;
    }
}

// The same combine as ONE kernel over NVLink peer memory: every GPU owns a slab of the volume, reads that slab of
// every peer's saturated u8 partial straight out of the peer's HBM (P2P loads), adds with saturation, and stores
// the finished slab into every peer's output volume (P2P stores) -- reduce-scatter, clamp and all-gather without
// an intermediate buffer or a second pass.  The caller brackets it with two device-side barriers of the
// symmetric-memory group (partials complete before, outputs complete after).
constexpr uint32_t kMaxPeers = 16;
struct PeerPtrs {
    const uint4* part[kMaxPeers];
    uint4* out[kMaxPeers];
    uint32_t n;
};
__global__ void __launch_bounds__(256)
k_combine_peer_u8(const __grid_constant__ PeerPtrs P, uint64_t off16, uint64_t slab16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slab16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v[kMaxPeers];
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; ++r)                 // all loads first: one NVLink round trip, not n
            if (r < P.n) v[r] = __ldcg(P.part[r] + off16 + i);
        uint4 acc = v[0];
#pragma unroll
        for (uint32_t r = 1; r < kMaxPeers; ++r)
            if (r < P.n) {
                acc.x = __vaddus4(acc.x, v[r].x); acc.y = __vaddus4(acc.y, v[r].y);
                acc.z = __vaddus4(acc.z, v[r].z); acc.w = __vaddus4(acc.w, v[r].w);
            }
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; ++r)
            if (r < P.n) __stcg(P.out[r] + off16 + i, acc);
    }
}

// Sparse form of the fused combine.  A hair volume is mostly zeros and a strand shard's partial even more so: each
// rank publishes one bit per 16-byte chunk of its partial (k_chunk_bitmap), the owner of a slab reads the peers'
// bitmap words (one broadcast load per warp and peer), fetches a chunk only from the peers that have something in it,
// and stores only non-zero results -- the outputs are zeroed by their owners before the first barrier.  NVLink then
// carries the hair, not the empty space.
__global__ void __launch_bounds__(256)
k_chunk_bitmap(const uint4* __restrict__ vol, uint64_t n16, uint32_t* __restrict__ bitmap) {
    // n16 is a multiple of 32 (the grid is padded to 16 * 32 * world bytes): whole warps, one word per warp-iteration
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(vol + i);
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (v.x | v.y | v.z | v.w) != 0u);
        if ((threadIdx.x & 31u) == 0u) bitmap[i >> 5] = bits;
    }
}
struct PeerPtrsSparse {
    const uint4* part[kMaxPeers];
    const uint32_t* bits[kMaxPeers];
    uint4* out[kMaxPeers];
    uint32_t n;
};
__global__ void __launch_bounds__(256)
k_combine_peer_u8_sparse(const __grid_constant__ PeerPtrsSparse P, uint64_t off16, uint64_t slab16) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slab16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = off16 + i;                            // chunk index in the whole grid; c >> 5 is warp-uniform
        uint32_t have = 0;                                       // peers with a non-zero chunk here
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; ++r)
            if (r < P.n) have |= ((__ldcg(P.bits[r] + (c >> 5)) >> lane) & 1u) << r;
        if (!have) continue;
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; ++r)
            if (r < P.n && ((have >> r) & 1u)) {
                const uint4 v = __ldcg(P.part[r] + c);
                acc.x = __vaddus4(acc.x, v.x); acc.y = __vaddus4(acc.y, v.y); acc.z = __vaddus4(acc.z, v.z); acc.w = __vaddus4(acc.w, v.w);
            }
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; ++r)
            if (r < P.n) __stcg(P.out[r] + c, acc);
    }
}

// Device-side barrier of the ranks of a strand-sharded voxelisation, over peer memory: every rank owns a signal pad
// (kMaxPeers words per slot) mapped into all ranks; rank r stores the call's epoch into word [slot][r] of EVERY pad and
// then waits until all words of its own pad have reached the epoch.  One CTA; the stream order of the launching rank puts
// everything it wrote before the barrier (partial volume, chunk bitmap, zeroed output) ahead of the signal
// (__threadfence_system), and the waiting rank's later kernels behind the acquire.  Epochs only grow: no reset, no ABA.
struct PeerSignals {
    uint32_t* pad[kMaxPeers];
    uint32_t n, rank;
};
__global__ void __launch_bounds__(32)
k_peer_barrier(const __grid_constant__ PeerSignals S, uint32_t slot, uint32_t epoch) {
    const uint32_t r = threadIdx.x;
    __threadfence_system();
    if (r < S.n) {
        uint32_t* remote = S.pad[r] + slot * kMaxPeers + S.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
        const uint32_t* mine = S.pad[S.rank] + slot * kMaxPeers + r;
        uint32_t v;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            __nanosleep(200);
            if (spin > (1u << 24)) __trap();                       // a rank that never arrives must fail, not hang the device
        }
    }
    __syncwarp();
    __threadfence_system();
}

// Same for an output grid that is not 16-byte aligned (a view into a caller's buffer): one voxel per thread.
template <bool ZERO>
__global__ void __launch_bounds__(256)
k_clamp_counts_unaligned(uint32_t* __restrict__ counts, uint64_t n, uint8_t* __restrict__ dens) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        dens[i] = (uint8_t)min(counts[i], 255u);
        if (ZERO) counts[i] = 0u;
    }
}

// ---------------------------------------------------------------------------
// Volume::normalize (hair_style.cc:344-357): min/max, then a 256-entry map
//   d -> (uchar)(float(uchar(d - lo)) * (255.0f / float(hi - lo))).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_minmax_u8(const uint8_t* __restrict__ dens, uint64_t n, uint32_t* __restrict__ lohi /* [0]=min, [1]=max */) {
    uint32_t lo = 255u, hi = 0u;
    const uint64_t n16 = n >> 4;
    const uint4* d4 = reinterpret_cast<const uint4*>(dens);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = d4[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // per-byte min/max of one word against the running values
            uint32_t x = w[k];
            uint32_t b0 = x & 0xFFu, b1 = (x >> 8) & 0xFFu, b2 = (x >> 16) & 0xFFu, b3 = x >> 24;
            lo = min(lo, min(min(b0, b1), min(b2, b3)));
            hi = max(hi, max(max(b0, b1), max(b2, b3)));
        }
    }
    if (blockIdx.x == 0)
        for (uint64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) {
            lo = min(lo, (uint32_t)dens[i]);
            hi = max(hi, (uint32_t)dens[i]);
        }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != 255u) atomicMin(lohi, lo);
        if (hi != 0u) atomicMax(lohi + 1, hi);
    }
}

__global__ void k_minmax_init(uint32_t* lohi) { lohi[0] = 255u; lohi[1] = 0u; }

__global__ void __launch_bounds__(256)
k_normalize_apply(uint8_t* __restrict__ dens, uint64_t n, const uint32_t* __restrict__ lohi) {
    __shared__ uint8_t lut[256];
    const uint32_t lo = lohi[0], hi = lohi[1];
    if (hi == lo) return;                                  // 255/0 in the reference: leave unchanged
    const float scaling = __fdiv_rn(255.0f, (float)(int)(hi - lo));
    for (uint32_t v = threadIdx.x; v < 256; v += blockDim.x) {
        const uint32_t d = (v - lo) & 0xFFu;                // unsigned char wrap of `d -= min`
        lut[v] = (uint8_t)__float2int_rz(__fmul_rn((float)d, scaling));
    }
    __syncthreads();
    const uint64_t n16 = n >> 4;
    uint4* d4 = reinterpret_cast<uint4*>(dens);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = d4[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = w[k];
            w[k] = (uint32_t)lut[x & 0xFFu] | ((uint32_t)lut[(x >> 8) & 0xFFu] << 8) |
                   ((uint32_t)lut[(x >> 16) & 0xFFu] << 16) | ((uint32_t)lut[x >> 24] << 24);
        }
        d4[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (blockIdx.x == 0)
        for (uint64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) dens[i] = lut[dens[i]];
}

// ---------------------------------------------------------------------------
// Volume::downsample (hair_style.hh:228-257): 2x2x2 -> 1, output (W/2,H/2,D/2).
// filter: 0 max, 1 sum/8, 2 (uchar)sum, 3 min.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_downsample(const uint8_t* __restrict__ in, uint32_t W, uint32_t H, uint32_t w, uint32_t h, uint32_t d,
             int filter, uint8_t* __restrict__ out) {
    const uint64_t n = (uint64_t)w * h * d;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n;
         o += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(o % w), j = (uint32_t)((o / w) % h), k = (uint32_t)(o / ((uint64_t)w * h));
        uint32_t s = 0, mx = 0, mn = 255;
#pragma unroll
        for (int z = 0; z < 2; ++z)
#pragma unroll
            for (int y = 0; y < 2; ++y) {
                const uint8_t* p = in + (2ull * i) + (uint64_t)(2 * j + y) * W + (uint64_t)(2 * k + z) * W * H;
                uint32_t a = p[0], b = p[1];
                s += a + b;
                mx = max(mx, max(a, b));
                mn = min(mn, min(a, b));
            }
        out[o] = (uint8_t)(filter == 0 ? mx : filter == 1 ? s / 8 : filter == 2 ? s : mn);
    }
}

// ---------------------------------------------------------------------------
// HairStyle::generate_indices (hair_style.cc:196-213) for strands of different lengths: segment j of the style
// belongs to the strand s with seg_prefix[s] <= j < seg_prefix[s+1] and joins vertices (j + s, j + s + 1) --
// every earlier strand contributes one vertex that starts no segment.  One thread per segment.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_generate_indices(const uint32_t* __restrict__ seg_prefix, uint32_t n_strands, uint32_t n_segments, uint2* __restrict__ pairs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_segments) return;
    uint32_t lo = 0, hi = n_strands;                        // last s with seg_prefix[s] <= j
    while (hi - lo > 1u) { const uint32_t m = (lo + hi) >> 1; if (__ldg(seg_prefix + m) <= j) lo = m; else hi = m; }
    pairs[j] = make_uint2(j + lo, j + lo + 1u);
}

// ---------------------------------------------------------------------------
// HairStyle::generate_bounding_box (hair_style.cc:215-234): min/max folded
// from (0,0,0).  Floats are mapped to order-preserving u32 keys so the fold is
// an integer atomicMin/Max; keys[0..2] = min keys, keys[3..5] = max keys,
// initialised to key(0.0f) by k_aabb_init and decoded by k_aabb_decode.
//
// Signed zeros.  The reference folds with glm::min(position, min) = (min < position) ? min : position and
// glm::max(position, max) = (position < max) ? max : position: on a TIE the new position wins, and -0.0f ties with
// +0.0f.  So when the minimum (or maximum) of an axis is zero, its SIGN is that of the last zero-valued coordinate in
// vertex order (+ when no coordinate is zero: the initial value).  keys[6..8] track that coordinate per axis:
// ((vertex + 1) << 1 | sign bit), folded with atomicMax.  Densities never see the difference ((v - -0.0f) == (v - 0.0f)
// up to the sign of a zero, which floor / the index sum discard), but the AABB the caller gets back is byte-identical.
// NaN coordinates: the reference's fold is order-dependent garbage there (a NaN replaces the running value, the next
// vertex replaces the NaN); defined here as ignored.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__global__ void k_aabb_init(uint32_t* keys) {
    if (threadIdx.x < 6) keys[threadIdx.x] = f2key(0.0f);
    else if (threadIdx.x < 9) keys[threadIdx.x] = 0u;
}
__global__ void __launch_bounds__(256)
k_aabb_reduce(const float* __restrict__ xyz, uint32_t n_vertices, uint32_t* __restrict__ keys) {
    // A block-iteration covers 768 consecutive floats (a multiple of 3), so the
    // float at base + 256*r + t is component (t + r) % 3: each thread keeps three
    // running (min,max) pairs, one per r, with fully coalesced loads.
    __shared__ uint32_t s_keys[9];
    const uint32_t z = f2key(0.0f);
    if (threadIdx.x < 6) s_keys[threadIdx.x] = z;
    else if (threadIdx.x < 9) s_keys[threadIdx.x] = 0u;
    __syncthreads();
    const uint64_t n = 3ull * n_vertices;
    uint32_t lo[3] = {z, z, z}, hi[3] = {z, z, z}, zero[3] = {0u, 0u, 0u};
    const uint64_t stride = (uint64_t)gridDim.x * 768ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 768ull; base < n; base += stride) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint64_t i = base + r * 256 + threadIdx.x;
            if (i < n) {
                const float f = xyz[i];
                if (f == 0.0f) zero[r] = max(zero[r], (((uint32_t)(i / 3ull) + 1u) << 1) | (__float_as_uint(f) >> 31));
                else if (f == f) {
                    const uint32_t k = f2key(f);
                    lo[r] = min(lo[r], k);
                    hi[r] = max(hi[r], k);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int comp = (threadIdx.x + r) % 3;
        if (lo[r] != z) atomicMin(&s_keys[comp], lo[r]);
        if (hi[r] != z) atomicMax(&s_keys[3 + comp], hi[r]);
        if (zero[r]) atomicMax(&s_keys[6 + comp], zero[r]);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(keys + threadIdx.x, s_keys[threadIdx.x]);
    else if (threadIdx.x < 9) atomicMax(keys + threadIdx.x, s_keys[threadIdx.x]);
}
__global__ void k_aabb_decode(const uint32_t* keys, float* out6) {
    if (threadIdx.x < 6) {
        float f = key2f(keys[threadIdx.x]);
        if (f == 0.0f) f = (keys[6 + threadIdx.x % 3] & 1u) ? -0.0f : 0.0f;    // the last zero coordinate of the axis decides the sign
        out6[threadIdx.x] = f;
    }
}

}  // namespace vkhr_b200
