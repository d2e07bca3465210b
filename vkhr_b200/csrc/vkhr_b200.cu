// vkhr_b200.cu -- the C ABI of include/vkhr_b200.h over the sm_100a kernels.
//
// Replaces, behind plain C entry points, the reference's
//   HairStyle::voxelize_segments   (src/vkhr/scene_graph/hair_style.cc:296-342)
//   HairStyle::voxelize_vertices   (src/vkhr/scene_graph/hair_style.cc:257-294)
//   Volume::normalize / downsample (hair_style.cc:344-357, hair_style.hh:228-257)
//   HairStyle::generate_bounding_box (hair_style.cc:215-234)
// There is no CPU fallback anywhere in this file: every entry point either
// runs the CUDA kernels or returns an error.
#include "../../include/vkhr_b200.h"
#include "kernels.cuh"
#include "prefilter.cuh"

#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cg = cooperative_groups;
using namespace vkhr_b200;

// ---------------------------------------------------------------------------
// PACKED8 / BRICK8 repair: one persistent cooperative kernel.  Instances whose overflow
// flag is set (some voxel received more than 255 hits) are handled one after
// the other on a single shared u32 scratch of kRepairChunk voxels (64 MiB; a
// whole 256^3 volume, 1/8 of a 512^3, 1/64 of a 1024^3 one -- the scratch used
// to be a full W*H*D u32 grid, 4 GiB at 1024^3, reserved eagerly for a pass that
// almost never runs).  Per chunk of the volume:
//   zero the scratch entries of flagged words -> grid barrier ->
//   re-walk the instance, counting only samples that land in flagged words of the chunk ->
//   grid barrier -> rewrite those words as min(count, 255).
// Chunks without a flagged word are skipped (one extra barrier to agree on that).
// With no flag set (the common case) every CTA returns after reading n flags.
// ---------------------------------------------------------------------------
constexpr uint32_t kRepairChunk = 1u << 24;                    // voxels; a multiple of 128 (whole bitmap words)

template <bool VERTICES>
__global__ void __launch_bounds__(kWalkThreads)
k_repair_packed(const __grid_constant__ Batch B, uint32_t* __restrict__ scratch) {
    // common case first: no instance overflowed -> one parallel look at the flags and out
    const uint32_t n_inst = B.n;
    const InstanceDev* inst = B.inst;
    // (BRICK8 has no per-word flags: the verdict sets the flag to 2 when the byte sum of the volume differs from
    // the number of samples added, and then every word is recounted)
    int any = 0;
    for (uint32_t k = threadIdx.x; k < n_inst; k += blockDim.x) any |= (*inst[k].ovf_flag != 0u);
    if (!__syncthreads_or(any)) return;                        // same answer in every CTA
    cg::grid_group grid = cg::this_grid();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t k = 0; k < n_inst; ++k) {
        const InstanceDev& I = inst[k];
        const uint32_t flag = *I.ovf_flag;                     // nothing below writes it: uniform across the grid
        if (flag == 0u) continue;
        const bool all = flag == 2u;
        const GridParams g = I.grid;
        uint32_t* words = reinterpret_cast<uint32_t*>(I.densities);
        uint4* counts4 = reinterpret_cast<uint4*>(scratch);
        volatile uint32_t* chunk_any = I.ovf_flag + 1;         // header word [1]: does this chunk hold a flagged word?
        for (uint32_t c0 = 0; c0 < g.n_voxels; c0 += kRepairChunk) {
            const uint32_t span = min(kRepairChunk, g.n_voxels - c0);          // voxels of this chunk
            const uint32_t w0 = c0 >> 2, n_words = span >> 2;                  // its 32-bit words (n_voxels % 16 == 0)
            const uint32_t b0 = w0 >> 5, n_bm = (n_words + 31u) >> 5;          // its bitmap words
            if (!all) {
                if (tid == 0) *chunk_any = 0u;
                grid.sync();
            }
            bool mine = false;
            for (uint32_t b = tid; b < n_bm; b += nthreads) {
                uint32_t m = all ? 0xFFFFFFFFu : I.ovf_bitmap[b0 + b];
                mine = mine || m != 0u;
                while (m) {
                    const uint32_t w = b * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    if (w < n_words) counts4[w] = make_uint4(0, 0, 0, 0);
                }
            }
            if (!all && mine) *chunk_any = 1u;
            grid.sync();
            if (!all && *chunk_any == 0u) continue;            // uniform: written before the barrier, not after
            SinkRecount sink{all ? nullptr : I.ovf_bitmap, scratch, c0, span};
            if (VERTICES) {
                for (uint32_t i = tid; i < I.n_vertices; i += nthreads) {
                    const float* v = I.vertices + 3ull * i;
                    uint32_t idx;
                    if (voxel_index(g, to_voxel_space(__ldg(v), g.ox, g.vsx, g.rvx), to_voxel_space(__ldg(v + 1), g.oy, g.vsy, g.rvy),
                                    to_voxel_space(__ldg(v + 2), g.oz, g.vsz, g.rvz), idx))
                        sink.put<0>(idx);
                }
            } else if (I.indices) {
                for (uint64_t s = tid; s < I.n_segments; s += nthreads) {
                    const uint2 pr = __ldg(reinterpret_cast<const uint2*>(I.indices) + s);
                    if (pr.x >= I.n_vertices || pr.y >= I.n_vertices) continue;      // as the walk: dropped, not read
                    const float* a = I.vertices + 3ull * pr.x;
                    const float* b = I.vertices + 3ull * pr.y;
                    walk_segment(g, __ldg(a), __ldg(a + 1), __ldg(a + 2), __ldg(b), __ldg(b + 1), __ldg(b + 2), sink);
                }
            } else {
                const uint32_t vps = I.segs_per_strand + 1u;
                for (uint32_t v = tid; v + 1u < I.n_vertices; v += nthreads) {
                    if (v % vps == vps - 1u) continue;             // last vertex of a strand starts no segment
                    const float* a = I.vertices + 3ull * v;
                    walk_segment(g, __ldg(a), __ldg(a + 1), __ldg(a + 2), __ldg(a + 3), __ldg(a + 4), __ldg(a + 5), sink);
                }
            }
            grid.sync();
            for (uint32_t b = tid; b < n_bm; b += nthreads) {
                uint32_t m = all ? 0xFFFFFFFFu : I.ovf_bitmap[b0 + b];
                while (m) {
                    const uint32_t w = b * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    if (w < n_words) words[w0 + w] = clamp4(counts4[w]);
                }
            }
            grid.sync();                                           // scratch is reused by the next chunk / flagged instance
        }
    }
}

// ---------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct vkhr_b200_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_per_sm = 0;       // shared memory of one SM (bytes): how many CTAs of the tiled prefilter fit
    cudaStream_t stream = nullptr;
    // every entry point shares the context's scratch: a call on another stream than the previous call's is ordered
    // behind it (pick())
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t order_event = nullptr;
    std::string err;
    uint64_t launches = 0;
    DevBuf counts;        // u32 scratch grid(s)
    size_t counts_clean_bytes = 0;   // leading bytes of `counts` known to be zero
    DevBuf bitmap;        // PACKED8 overflow bitmaps (+ flags at the front)
    uint32_t last_strategy = 0;   // VKHR_B200_STRATEGY_* of the last run_voxelize
    DevBuf brick;         // BRICK8 scratch volumes (brick order); all zero between calls up to brick_clean_bytes
    size_t brick_clean_bytes = 0;
    uint32_t shard_epoch = 0;     // barriers of the sharded entry point pair up by call count
    DevBuf frame_ctl;     // two FrameCtl blocks of the frame kernel (they alternate; each call zeroes the next call's)
    uint64_t frame_calls = 0;
    size_t ring_budget = size_t(64) << 20;        // bytes of BRICK8 scratch the frame kernel keeps in flight (L2-resident ring)
    DevBuf small;         // lohi[2] + aabb keys[6] + aabb floats[6]
    DevBuf tacc;          // tangent mode: 16-byte accumulator per (non-empty) voxel
    DevBuf tmeta;         // sparse tangent mode: voxel bitmap, per-word counts and slots, scan scratch
    size_t tacc_clean_bytes = 0;     // leading bytes of `tacc` known to be zero
    DevBuf st_vertices, st_indices, st_tangents, st_dens, st_tang_out;   // host-API staging
    // pipelined host crowd API: a ring of staging slots and two copy streams (one per PCIe direction)
    static constexpr int kSlots = 3;
    struct Slot { DevBuf vertices, indices, dens; cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr; };
    Slot slots[kSlots];
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    int repair_blocks[2] = {0, 0};
    int frame_slots[2] = {0, 0};  // CTAs of k_frame<3,3> / <4,4> the device holds at once (occupancy x SMs)
    DevBuf pf_occ;                // prefilter: occupancy cells + tile activity bytes
    DevBuf adsm_occ;              // ADSM coarse occupancy bits
    DevBuf adsm_table;            // the ADSM march's accumulated t sequence for `adsm_steps`
    float adsm_steps = 0.0f;
    uint32_t adsm_n = 0;
    uint32_t pf_smem_opted[kPfVariantCount] = {};   // dynamic shared memory each tiled prefilter instantiation has been opted into
    Batch batch;          // host copy of the kernel-parameter batch being launched
    // optional per-phase device timing (vkhr_b200_profile_*): CUDA events recorded on the
    // launching stream around each phase of run_voxelize
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int phase; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

enum Phase { PH_CLEAR = 0, PH_WALK = 1, PH_FINISH = 2, PH_NORMALIZE = 3, PH_PREFILTER = 4, PH_COUNT = 5 };

static thread_local std::string g_create_error;

namespace {

int fail(vkhr_b200_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define CU_CHECK(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? VKHR_B200_ERR_OUT_OF_MEMORY          \
                                                          : VKHR_B200_ERR_CUDA;                  \
            return fail(ctx, code_, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
        }                                                                                        \
    } while (0)

#define RET_IF(expr)                     \
    do {                                 \
        int rc_ = (expr);                \
        if (rc_ != VKHR_B200_OK) return rc_; \
    } while (0)

int bind(vkhr_b200_ctx* ctx) {
    if (!ctx) return fail(nullptr, VKHR_B200_ERR_INVALID_ARGUMENT, "null context");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    return VKHR_B200_OK;
}

int reserve(vkhr_b200_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return VKHR_B200_OK;
    if (b.p) {
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        CU_CHECK(ctx, cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { (void)cudaGetLastError(); want = bytes; e = cudaMalloc(&b.p, want); }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b.p = nullptr;
        return fail(ctx, VKHR_B200_ERR_OUT_OF_MEMORY, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    b.cap = want;
    return VKHR_B200_OK;
}

// The stream of this call.  The context's scratch (counters, flags, staging, the "known zero" invariants) is shared
// between calls, so a call that arrives on a different stream than the previous one waits for it: an event recorded on
// the old stream, awaited on the new one.  (A stream the caller has destroyed meanwhile has no pending work: the failing
// record is ignored.)
inline cudaStream_t pick(vkhr_b200_ctx* ctx, void* stream) {
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (ctx->last_stream_valid && s != ctx->last_stream && ctx->order_event) {
        if (cudaEventRecord(ctx->order_event, ctx->last_stream) == cudaSuccess) cudaStreamWaitEvent(s, ctx->order_event, 0);
        else (void)cudaGetLastError();
    }
    ctx->last_stream = s;
    ctx->last_stream_valid = true;
    return s;
}

cudaEvent_t take_event(vkhr_b200_ctx* ctx) {
    if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Scoped phase marker: records an event pair on `s` when profiling is on, else does nothing.
struct PhaseMark {
    vkhr_b200_ctx* ctx; cudaStream_t s; cudaEvent_t a = nullptr; int phase;
    PhaseMark(vkhr_b200_ctx* c, cudaStream_t st, int ph) : ctx(c), s(st), phase(ph) {
        if (ctx->profiling) { a = take_event(ctx); cudaEventRecord(a, s); }
    }
    ~PhaseMark() {
        if (a) { cudaEvent_t b = take_event(ctx); cudaEventRecord(b, s); ctx->spans.push_back({a, b, phase}); }
    }
};

inline unsigned stride_blocks(vkhr_b200_ctx* ctx, uint64_t items, unsigned per_block, unsigned waves) {
    uint64_t need = (items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)ctx->sm_count * waves;
    uint64_t g = need < cap ? need : cap;
    return (unsigned)(g ? g : 1);
}

int make_grid(vkhr_b200_ctx* ctx, const float origin[3], const float size[3],
              uint32_t W, uint32_t H, uint32_t D, uint32_t flags, GridParams& g) {
    if (!origin || !size) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null AABB");
    if (W == 0 || H == 0 || D == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "zero resolution");
    const unsigned long long n = (unsigned long long)W * H * D;
    if (n >= (1ull << 32)) return fail(ctx, VKHR_B200_ERR_UNSUPPORTED, "W*H*D must be < 2^32");
    for (int c = 0; c < 3; ++c)
        if (!std::isfinite(origin[c]) || !std::isfinite(size[c]) || !(size[c] > 0.0f))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "AABB must be finite with size > 0");
    // hair_style.cc:297-307: resolution is a glm::vec3 of floats, voxel_size = size / resolution
    const float rx = (float)(size_t)W, ry = (float)(size_t)H, rz = (float)(size_t)D;
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.vsx = size[0] / rx; g.vsy = size[1] / ry; g.vsz = size[2] / rz;
    g.rx1 = rx - 1.0f; g.ry1 = ry - 1.0f; g.rz1 = rz - 1.0f;
    g.Wf = rx; g.Hf = ry;
    g.W = W; g.H = H; g.D = D;
    g.n_voxels = (uint32_t)n;
    g.index_exact = (flags & VKHR_B200_INDEX_EXACT) ? 1u : 0u;
    // RN(1/voxel_size) for div_exact (walk.cuh); outside [2^-40, 2^40] the kernels use the plain IEEE division
    auto recip = [](float vs) { return (vs >= 9.094947e-13f && vs <= 1.0995116e12f) ? 1.0f / vs : 0.0f; };
    g.rvx = recip(g.vsx); g.rvy = recip(g.vsy); g.rvz = recip(g.vsz);
    g.fast_div = (g.rvx != 0.0f && g.rvy != 0.0f && g.rvz != 0.0f) ? 1u : 0u;
    // Small grids (<= 2^24 voxels): the fp32 index expression is exact for bounded positions, so the hot kernel
    // may compute it in int32 (walk.cuh, sample_index<2>).  The bound keeps |y*W| < 2^22, |z*W*H| < 2^30 and
    // leaves room for the rounding drift of the accumulated `root += dir`.
    g.small_grid = (n <= (1ull << 24) && !g.index_exact) ? 1u : 0u;
    g.pos_limit = 3.0e38f;
    if (g.small_grid) {
        const unsigned long long m = std::max(std::max(W, H), D);
        const unsigned long long lim = std::min<unsigned long long>(8192ull, std::min((1ull << 21) / m, (1ull << 29) / ((unsigned long long)W * H)));
        if (lim >= 8) g.pos_limit = (float)lim; else g.small_grid = 0u;
    }
    return VKHR_B200_OK;
}

// Number of segments described by (indices, n_indices) or by uniform strands.
int segment_count(vkhr_b200_ctx* ctx, const uint32_t* d_indices, uint64_t n_indices, uint32_t n_vertices,
                  uint32_t segs, uint64_t& n_segments) {
    if (n_vertices > 0xFFFFFFFFu / 3u - 1024u)                   // the kernels address floats with 32-bit offsets
        return fail(ctx, VKHR_B200_ERR_UNSUPPORTED, "more than 2^32 / 3 vertices");
    if (d_indices) {
        if (reinterpret_cast<uintptr_t>(d_indices) & 7u)
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "indices must be 8-byte aligned");
        n_segments = n_indices / 2;      // the reference loop `i < size()-1; i += 2` (hair_style.cc:311)
    } else {
        if (segs == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "indices == NULL needs segs_per_strand > 0");
        if (n_vertices % (segs + 1) != 0)
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "n_vertices is not a multiple of segs_per_strand + 1");
        n_segments = (uint64_t)(n_vertices / (segs + 1)) * segs;
    }
    return VKHR_B200_OK;
}

enum Strategy { COUNT32 = 0, PACKED8 = 1 };

struct Job {                 // one instance, host side
    const float* d_vertices;
    const uint32_t* d_indices;
    uint64_t n_segments;
    uint32_t n_vertices;
    uint32_t segs;
    GridParams grid;
    uint8_t* d_dens;
};

bool packed_ok(const Job& j) {
    return (j.grid.n_voxels % 16u) == 0 && (reinterpret_cast<uintptr_t>(j.d_dens) & 15u) == 0;
}

// Fill ctx->batch with `n` (<= kMaxBatch) jobs.
struct BatchPlan {
    uint32_t max_tiles[3] = {0, 0, 0};      // per WalkKind: grid.x of that kernel (0 = not needed)
};

BatchPlan fill_batch(vkhr_b200_ctx* ctx, const Job* jobs, uint32_t n, bool vertices_mode, bool frame = false) {
    BatchPlan plan;
    Batch& B = ctx->batch;
    B.n = n;
    for (uint32_t k = 0; k < n; ++k) {
        InstanceDev& I = B.inst[k];
        std::memset(&I, 0, sizeof I);
        const uint32_t kind = vertices_mode ? WK_SPLAT : (jobs[k].d_indices ? WK_INDEXED : WK_UNIFORM);
        I.vertices = jobs[k].d_vertices;
        I.indices = jobs[k].d_indices;
        I.n_segments = jobs[k].n_segments;
        I.n_vertices = jobs[k].n_vertices;
        I.segs_per_strand = jobs[k].segs;
        I.grid = jobs[k].grid;
        I.densities = jobs[k].d_dens;
        I.kind = kind;
        if (kind == WK_UNIFORM) {
            // warp-tiles of kTileStride vertices, kTilesPerWarp per warp, kWarpsPerBlock warps per CTA
            const uint64_t warp_tiles = ((uint64_t)jobs[k].n_vertices + kTileStride - 1) / kTileStride;
            const uint64_t per_cta = (uint64_t)kWarpsPerBlock * kTilesPerWarp;
            const uint64_t per_item = frame ? per_cta * kFrameRanges : per_cta;            // the frame kernel's items hold several ranges per warp
            I.n_tiles = (uint32_t)((warp_tiles + per_item - 1) / per_item);
        } else {
            const uint64_t items = (kind == WK_INDEXED) ? jobs[k].n_segments : jobs[k].n_vertices;
            const uint64_t per_item = frame ? kFrameIndexedSegs : kWalkThreads;     // the frame kernel's walk items are larger
            I.n_tiles = (uint32_t)((items + per_item - 1) / per_item);
        }
        I.vps_magic = (uint32_t)((1ull << 32) / (uint64_t)(jobs[k].segs + 1u)) + 1u;
        I.tile_step = kTileStride % (jobs[k].segs + 1u);
        plan.max_tiles[kind] = std::max(plan.max_tiles[kind], I.n_tiles);
    }
    return plan;
}

// Launch the walk (or splat) of instances [first, first + count) of ctx->batch.  MODE as in kernels.cuh.
template <int MODE>
int launch_walk(vkhr_b200_ctx* ctx, const BatchPlan& plan, bool exact, uint32_t first, uint32_t count, cudaStream_t s) {
    const Batch& B = ctx->batch;
    uint32_t tiles[3] = {0, 0, 0};
    for (uint32_t k = first; k < first + count; ++k) tiles[B.inst[k].kind] = std::max(tiles[B.inst[k].kind], B.inst[k].n_tiles);
    (void)plan;
    if (tiles[WK_UNIFORM]) {
        const dim3 grid(tiles[WK_UNIFORM], count);
        bool small = !exact;                                   // int32 index: every instance on a small grid
        for (uint32_t k = first; k < first + count; ++k) small = small && B.inst[k].grid.small_grid;
        if (exact)      k_walk_uniform<MODE, 1><<<grid, kWalkThreads, 0, s>>>(B, first);
        else if (small) k_walk_uniform<MODE, 2><<<grid, kWalkThreads, 0, s>>>(B, first);
        else            k_walk_uniform<MODE, 0><<<grid, kWalkThreads, 0, s>>>(B, first);
        ctx->launches++;
    }
    if (tiles[WK_INDEXED]) {
        const dim3 grid(tiles[WK_INDEXED], count);
        if (exact) k_walk_indexed<MODE, 1><<<grid, kWalkThreads, 0, s>>>(B, first);
        else       k_walk_indexed<MODE, 0><<<grid, kWalkThreads, 0, s>>>(B, first);
        ctx->launches++;
    }
    if (tiles[WK_SPLAT]) {
        const dim3 grid(tiles[WK_SPLAT], count);
        if (exact) k_splat_batch<MODE, 1><<<grid, kWalkThreads, 0, s>>>(B, first);
        else       k_splat_batch<MODE, 0><<<grid, kWalkThreads, 0, s>>>(B, first);
        ctx->launches++;
    }
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

// The BRICK8 walk of instances [first, first + count): segments only (uniform strands and / or index pairs), never in
// exact-index mode; small grids take the int32 index (<3, 3>), larger ones -- W, H powers of two, checked by the
// caller -- the reference's rounded fp32 index decomposed by bit fields (<4, 4>).
int launch_walk_brick(vkhr_b200_ctx* ctx, uint32_t first, uint32_t count, cudaStream_t s) {
    const Batch& B = ctx->batch;
    uint32_t tiles[3] = {0, 0, 0};
    bool small = true;
    for (uint32_t k = first; k < first + count; ++k) {
        tiles[B.inst[k].kind] = std::max(tiles[B.inst[k].kind], B.inst[k].n_tiles);
        small = small && B.inst[k].grid.small_grid;
    }
    if (tiles[WK_SPLAT]) return fail(ctx, VKHR_B200_ERR_UNSUPPORTED, "BRICK8 walks segments only");
    if (tiles[WK_UNIFORM]) {
        const dim3 grid(tiles[WK_UNIFORM], count);
        if (small) k_walk_uniform<3, 3><<<grid, kWalkThreads, 0, s>>>(B, first);
        else       k_walk_uniform<4, 4><<<grid, kWalkThreads, 0, s>>>(B, first);
        ctx->launches++;
    }
    if (tiles[WK_INDEXED]) {                                   // explicit index pairs: what the reference's own caller passes
        const dim3 grid(tiles[WK_INDEXED], count);
        if (small) k_walk_indexed<3, 3><<<grid, kWalkThreads, 0, s>>>(B, first);
        else       k_walk_indexed<4, 4><<<grid, kWalkThreads, 0, s>>>(B, first);
        ctx->launches++;
    }
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

template <bool VERTICES>
int launch_repair(vkhr_b200_ctx* ctx, uint32_t* scratch, cudaStream_t s) {
    int& blocks = ctx->repair_blocks[VERTICES ? 1 : 0];
    if (blocks == 0) {
        int per_sm = 0;
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_repair_packed<VERTICES>, kWalkThreads, 0));
        if (per_sm < 1) return fail(ctx, VKHR_B200_ERR_CUDA, "repair kernel does not fit on an SM");
        // one CTA slot per SM is left free: a cooperative grid only starts when ALL its CTAs fit at once, and a small kernel
        // of another stream that is waiting for this one's stream (the device-side barrier of a sharded call on the same
        // device) must not be able to keep it out for ever
        blocks = std::max(per_sm - 1, 1) * ctx->sm_count;
    }
    void* args[] = {(void*)&ctx->batch, (void*)&scratch};
    CU_CHECK(ctx, cudaLaunchCooperativeKernel((void*)k_repair_packed<VERTICES>, dim3(blocks), dim3(kWalkThreads), args, 0, s));
    ctx->launches++;
    return VKHR_B200_OK;
}

#ifndef VKHR_FRAME_COPIERS
#define VKHR_FRAME_COPIERS 256u
#endif
// BRICK8 as ONE launch per batch: the frame kernel (kernels.cuh, k_frame) walks and copies out through a ring of
// L2-resident scratch volumes; then the repair kernel looks at the verdict flags.  `small`: every grid <= 2^24 voxels.
int run_frame(vkhr_b200_ctx* ctx, const Job* jobs, uint32_t n, bool small, cudaStream_t s) {
    const uint64_t nv = jobs[0].grid.n_voxels;
    uint32_t ring = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ctx->ring_budget / nv, 1), 8);
    ring = std::min(ring, std::min(n, kMaxBatch));
    const size_t need = (size_t)ring * nv;
    if (need > ctx->brick.cap) ctx->brick_clean_bytes = 0;         // reserve() reallocates: contents undefined
    RET_IF(reserve(ctx, ctx->brick, need));
    if (ctx->brick_clean_bytes < need) {
        PhaseMark mk(ctx, s, PH_CLEAR);
        CU_CHECK(ctx, cudaMemsetAsync(ctx->brick.p, 0, need, s));
    }
    ctx->brick_clean_bytes = 0;                                    // until every launch below has been queued
    if (!ctx->frame_ctl.p) {
        RET_IF(reserve(ctx, ctx->frame_ctl, 2 * sizeof(FrameCtl)));
        CU_CHECK(ctx, cudaMemsetAsync(ctx->frame_ctl.p, 0, 2 * sizeof(FrameCtl), s));
        ctx->frame_calls = 0;
    }
    const size_t bm_words = 4;                                     // no per-word bitmap: the verdict flags whole instances
    constexpr size_t kHdr = 4;
    const uint32_t chunk = std::min<uint32_t>(n, kMaxBatch);
    RET_IF(reserve(ctx, ctx->bitmap, (size_t)chunk * (bm_words + kHdr) * 4));
    RET_IF(reserve(ctx, ctx->counts, std::min<uint64_t>(nv, kRepairChunk) * 4));   // the repair's chunk scratch
    ctx->counts_clean_bytes = 0;
    uint32_t* base = static_cast<uint32_t*>(ctx->bitmap.p);
    for (uint32_t first = 0; first < n; first += chunk) {
        const uint32_t m = std::min(chunk, n - first);
        fill_batch(ctx, jobs + first, m, false, true);
        FramePlan P{};
        P.ring = std::min(ring, m);
        // copiers: the last CTAs of instance i + 1 copy out instance i after their own walk (whose CTAs all started
        // before them: nothing to wait for); the batch's last instance is copied out by its own last CTAs
        if (P.ring < 2u && m > 1u) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "the frame kernel needs a ring of two scratch volumes");
        P.copiers = VKHR_FRAME_COPIERS;
        // The last instance's own copiers are the one wait on LARGER block indices in the kernel (each waits for every CTA
        // of its instance, the other copiers behind it included): all of them must fit on the device at once.  592 slots
        // on a B200; half of what this device holds, on a smaller part.
        int& slots = ctx->frame_slots[small ? 0 : 1];
        if (slots == 0) {
            int per_sm = 0;
            if (small) CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frame<3, 3>, kWalkThreads, 0));
            else       CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frame<4, 4>, kWalkThreads, 0));
            if (per_sm < 1) return fail(ctx, VKHR_B200_ERR_CUDA, "frame kernel does not fit on an SM");
            slots = per_sm * ctx->sm_count;
        }
        P.copiers_last = std::min<uint32_t>(256u, std::max<uint32_t>(1u, (uint32_t)slots / 2u));
        P.n_bricks = (uint32_t)(nv / 32);
        uint32_t max_items = 1;
        for (uint32_t p = 0; p < m; ++p) max_items = std::max(max_items, ctx->batch.inst[p].n_tiles);
        uint8_t* const ring_base = static_cast<uint8_t*>(ctx->brick.p);
        FrameCtl* ctl = static_cast<FrameCtl*>(ctx->frame_ctl.p);
        P.ctl = ctl + (ctx->frame_calls & 1u);
        P.ctl_next = ctl + ((ctx->frame_calls + 1u) & 1u);
        ctx->frame_calls++;
        for (uint32_t k = 0; k < m; ++k) {
            ctx->batch.inst[k].ovf_flag = base + (size_t)k * (bm_words + kHdr);
            ctx->batch.inst[k].ovf_bitmap = base + (size_t)k * (bm_words + kHdr) + kHdr;
            ctx->batch.inst[k].counts = static_cast<uint32_t*>(ctx->counts.p);
            ctx->batch.inst[k].brick = ring_base + (size_t)(k % P.ring) * nv;   // instance k counts in slot k mod ring
        }
        {
            PhaseMark mk(ctx, s, PH_WALK);
            const dim3 grid(max_items, m);
            if (small) k_frame<3, 3><<<grid, kWalkThreads, 0, s>>>(ctx->batch, P);
            else       k_frame<4, 4><<<grid, kWalkThreads, 0, s>>>(ctx->batch, P);
            // (the ring as a persisting access-policy window of this launch was measured: same time, same L2 misses,
            // profiles/r02_aj_*)
            ctx->launches++;
            CU_CHECK(ctx, cudaGetLastError());
        }
        PhaseMark mk(ctx, s, PH_FINISH);
        RET_IF(launch_repair<false>(ctx, static_cast<uint32_t*>(ctx->counts.p), s));
    }
    ctx->brick_clean_bytes = need;                                 // every copy-out re-zeroed what its walk touched
    return VKHR_B200_OK;
}

// The voxelisation of `n` instances at one resolution into their u8 grids, kMaxBatch at a time.
int run_voxelize(vkhr_b200_ctx* ctx, const Job* jobs, uint32_t n, bool vertices_mode, uint32_t flags, cudaStream_t s) {
    if (n == 0) return VKHR_B200_OK;
    const uint64_t nv = jobs[0].grid.n_voxels;
    const bool exact = (flags & VKHR_B200_INDEX_EXACT) != 0;
    bool packed = true;
    for (uint32_t k = 0; k < n; ++k) packed = packed && packed_ok(jobs[k]);
    if (flags & VKHR_B200_STRATEGY_COUNT32) packed = false;
    if ((flags & VKHR_B200_STRATEGY_PACKED8) && !packed)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "PACKED8 needs W*H*D % 16 == 0 and 16-byte aligned densities");

    // BRICK8: segments (uniform strands or index pairs), whole bricks, a small or power-of-two grid; otherwise plain PACKED8
    // (the default where it can run: 1.63 ms against 1.97 ms for the 64-instance crowd frame at 256^3 on B200)
    bool brick = packed && !(flags & VKHR_B200_STRATEGY_PACKED8) && !vertices_mode && !exact;
    uint64_t total_segments = 0;
    for (uint32_t k = 0; brick && k < n; ++k) {
        const GridParams& g = jobs[k].grid;
        const bool pow2 = (g.W & (g.W - 1u)) == 0 && (g.H & (g.H - 1u)) == 0;
        brick = (g.small_grid || pow2) && g.W % 4u == 0 && g.H % 4u == 0 && g.D % 2u == 0 &&
                g.W / 4u < 65536u && g.H / 4u < 65536u;
        total_segments += jobs[k].n_segments;
    }
    // the copy-out moves 2-3 bytes per voxel: it pays once there is about one segment per 100 voxels (measured on B200:
    // 0.1 segments per voxel, the crowd at 256^3: 1.97 -> 1.63 ms; 0.24, 32 M segments at 512^3: 0.87 -> 0.67 ms;
    // 0.024, 3.25 M at 512^3: 0.151 -> 0.129 ms; 0.0015, 1.6 M at 1024^3: 0.52 -> 0.59 ms, the clear pass is cheaper)
    if (brick && !(flags & VKHR_B200_STRATEGY_BRICK8) && total_segments * 100ull < (uint64_t)n * nv) brick = false;

    ctx->last_strategy = brick ? VKHR_B200_STRATEGY_BRICK8 : packed ? VKHR_B200_STRATEGY_PACKED8 : VKHR_B200_STRATEGY_COUNT32;
    // the frame kernel pays where at least two scratch volumes fit its L2-resident ring (256^3: 16 MiB of 64); one big
    // volume (512^3 and up) is copied out faster by the whole machine at once -- the separate kernels (measured on B200:
    // 3.25 M segments at 512^3 0.109 ms split, 0.141 ms frame; one ponytail at 256^3 0.042 ms split, 0.035 ms frame)
    if (brick && !(flags & VKHR_B200_BRICK8_SPLIT) && 2 * nv <= ctx->ring_budget) {
        bool small = true;
        for (uint32_t k = 0; k < n; ++k) small = small && jobs[k].grid.small_grid;
        const int rc = run_frame(ctx, jobs, n, small, s);
        if (rc == VKHR_B200_ERR_OUT_OF_MEMORY && !(flags & VKHR_B200_STRATEGY_BRICK8)) {
            brick = false;                                             // no room for the scratch ring: count in the output volumes
            ctx->last_strategy = VKHR_B200_STRATEGY_PACKED8;
        } else {
            RET_IF(rc);
            if (flags & VKHR_B200_NORMALIZE) {
                PhaseMark mk(ctx, s, PH_NORMALIZE);
                for (uint32_t k = 0; k < n; ++k) RET_IF(vkhr_b200_normalize_dev(ctx, jobs[k].d_dens, nv, s));
            }
            return VKHR_B200_OK;
        }
    }
    if (packed) {
        // scratch: per-instance overflow bitmap + flag, one shared u32 recount grid
        // per instance: a header (flag at word [0], the 2 x kStatSlots u64 statistics of BRICK8 from word [4]) + the bitmap
        const size_t bm_words = ((nv / 4 + 31) / 32 + 3) & ~size_t(3);
        constexpr size_t kHdr = 4 + 4 * (size_t)kStatSlots;
        uint32_t chunk = std::min<uint32_t>(n, kMaxBatch);
        if (brick) {
            // one scratch volume per instance of a chunk, within a 4 GiB budget (64 instances at 256^3 take 1 GiB; a
            // 1024^3 volume is 1 GiB by itself); if even that cannot be allocated the default falls back to counting in
            // the output volumes (PACKED8)
            const size_t budget = size_t(4) << 30;
            const uint32_t bchunk = (uint32_t)std::min<size_t>(chunk, std::max<size_t>(1, budget / nv));
            const size_t need = (size_t)bchunk * nv;
            if (need > ctx->brick.cap) ctx->brick_clean_bytes = 0;         // reserve() reallocates: contents undefined
            const int rc = reserve(ctx, ctx->brick, need);
            if (rc != VKHR_B200_OK) {
                if (flags & VKHR_B200_STRATEGY_BRICK8) return rc;
                brick = false;
                ctx->brick_clean_bytes = 0;
                ctx->last_strategy = VKHR_B200_STRATEGY_PACKED8;
            } else {
                chunk = bchunk;
                if (ctx->brick_clean_bytes < need) {
                    PhaseMark mk(ctx, s, PH_CLEAR);
                    CU_CHECK(ctx, cudaMemsetAsync(ctx->brick.p, 0, need, s));
                }
                ctx->brick_clean_bytes = 0;                                // until the copy-out of every chunk has been queued
            }
        }
        RET_IF(reserve(ctx, ctx->bitmap, (size_t)chunk * (bm_words + kHdr) * 4));
        RET_IF(reserve(ctx, ctx->counts, std::min<uint64_t>(nv, kRepairChunk) * 4));   // the repair's chunk scratch, not a whole u32 grid
        ctx->counts_clean_bytes = 0;                   // the recount may leave entries behind
        uint32_t* base = static_cast<uint32_t*>(ctx->bitmap.p);
        for (uint32_t first = 0; first < n; first += chunk) {
            const uint32_t m = std::min(chunk, n - first);
            const BatchPlan plan = fill_batch(ctx, jobs + first, m, vertices_mode);
            for (uint32_t k = 0; k < m; ++k) {
                ctx->batch.inst[k].ovf_flag = base + (size_t)k * (bm_words + kHdr);
                ctx->batch.inst[k].ovf_bitmap = base + (size_t)k * (bm_words + kHdr) + kHdr;
                ctx->batch.inst[k].stats = brick ? reinterpret_cast<unsigned long long*>(base + (size_t)k * (bm_words + kHdr) + 4) : nullptr;
                ctx->batch.inst[k].counts = static_cast<uint32_t*>(ctx->counts.p);
            }
            for (uint32_t k = 0; brick && k < m; ++k)
                ctx->batch.inst[k].brick = static_cast<uint8_t*>(ctx->brick.p) + (size_t)k * nv;
            const unsigned gx = brick ? 1u : stride_blocks(ctx, nv / 16, 256, m >= 8 ? 2 : 8);
            {
                PhaseMark mk(ctx, s, PH_CLEAR);
                k_clear_packed_batch<<<dim3(gx, m), 256, 0, s>>>(ctx->batch, 0u, brick ? 0u : 1u);
                ctx->launches++;
            }
            const bool any_work = plan.max_tiles[0] + plan.max_tiles[1] + plan.max_tiles[2] != 0;
            if (brick) {
                if (any_work) {
                    PhaseMark mk(ctx, s, PH_WALK);
                    RET_IF(launch_walk_brick(ctx, 0, m, s));
                }
                PhaseMark mk(ctx, s, PH_FINISH);                           // the copy-out also writes the zeros of an empty volume
                k_untile_batch<<<dim3(stride_blocks(ctx, nv / 32, 256, m >= 8 ? 2 : 8), m), 256, 0, s>>>(ctx->batch, 0u);
                k_brick_verdict<<<m, kStatSlots, 0, s>>>(ctx->batch, 0u);
                ctx->launches += 2;
            }
            if (any_work) {
                if (!brick) {
                    PhaseMark mk(ctx, s, PH_WALK);
                    RET_IF(launch_walk<1>(ctx, plan, exact, 0, m, s));
                }
                PhaseMark mk(ctx, s, PH_FINISH);
                if (vertices_mode) RET_IF(launch_repair<true>(ctx, static_cast<uint32_t*>(ctx->counts.p), s));
                else               RET_IF(launch_repair<false>(ctx, static_cast<uint32_t*>(ctx->counts.p), s));
            }
            CU_CHECK(ctx, cudaGetLastError());
        }
        if (brick) ctx->brick_clean_bytes = (size_t)chunk * nv;            // every copy-out re-zeroed what its walk touched
    } else {
        // COUNT32 in chunks of instances bounded by a 1 GiB scratch budget
        const size_t per = nv * 4;
        uint32_t chunk = (uint32_t)std::max<size_t>(1, (size_t(1) << 30) / per);
        chunk = std::min(std::min(chunk, n), kMaxBatch);
        const size_t need = per * chunk;
        if (need > ctx->counts.cap) ctx->counts_clean_bytes = 0;
        RET_IF(reserve(ctx, ctx->counts, need));
        for (uint32_t first = 0; first < n; first += chunk) {
            const uint32_t m = std::min(chunk, n - first);
            if (ctx->counts_clean_bytes < per * m) {
                PhaseMark mk(ctx, s, PH_CLEAR);
                CU_CHECK(ctx, cudaMemsetAsync(ctx->counts.p, 0, per * m, s));
                ctx->counts_clean_bytes = 0;           // until the ZERO clamp below has run
            }
            const BatchPlan plan = fill_batch(ctx, jobs + first, m, vertices_mode);
            for (uint32_t k = 0; k < m; ++k)
                ctx->batch.inst[k].counts = static_cast<uint32_t*>(ctx->counts.p) + (size_t)k * nv;
            {
                PhaseMark mk(ctx, s, PH_WALK);
                RET_IF(launch_walk<0>(ctx, plan, exact, 0, m, s));
            }
            PhaseMark mk(ctx, s, PH_FINISH);
            for (uint32_t k = 0; k < m; ++k) {
                uint32_t* c = static_cast<uint32_t*>(ctx->counts.p) + (size_t)k * nv;
                uint8_t* d = jobs[first + k].d_dens;
                if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0 && ((size_t)k * nv) % 4 == 0)
                    k_clamp_counts<true><<<stride_blocks(ctx, nv / 16 + 1, 256, 8), 256, 0, s>>>(c, nv, d);
                else
                    k_clamp_counts_unaligned<true><<<stride_blocks(ctx, nv, 256, 16), 256, 0, s>>>(c, nv, d);
                ctx->launches++;
            }
            CU_CHECK(ctx, cudaGetLastError());
            ctx->counts_clean_bytes = std::max(ctx->counts_clean_bytes, per * m);   // ZERO clamp restored the invariant
        }
    }
    if (flags & VKHR_B200_NORMALIZE) {
        PhaseMark mk(ctx, s, PH_NORMALIZE);
        for (uint32_t k = 0; k < n; ++k) RET_IF(vkhr_b200_normalize_dev(ctx, jobs[k].d_dens, nv, s));
    }
    return VKHR_B200_OK;
}

// ADD the hits of one shard into a caller-owned u32 grid (the multi-GPU partial).
int run_count(vkhr_b200_ctx* ctx, const Job& j, bool vertices_mode, uint32_t flags, uint32_t* d_counts, cudaStream_t s) {
    const BatchPlan plan = fill_batch(ctx, &j, 1, vertices_mode);
    if (plan.max_tiles[0] + plan.max_tiles[1] + plan.max_tiles[2] == 0) return VKHR_B200_OK;
    ctx->batch.inst[0].counts = d_counts;
    PhaseMark mk(ctx, s, PH_WALK);
    return launch_walk<0>(ctx, plan, (flags & VKHR_B200_INDEX_EXACT) != 0, 0, 1, s);
}

// Launch k_walk_tangent for one job.
int launch_walk_tangent(vkhr_b200_ctx* ctx, const Job& j, const float* d_tangents_in, bool vertices_mode, bool exact,
                        unsigned long long* acc, const uint32_t* bits, const uint32_t* prefix, cudaStream_t s) {
    const uint64_t items = vertices_mode ? j.n_vertices : j.n_segments;
    if (!items) return VKHR_B200_OK;
    const unsigned blocks = (unsigned)((items + kWalkThreads - 1) / kWalkThreads);
#define VKHR_LAUNCH_TANGENT(KIND)                                                                                      \
    do {                                                                                                               \
        if (exact) k_walk_tangent<KIND, 1><<<blocks, kWalkThreads, 0, s>>>(j.d_vertices, j.d_indices, d_tangents_in, items, j.n_vertices, j.segs, j.grid, acc, bits, prefix); \
        else       k_walk_tangent<KIND, 0><<<blocks, kWalkThreads, 0, s>>>(j.d_vertices, j.d_indices, d_tangents_in, items, j.n_vertices, j.segs, j.grid, acc, bits, prefix); \
    } while (0)
    if (vertices_mode) VKHR_LAUNCH_TANGENT(WK_SPLAT);
    else if (j.d_indices) VKHR_LAUNCH_TANGENT(WK_INDEXED);
    else VKHR_LAUNCH_TANGENT(WK_UNIFORM);
#undef VKHR_LAUNCH_TANGENT
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

// Densities AND the tangent volume of one instance (Volume::tangents).
// Segments on grids of whole 32-voxel words take the SPARSE form: the densities come from the ordinary (fast) density
// path; then the tangents are summed only where a voxel is non-empty -- a bitmap of the density volume, an exclusive
// scan of its per-word popcounts (cub::DeviceScan) and 16 bytes of accumulator per non-empty voxel, instead of 16
// bytes per voxel (16 GiB at 1024^3).  One device->host read of the non-empty count sizes the accumulator (this is the
// load-time path of the reference, rasterizer/hair_style.cc:75; the per-frame path does not ask for tangents).
// Everything else (vertices mode, odd grid sizes) counts and sums in a dense 16-byte-per-voxel accumulator.
int run_tangent(vkhr_b200_ctx* ctx, const Job& j, const float* d_tangents_in, bool vertices_mode, uint32_t flags,
                int8_t* d_tangents_out, cudaStream_t s) {
    if (reinterpret_cast<uintptr_t>(d_tangents_out) & 3u)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "tangents_out must be 4-byte aligned");
    if (vertices_mode && !d_tangents_in)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "voxelize_vertices needs tangents_in to produce tangents_out");
    const uint64_t nv = j.grid.n_voxels;
    const bool exact = (flags & VKHR_B200_INDEX_EXACT) != 0;
    const bool sparse = !vertices_mode && nv % 32u == 0 && (reinterpret_cast<uintptr_t>(j.d_dens) & 15u) == 0;
    if (sparse) {
        RET_IF(run_voxelize(ctx, &j, 1, false, flags & ~(uint32_t)VKHR_B200_NORMALIZE, s));
        const uint32_t n_words = (uint32_t)(nv / 32);
        size_t scan_bytes = 0;
        CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_words, s));
        const size_t words_bytes = ((size_t)n_words * 4 + 255) & ~size_t(255);
        RET_IF(reserve(ctx, ctx->tmeta, 3 * words_bytes + scan_bytes + 256));
        uint32_t* bits = static_cast<uint32_t*>(ctx->tmeta.p);
        uint32_t* counts = bits + words_bytes / 4;
        uint32_t* prefix = counts + words_bytes / 4;
        void* scan_tmp = prefix + words_bytes / 4;
        uint32_t last[2] = {0, 0};
        {
            PhaseMark mk(ctx, s, PH_CLEAR);
            k_tangent_bitmap<<<(n_words + 255) / 256, 256, 0, s>>>(j.d_dens, n_words, bits, counts);
            ctx->launches++;
            CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, prefix, (int)n_words, s));
            CU_CHECK(ctx, cudaMemcpyAsync(&last[0], prefix + n_words - 1, 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(ctx, cudaMemcpyAsync(&last[1], counts + n_words - 1, 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(ctx, cudaStreamSynchronize(s));
        }
        const size_t nnz = (size_t)last[0] + last[1];
        const size_t need = std::max<size_t>(nnz, 1) * 16;
        if (need > ctx->tacc.cap) ctx->tacc_clean_bytes = 0;
        RET_IF(reserve(ctx, ctx->tacc, need));
        unsigned long long* acc = static_cast<unsigned long long*>(ctx->tacc.p);
        CU_CHECK(ctx, cudaMemsetAsync(acc, 0, need, s));
        ctx->tacc_clean_bytes = 0;
        {
            PhaseMark mk(ctx, s, PH_WALK);
            RET_IF(launch_walk_tangent(ctx, j, d_tangents_in, false, exact, acc, bits, prefix, s));
        }
        PhaseMark mk(ctx, s, PH_FINISH);
        k_finish_tangent_sparse<<<stride_blocks(ctx, nv, 256, 16), 256, 0, s>>>(acc, bits, prefix, nv, reinterpret_cast<uint32_t*>(d_tangents_out));
        ctx->launches++;
        CU_CHECK(ctx, cudaGetLastError());
    } else {
        const size_t need = (size_t)nv * 16;
        if (need > ctx->tacc.cap) ctx->tacc_clean_bytes = 0;
        RET_IF(reserve(ctx, ctx->tacc, need));
        unsigned long long* acc = static_cast<unsigned long long*>(ctx->tacc.p);
        if (ctx->tacc_clean_bytes < need) {
            PhaseMark mk(ctx, s, PH_CLEAR);
            CU_CHECK(ctx, cudaMemsetAsync(acc, 0, need, s));
        }
        ctx->tacc_clean_bytes = 0;
        {
            PhaseMark mk(ctx, s, PH_WALK);
            RET_IF(launch_walk_tangent(ctx, j, d_tangents_in, vertices_mode, exact, acc, nullptr, nullptr, s));
        }
        {
            PhaseMark mk(ctx, s, PH_FINISH);
            k_finish_tangent<<<stride_blocks(ctx, nv, 256, 16), 256, 0, s>>>(acc, nv, j.d_dens, reinterpret_cast<uint32_t*>(d_tangents_out));
            ctx->launches++;
            CU_CHECK(ctx, cudaGetLastError());
        }
        ctx->tacc_clean_bytes = need;                   // the finish pass zeroed every entry it found non-empty
    }
    if (flags & VKHR_B200_NORMALIZE) {
        PhaseMark mk(ctx, s, PH_NORMALIZE);
        RET_IF(vkhr_b200_normalize_dev(ctx, j.d_dens, nv, s));
    }
    return VKHR_B200_OK;
}

int stage_in(vkhr_b200_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
    RET_IF(reserve(ctx, b, bytes ? bytes : 16));
    if (bytes) CU_CHECK(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return VKHR_B200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// exported functions
// ---------------------------------------------------------------------------
extern "C" {

const char* vkhr_b200_version(void) { return "vkhr_b200 0.1.0 (sm_100a)"; }

int vkhr_b200_create(int device, vkhr_b200_ctx** out) {
    if (!out) return fail(nullptr, VKHR_B200_ERR_INVALID_ARGUMENT, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, VKHR_B200_ERR_NO_DEVICE,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (this library has no CPU fallback)");
    }
    if (device < 0 || device >= count) return fail(nullptr, VKHR_B200_ERR_INVALID_ARGUMENT, "device index out of range");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(nullptr, VKHR_B200_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, VKHR_B200_ERR_NO_DEVICE,
                    std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                        "; this library carries sm_100a code only");
    vkhr_b200_ctx* ctx = new vkhr_b200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, VKHR_B200_ERR_CUDA, "cannot create a stream on the device");
    }
    if (cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); ctx->order_event = nullptr; }
    if (reserve(ctx, ctx->small, 256) != VKHR_B200_OK) {
        g_create_error = ctx->err;
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return VKHR_B200_ERR_OUT_OF_MEMORY;
    }
    // CUDA loads kernels lazily, at their first launch, and loading can synchronise the context -- behind a kernel of
    // another rank that is WAITING for this rank (the device-side barrier of a sharded call on a shared device), that is a
    // deadlock.  So the kernels a sharded call can launch are loaded here, while nothing is waiting for anything.
    {
        cudaFuncAttributes fa;
        const void* eager[] = {(const void*)k_peer_barrier, (const void*)k_chunk_bitmap, (const void*)k_combine_peer_u8_sparse,
                               (const void*)k_combine_peer_u8, (const void*)k_frame<3, 3>, (const void*)k_frame<4, 4>,
                               (const void*)k_walk_uniform<3, 3>, (const void*)k_walk_uniform<4, 4>, (const void*)k_walk_indexed<3, 3>,
                               (const void*)k_walk_indexed<4, 4>, (const void*)k_walk_uniform<1, 0>, (const void*)k_walk_uniform<1, 1>,
                               (const void*)k_walk_uniform<1, 2>, (const void*)k_walk_indexed<1, 0>, (const void*)k_walk_indexed<1, 1>,
                               (const void*)k_walk_uniform<0, 0>, (const void*)k_walk_uniform<0, 1>, (const void*)k_walk_uniform<0, 2>,
                               (const void*)k_walk_indexed<0, 0>, (const void*)k_walk_indexed<0, 1>, (const void*)k_clamp_counts<true>,
                               (const void*)k_clamp_counts_unaligned<true>, (const void*)k_untile_batch, (const void*)k_clear_packed_batch,
                               (const void*)k_brick_verdict, (const void*)k_repair_packed<false>, (const void*)k_minmax_init,
                               (const void*)k_minmax_u8, (const void*)k_normalize_apply};
        for (const void* k : eager)
            if (cudaFuncGetAttributes(&fa, k) != cudaSuccess) (void)cudaGetLastError();
    }
    *out = ctx;
    return VKHR_B200_OK;
}

void vkhr_b200_destroy(vkhr_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->counts, &ctx->bitmap, &ctx->brick, &ctx->frame_ctl, &ctx->small, &ctx->tacc, &ctx->tmeta, &ctx->adsm_table, &ctx->adsm_occ, &ctx->pf_occ, &ctx->st_vertices,
                      &ctx->st_indices, &ctx->st_tangents, &ctx->st_dens, &ctx->st_tang_out};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (auto& sl : ctx->slots) {
        DevBuf* sb[] = {&sl.vertices, &sl.indices, &sl.dens};
        for (DevBuf* b : sb) if (b->p) cudaFree(b->p);
        cudaEvent_t ev[] = {sl.uploaded, sl.computed, sl.downloaded};
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    }
    if (ctx->order_event) cudaEventDestroy(ctx->order_event);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* vkhr_b200_last_error(const vkhr_b200_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

void* vkhr_b200_stream(vkhr_b200_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

int vkhr_b200_synchronize(vkhr_b200_ctx* ctx) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

uint64_t vkhr_b200_launch_count(const vkhr_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint32_t vkhr_b200_last_strategy(const vkhr_b200_ctx* ctx) { return ctx ? ctx->last_strategy : 0u; }

int vkhr_b200_set_scratch_ring_bytes(vkhr_b200_ctx* ctx, size_t bytes) {
    RET_IF(bind(ctx));
    if (bytes == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "ring bytes must be > 0");
    ctx->ring_budget = bytes;
    return VKHR_B200_OK;
}

int vkhr_b200_profile_enable(vkhr_b200_ctx* ctx, int enable) {
    RET_IF(bind(ctx));
    ctx->profiling = enable != 0;
    return VKHR_B200_OK;
}

int vkhr_b200_profile_read_ex(vkhr_b200_ctx* ctx, double* ms_out, uint32_t* spans_out, uint32_t n_phases) {
    RET_IF(bind(ctx));
    if (!ms_out || !spans_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null output");
    for (uint32_t k = 0; k < n_phases; ++k) { ms_out[k] = 0.0; spans_out[k] = 0; }
    for (auto& sp : ctx->spans) {
        CU_CHECK(ctx, cudaEventSynchronize(sp.b));
        float ms = 0.0f;
        CU_CHECK(ctx, cudaEventElapsedTime(&ms, sp.a, sp.b));
        if ((uint32_t)sp.phase < n_phases) { ms_out[sp.phase] += ms; spans_out[sp.phase] += 1; }
        ctx->event_pool.push_back(sp.a);
        ctx->event_pool.push_back(sp.b);
    }
    ctx->spans.clear();
    return VKHR_B200_OK;
}

int vkhr_b200_profile_read(vkhr_b200_ctx* ctx, double ms_out[4], uint32_t spans_out[4]) {
    return vkhr_b200_profile_read_ex(ctx, ms_out, spans_out, 4);
}

// ---- device-pointer API -----------------------------------------------------
int vkhr_b200_voxelize_segments_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                    const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
                                    const float* d_tangents_in, const float aabb_origin[3], const float aabb_size[3],
                                    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                    uint8_t* d_densities_out, int8_t* d_tangents_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_vertices || !d_densities_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices or densities");
    Job j{};
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, j.grid));
    RET_IF(segment_count(ctx, d_indices, n_indices, n_vertices, segs_per_strand, j.n_segments));
    j.d_vertices = d_vertices; j.d_indices = d_indices; j.n_vertices = n_vertices;
    j.segs = segs_per_strand; j.d_dens = d_densities_out;
    if (d_tangents_out) return run_tangent(ctx, j, d_tangents_in, false, flags, d_tangents_out, pick(ctx, stream));
    return run_voxelize(ctx, &j, 1, false, flags, pick(ctx, stream));
}

int vkhr_b200_voxelize_vertices_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                    const float* d_tangents_in, const float aabb_origin[3], const float aabb_size[3],
                                    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                    uint8_t* d_densities_out, int8_t* d_tangents_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_vertices || !d_densities_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices or densities");
    Job j{};
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, j.grid));
    j.d_vertices = d_vertices; j.n_vertices = n_vertices; j.d_dens = d_densities_out;
    if (d_tangents_out) return run_tangent(ctx, j, d_tangents_in, true, flags, d_tangents_out, pick(ctx, stream));
    return run_voxelize(ctx, &j, 1, true, flags, pick(ctx, stream));
}

int vkhr_b200_voxelize_segments_batch_dev(vkhr_b200_ctx* ctx, const vkhr_b200_instance* instances, uint32_t n,
                                          uint32_t W, uint32_t H, uint32_t D, uint32_t flags, void* stream) {
    RET_IF(bind(ctx));
    if (n == 0) return VKHR_B200_OK;
    if (!instances) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null instance array");
    std::vector<Job> jobs(n);
    for (uint32_t k = 0; k < n; ++k) {
        const vkhr_b200_instance& in = instances[k];
        if ((!in.d_vertices && in.n_vertices) || !in.d_densities_out)      // (an instance without vertices gets an empty volume)
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "instance " + std::to_string(k) + ": null vertices or densities");
        Job& j = jobs[k];
        RET_IF(make_grid(ctx, in.aabb_origin, in.aabb_size, W, H, D, flags, j.grid));
        RET_IF(segment_count(ctx, in.d_indices, in.n_indices, in.n_vertices, in.segs_per_strand, j.n_segments));
        j.d_vertices = in.d_vertices; j.d_indices = in.d_indices; j.n_vertices = in.n_vertices;
        j.segs = in.segs_per_strand; j.d_dens = in.d_densities_out;
    }
    return run_voxelize(ctx, jobs.data(), n, false, flags, pick(ctx, stream));
}

// ---- multi-GPU building blocks -----------------------------------------------
int vkhr_b200_count_segments_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                 const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
                                 const float aabb_origin[3], const float aabb_size[3],
                                 uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                 uint32_t* d_counts_inout, void* stream) {
    RET_IF(bind(ctx));
    if (!d_counts_inout) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null counts");
    if (n_vertices == 0) return VKHR_B200_OK;
    if (!d_vertices) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices");
    Job j{};
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, j.grid));
    RET_IF(segment_count(ctx, d_indices, n_indices, n_vertices, segs_per_strand, j.n_segments));
    j.d_vertices = d_vertices; j.d_indices = d_indices; j.n_vertices = n_vertices; j.segs = segs_per_strand;
    return run_count(ctx, j, false, flags, d_counts_inout, pick(ctx, stream));
}

int vkhr_b200_count_vertices_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                 const float aabb_origin[3], const float aabb_size[3],
                                 uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                 uint32_t* d_counts_inout, void* stream) {
    RET_IF(bind(ctx));
    if (!d_counts_inout) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null counts");
    if (n_vertices == 0) return VKHR_B200_OK;
    if (!d_vertices) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices");
    Job j{};
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, j.grid));
    j.d_vertices = d_vertices; j.n_vertices = n_vertices;
    return run_count(ctx, j, true, flags, d_counts_inout, pick(ctx, stream));
}

int vkhr_b200_clamp_counts_dev(vkhr_b200_ctx* ctx, const uint32_t* d_counts, uint64_t n_voxels, uint32_t flags,
                               uint8_t* d_densities_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_counts || !d_densities_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    if ((reinterpret_cast<uintptr_t>(d_counts) & 15u) || (reinterpret_cast<uintptr_t>(d_densities_out) & 15u))
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "counts and densities must be 16-byte aligned");
    cudaStream_t s = pick(ctx, stream);
    k_clamp_counts<false><<<stride_blocks(ctx, n_voxels / 16 + 1, 256, 8), 256, 0, s>>>(
        const_cast<uint32_t*>(d_counts), n_voxels, d_densities_out);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    if (flags & VKHR_B200_NORMALIZE) RET_IF(vkhr_b200_normalize_dev(ctx, d_densities_out, n_voxels, s));
    return VKHR_B200_OK;
}

int vkhr_b200_saturating_sum_u8_dev(vkhr_b200_ctx* ctx, const uint8_t* d_slabs, uint32_t n_slabs, uint64_t slab_bytes,
                                    uint8_t* d_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_slabs || !d_out || n_slabs == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer or no slabs");
    if ((slab_bytes & 15u) || (reinterpret_cast<uintptr_t>(d_slabs) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "slabs and output must be 16-byte aligned, slab_bytes a multiple of 16");
    if (slab_bytes == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    k_saturating_sum_u8<<<stride_blocks(ctx, slab_bytes / 16, 256, 8), 256, 0, s>>>(
        reinterpret_cast<const uint4*>(d_slabs), n_slabs, slab_bytes / 16, reinterpret_cast<uint4*>(d_out));
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_combine_peer_u8_dev(vkhr_b200_ctx* ctx, const void* const* d_partials, void* const* d_outs, uint32_t n_peers,
                                  uint64_t slab_offset_bytes, uint64_t slab_bytes, void* stream) {
    RET_IF(bind(ctx));
    if (!d_partials || !d_outs || n_peers == 0 || n_peers > kMaxPeers)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "peer pointer arrays: 1.." + std::to_string(kMaxPeers) + " peers");
    if ((slab_offset_bytes & 15u) || (slab_bytes & 15u)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "slab offset and size must be multiples of 16");
    PeerPtrs P{};
    P.n = n_peers;
    for (uint32_t r = 0; r < n_peers; ++r) {
        if (!d_partials[r] || !d_outs[r] || (reinterpret_cast<uintptr_t>(d_partials[r]) & 15u) || (reinterpret_cast<uintptr_t>(d_outs[r]) & 15u))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "peer buffers must be non-null and 16-byte aligned");
        P.part[r] = static_cast<const uint4*>(d_partials[r]);
        P.out[r] = static_cast<uint4*>(d_outs[r]);
    }
    if (slab_bytes == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    k_combine_peer_u8<<<stride_blocks(ctx, slab_bytes / 16, 256, 8), 256, 0, s>>>(P, slab_offset_bytes / 16, slab_bytes / 16);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_chunk_bitmap_dev(vkhr_b200_ctx* ctx, const uint8_t* d_volume, uint64_t n_bytes, uint32_t* d_bitmap, void* stream) {
    RET_IF(bind(ctx));
    if (!d_volume || !d_bitmap) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    if ((n_bytes & 511u) || (reinterpret_cast<uintptr_t>(d_volume) & 15u))
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "volume must be 16-byte aligned and a multiple of 512 bytes (32 chunks per bitmap word)");
    if (n_bytes == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    k_chunk_bitmap<<<stride_blocks(ctx, n_bytes / 16, 256, 8), 256, 0, s>>>(reinterpret_cast<const uint4*>(d_volume), n_bytes / 16, d_bitmap);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_combine_peer_u8_sparse_dev(vkhr_b200_ctx* ctx, const void* const* d_partials, const void* const* d_bitmaps,
                                         void* const* d_outs, uint32_t n_peers, uint64_t slab_offset_bytes, uint64_t slab_bytes,
                                         void* stream) {
    RET_IF(bind(ctx));
    if (!d_partials || !d_bitmaps || !d_outs || n_peers == 0 || n_peers > kMaxPeers)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "peer pointer arrays: 1.." + std::to_string(kMaxPeers) + " peers");
    if ((slab_offset_bytes & 511u) || (slab_bytes & 511u))
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "slab offset and size must be multiples of 512 (whole bitmap words)");
    PeerPtrsSparse P{};
    P.n = n_peers;
    for (uint32_t r = 0; r < n_peers; ++r) {
        if (!d_partials[r] || !d_bitmaps[r] || !d_outs[r] || (reinterpret_cast<uintptr_t>(d_partials[r]) & 15u) ||
            (reinterpret_cast<uintptr_t>(d_outs[r]) & 15u) || (reinterpret_cast<uintptr_t>(d_bitmaps[r]) & 3u))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "peer buffers must be non-null and aligned");
        P.part[r] = static_cast<const uint4*>(d_partials[r]);
        P.bits[r] = static_cast<const uint32_t*>(d_bitmaps[r]);
        P.out[r] = static_cast<uint4*>(d_outs[r]);
    }
    if (slab_bytes == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    k_combine_peer_u8_sparse<<<stride_blocks(ctx, slab_bytes / 16, 256, 8), 256, 0, s>>>(P, slab_offset_bytes / 16, slab_bytes / 16);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

uint64_t vkhr_b200_sharded_volume_bytes(uint32_t W, uint32_t H, uint32_t D, uint32_t world) {
    const uint64_t nv = (uint64_t)W * H * D, q = 512ull * (world ? world : 1u);
    return (nv + q - 1) / q * q;
}

int vkhr_b200_voxelize_segments_sharded_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                            const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
                                            const float aabb_origin[3], const float aabb_size[3],
                                            uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                            const vkhr_b200_shard_peers* peers, void* stream) {
    RET_IF(bind(ctx));
    if (!peers || !peers->partials || !peers->bitmaps || !peers->outs || !peers->signals)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null peer description");
    const uint32_t world = peers->world, rank = peers->rank;
    if (world == 0 || world > kMaxPeers || rank >= world)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "1.." + std::to_string(kMaxPeers) + " ranks, rank < world");
    for (uint32_t r = 0; r < world; ++r)
        if (!peers->partials[r] || !peers->bitmaps[r] || !peers->outs[r] || !peers->signals[r] ||
            (reinterpret_cast<uintptr_t>(peers->partials[r]) & 15u) || (reinterpret_cast<uintptr_t>(peers->outs[r]) & 15u) ||
            (reinterpret_cast<uintptr_t>(peers->bitmaps[r]) & 3u) || (reinterpret_cast<uintptr_t>(peers->signals[r]) & 3u))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "peer buffers must be non-null and aligned (volumes 16 bytes, words 4)");
    GridParams g;
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, g));
    const uint64_t nv = g.n_voxels, nvp = vkhr_b200_sharded_volume_bytes(W, H, D, world), slab = nvp / world;
    cudaStream_t s = pick(ctx, stream);
    uint8_t* partial = static_cast<uint8_t*>(peers->partials[rank]);
    uint8_t* out = static_cast<uint8_t*>(peers->outs[rank]);
    // 1. this rank's shard -> its saturated u8 partial (the ordinary single-GPU voxelisation; the pad stays zero)
    const bool empty = n_vertices == 0 || (d_indices && n_indices < 2);
    if (empty) CU_CHECK(ctx, cudaMemsetAsync(partial, 0, nv, s));
    else RET_IF(vkhr_b200_voxelize_segments_dev(ctx, d_vertices, n_vertices, d_indices, n_indices, segs_per_strand, nullptr, aabb_origin,
                                                aabb_size, W, H, D, flags & ~(uint32_t)VKHR_B200_NORMALIZE, partial, nullptr, s));
    // 2. which 16-byte chunks of it hold anything; the output starts from zero (peers store only non-zero results)
    {
        PhaseMark mk(ctx, s, PH_CLEAR);
        RET_IF(vkhr_b200_chunk_bitmap_dev(ctx, partial, nvp, static_cast<uint32_t*>(peers->bitmaps[rank]), s));
        CU_CHECK(ctx, cudaMemsetAsync(out, 0, nvp, s));
    }
    // 3. barrier, combine my slab from all partials into all outputs, barrier
    PeerSignals S{};
    S.n = world; S.rank = rank;
    for (uint32_t r = 0; r < world; ++r) S.pad[r] = static_cast<uint32_t*>(peers->signals[r]);
    const uint32_t epoch = ++ctx->shard_epoch;
    {
        PhaseMark mk(ctx, s, PH_NORMALIZE);                        // (profile slot [3]: the first barrier = waiting for the slowest rank)
        k_peer_barrier<<<1, 32, 0, s>>>(S, 0u, epoch);
        ctx->launches++;
    }
    {
        PhaseMark mk(ctx, s, PH_PREFILTER);                        // (profile slot [4]: the fused combine + the second barrier)
        RET_IF(vkhr_b200_combine_peer_u8_sparse_dev(ctx, reinterpret_cast<const void* const*>(peers->partials),
                                                    reinterpret_cast<const void* const*>(peers->bitmaps), peers->outs, world, (uint64_t)rank * slab, slab, s));
        k_peer_barrier<<<1, 32, 0, s>>>(S, 1u, epoch);
        ctx->launches++;
    }
    CU_CHECK(ctx, cudaGetLastError());
    if (flags & VKHR_B200_NORMALIZE) RET_IF(vkhr_b200_normalize_dev(ctx, out, nv, s));
    return VKHR_B200_OK;
}

// ---- Volume operations ---------------------------------------------------------
int vkhr_b200_normalize_dev(vkhr_b200_ctx* ctx, uint8_t* d_densities, uint64_t n_voxels, void* stream) {
    RET_IF(bind(ctx));
    if (!d_densities) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null densities");
    if (reinterpret_cast<uintptr_t>(d_densities) & 15u)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "densities must be 16-byte aligned");
    if (n_voxels == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    uint32_t* lohi = static_cast<uint32_t*>(ctx->small.p);
    const unsigned g = stride_blocks(ctx, n_voxels / 16 + 1, 256, 8);
    k_minmax_init<<<1, 32, 0, s>>>(lohi);
    k_minmax_u8<<<g, 256, 0, s>>>(d_densities, n_voxels, lohi);
    k_normalize_apply<<<g, 256, 0, s>>>(d_densities, n_voxels, lohi);
    ctx->launches += 3;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_downsample_dev(vkhr_b200_ctx* ctx, const uint8_t* d_densities, uint32_t W, uint32_t H, uint32_t D,
                             int filter, uint8_t* d_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_densities || !d_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    if (filter < 0 || filter > 3) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "unknown downsample filter");
    const uint32_t w = W / 2, h = H / 2, d = D / 2;
    const uint64_t n = (uint64_t)w * h * d;
    if (n == 0) return VKHR_B200_OK;
    cudaStream_t s = pick(ctx, stream);
    k_downsample<<<stride_blocks(ctx, n, 256, 16), 256, 0, s>>>(d_densities, W, H, w, h, d, filter, d_out);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_generate_bounding_box_dev(vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
                                        float* d_aabb_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_aabb_out || (n_vertices && !d_vertices)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t s = pick(ctx, stream);
    uint32_t* keys = static_cast<uint32_t*>(ctx->small.p) + 32;          // 9 words: min keys, max keys, last-zero trackers
    k_aabb_init<<<1, 32, 0, s>>>(keys);
    ctx->launches++;
    if (n_vertices) {
        k_aabb_reduce<<<stride_blocks(ctx, 3ull * n_vertices, 768, 8), 256, 0, s>>>(d_vertices, n_vertices, keys);
        ctx->launches++;
    }
    k_aabb_decode<<<1, 32, 0, s>>>(keys, d_aabb_out);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

// ---- host-pointer API ------------------------------------------------------------
int vkhr_b200_voxelize_segments(vkhr_b200_ctx* ctx, const float* vertices, uint32_t n_vertices,
                                const uint32_t* indices, uint64_t n_indices, uint32_t segs_per_strand,
                                const float* tangents_in, const float aabb_origin[3], const float aabb_size[3],
                                uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                uint8_t* densities_out, int8_t* tangents_out) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities_out || (n_vertices && !vertices)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices or densities");
    GridParams g;
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, g));
    const size_t nv = g.n_voxels;
    RET_IF(reserve(ctx, ctx->st_dens, nv));
    if (tangents_out) RET_IF(reserve(ctx, ctx->st_tang_out, nv * 4));
    // fewer than two indices: the reference underflows size()-1; defined here as an empty volume
    const bool empty = n_vertices == 0 || (indices && n_indices < 2);
    if (empty) {
        CU_CHECK(ctx, cudaMemsetAsync(ctx->st_dens.p, 0, nv, ctx->stream));
        if (tangents_out) CU_CHECK(ctx, cudaMemsetAsync(ctx->st_tang_out.p, 0, nv * 4, ctx->stream));
    } else {
        RET_IF(stage_in(ctx, ctx->st_vertices, vertices, (size_t)n_vertices * 12));
        if (indices) RET_IF(stage_in(ctx, ctx->st_indices, indices, (size_t)(n_indices / 2) * 8));
        if (tangents_out && tangents_in) RET_IF(stage_in(ctx, ctx->st_tangents, tangents_in, (size_t)n_vertices * 12));
        RET_IF(vkhr_b200_voxelize_segments_dev(ctx, static_cast<const float*>(ctx->st_vertices.p), n_vertices,
                                               indices ? static_cast<const uint32_t*>(ctx->st_indices.p) : nullptr,
                                               n_indices, segs_per_strand,
                                               (tangents_out && tangents_in) ? static_cast<const float*>(ctx->st_tangents.p) : nullptr,
                                               aabb_origin, aabb_size, W, H, D, flags, static_cast<uint8_t*>(ctx->st_dens.p),
                                               tangents_out ? static_cast<int8_t*>(ctx->st_tang_out.p) : nullptr, ctx->stream));
    }
    CU_CHECK(ctx, cudaMemcpyAsync(densities_out, ctx->st_dens.p, nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (tangents_out) CU_CHECK(ctx, cudaMemcpyAsync(tangents_out, ctx->st_tang_out.p, nv * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

int vkhr_b200_voxelize_vertices(vkhr_b200_ctx* ctx, const float* vertices, uint32_t n_vertices,
                                const float* tangents_in, const float aabb_origin[3], const float aabb_size[3],
                                uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
                                uint8_t* densities_out, int8_t* tangents_out) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities_out || (n_vertices && !vertices)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null vertices or densities");
    if (tangents_out && !tangents_in)
        return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "voxelize_vertices needs tangents_in to produce tangents_out");
    GridParams g;
    RET_IF(make_grid(ctx, aabb_origin, aabb_size, W, H, D, flags, g));
    const size_t nv = g.n_voxels;
    RET_IF(reserve(ctx, ctx->st_dens, nv));
    if (tangents_out) RET_IF(reserve(ctx, ctx->st_tang_out, nv * 4));
    if (n_vertices == 0) {
        CU_CHECK(ctx, cudaMemsetAsync(ctx->st_dens.p, 0, nv, ctx->stream));
        if (tangents_out) CU_CHECK(ctx, cudaMemsetAsync(ctx->st_tang_out.p, 0, nv * 4, ctx->stream));
    } else {
        RET_IF(stage_in(ctx, ctx->st_vertices, vertices, (size_t)n_vertices * 12));
        if (tangents_out) RET_IF(stage_in(ctx, ctx->st_tangents, tangents_in, (size_t)n_vertices * 12));
        RET_IF(vkhr_b200_voxelize_vertices_dev(ctx, static_cast<const float*>(ctx->st_vertices.p), n_vertices,
                                               tangents_out ? static_cast<const float*>(ctx->st_tangents.p) : nullptr,
                                               aabb_origin, aabb_size, W, H, D, flags,
                                               static_cast<uint8_t*>(ctx->st_dens.p),
                                               tangents_out ? static_cast<int8_t*>(ctx->st_tang_out.p) : nullptr, ctx->stream));
    }
    CU_CHECK(ctx, cudaMemcpyAsync(densities_out, ctx->st_dens.p, nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (tangents_out) CU_CHECK(ctx, cudaMemcpyAsync(tangents_out, ctx->st_tang_out.p, nv * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

int vkhr_b200_normalize(vkhr_b200_ctx* ctx, uint8_t* densities, uint64_t n_voxels) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null densities");
    if (n_voxels == 0) return VKHR_B200_OK;
    RET_IF(stage_in(ctx, ctx->st_dens, densities, n_voxels));
    RET_IF(vkhr_b200_normalize_dev(ctx, static_cast<uint8_t*>(ctx->st_dens.p), n_voxels, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(densities, ctx->st_dens.p, n_voxels, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

int vkhr_b200_downsample(vkhr_b200_ctx* ctx, const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
                         int filter, uint8_t* out) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities || !out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    const size_t n_in = (size_t)W * H * D, n_out = (size_t)(W / 2) * (H / 2) * (D / 2);
    if (n_out == 0) return VKHR_B200_OK;
    RET_IF(stage_in(ctx, ctx->st_dens, densities, n_in));
    RET_IF(reserve(ctx, ctx->st_tang_out, n_out));
    RET_IF(vkhr_b200_downsample_dev(ctx, static_cast<const uint8_t*>(ctx->st_dens.p), W, H, D, filter,
                                    static_cast<uint8_t*>(ctx->st_tang_out.p), ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(out, ctx->st_tang_out.p, n_out, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

int vkhr_b200_generate_bounding_box(vkhr_b200_ctx* ctx, const float* vertices, uint32_t n_vertices, float aabb_out[6]) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!aabb_out || (n_vertices && !vertices)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null pointer");
    RET_IF(stage_in(ctx, ctx->st_vertices, vertices, (size_t)n_vertices * 12));
    float* d_out = reinterpret_cast<float*>(static_cast<uint32_t*>(ctx->small.p) + 16);
    RET_IF(vkhr_b200_generate_bounding_box_dev(ctx, static_cast<const float*>(ctx->st_vertices.p), n_vertices, d_out, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(aabb_out, d_out, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

int vkhr_b200_volume_save(const char* path, const uint8_t* densities, uint64_t n_voxels) {
    if (!path || (!densities && n_voxels)) return VKHR_B200_ERR_INVALID_ARGUMENT;
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return VKHR_B200_ERR_INVALID_ARGUMENT;
    const bool ok = n_voxels == 0 || std::fwrite(densities, 1, n_voxels, f) == n_voxels;
    return (std::fclose(f) == 0 && ok) ? VKHR_B200_OK : VKHR_B200_ERR_INVALID_ARGUMENT;
}

// ---- host-pointer crowd API: upload / kernels / download of consecutive instances overlap ----------
int vkhr_b200_voxelize_segments_batch(vkhr_b200_ctx* ctx, const vkhr_b200_host_instance* instances, uint32_t n,
                                      uint32_t W, uint32_t H, uint32_t D, uint32_t flags) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (n == 0) return VKHR_B200_OK;
    if (!instances) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null instance array");
    // validate everything and size the slots before anything is enqueued
    std::vector<Job> jobs(n);
    size_t max_v = 16, max_i = 16;
    for (uint32_t k = 0; k < n; ++k) {
        const vkhr_b200_host_instance& in = instances[k];
        if (!in.densities_out || (in.n_vertices && !in.vertices))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "instance " + std::to_string(k) + ": null vertices or densities");
        Job& j = jobs[k];
        RET_IF(make_grid(ctx, in.aabb_origin, in.aabb_size, W, H, D, flags, j.grid));
        j.n_vertices = in.n_vertices; j.segs = in.segs_per_strand; j.n_segments = 0;
        const bool empty = in.n_vertices == 0 || (in.indices && in.n_indices < 2);
        if (!empty) {
            if (in.indices) j.n_segments = in.n_indices / 2;
            else RET_IF(segment_count(ctx, nullptr, 0, in.n_vertices, in.segs_per_strand, j.n_segments));
        }
        max_v = std::max(max_v, (size_t)in.n_vertices * 12);
        if (in.indices) max_i = std::max(max_i, (size_t)(in.n_indices / 2) * 8);
    }
    const size_t nv = jobs[0].grid.n_voxels;
    if (!ctx->h2d_stream) CU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) CU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    const uint32_t n_slots = std::min<uint32_t>(n, vkhr_b200_ctx::kSlots);
    for (uint32_t s = 0; s < n_slots; ++s) {
        auto& sl = ctx->slots[s];
        RET_IF(reserve(ctx, sl.vertices, max_v));
        RET_IF(reserve(ctx, sl.indices, max_i));
        RET_IF(reserve(ctx, sl.dens, nv));
        if (!sl.uploaded) {
            CU_CHECK(ctx, cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming));
            CU_CHECK(ctx, cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming));
            CU_CHECK(ctx, cudaEventCreateWithFlags(&sl.downloaded, cudaEventDisableTiming));
        }
    }
    cudaStream_t up = ctx->h2d_stream, run = ctx->stream, down = ctx->d2h_stream;
    for (uint32_t k = 0; k < n; ++k) {
        const vkhr_b200_host_instance& in = instances[k];
        auto& sl = ctx->slots[k % n_slots];
        Job& j = jobs[k];
        const bool reuse = k >= n_slots;
        // upload: the slot's input buffers are free once the kernels of its previous tenant have run
        if (reuse) CU_CHECK(ctx, cudaStreamWaitEvent(up, sl.computed, 0));
        if (j.n_segments) {
            CU_CHECK(ctx, cudaMemcpyAsync(sl.vertices.p, in.vertices, (size_t)in.n_vertices * 12, cudaMemcpyHostToDevice, up));
            if (in.indices) CU_CHECK(ctx, cudaMemcpyAsync(sl.indices.p, in.indices, (size_t)j.n_segments * 8, cudaMemcpyHostToDevice, up));
        }
        CU_CHECK(ctx, cudaEventRecord(sl.uploaded, up));
        // kernels: need the upload, and the slot's volume buffer back from the previous download
        CU_CHECK(ctx, cudaStreamWaitEvent(run, sl.uploaded, 0));
        if (reuse) CU_CHECK(ctx, cudaStreamWaitEvent(run, sl.downloaded, 0));
        j.d_vertices = static_cast<const float*>(sl.vertices.p);
        j.d_indices = in.indices ? static_cast<const uint32_t*>(sl.indices.p) : nullptr;
        j.d_dens = static_cast<uint8_t*>(sl.dens.p);
        if (j.n_segments) RET_IF(run_voxelize(ctx, &j, 1, false, flags, run));
        else CU_CHECK(ctx, cudaMemsetAsync(sl.dens.p, 0, nv, run));
        CU_CHECK(ctx, cudaEventRecord(sl.computed, run));
        // download
        CU_CHECK(ctx, cudaStreamWaitEvent(down, sl.computed, 0));
        CU_CHECK(ctx, cudaMemcpyAsync(in.densities_out, sl.dens.p, nv, cudaMemcpyDeviceToHost, down));
        CU_CHECK(ctx, cudaEventRecord(sl.downloaded, down));
    }
    CU_CHECK(ctx, cudaStreamSynchronize(down));
    CU_CHECK(ctx, cudaStreamSynchronize(run));
    CU_CHECK(ctx, cudaStreamSynchronize(up));
    return VKHR_B200_OK;
}

int vkhr_b200_host_register(vkhr_b200_ctx* ctx, void* ptr, size_t bytes) {
    RET_IF(bind(ctx));
    if (!ptr || !bytes) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null range");
    CU_CHECK(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return VKHR_B200_OK;
}
int vkhr_b200_host_unregister(vkhr_b200_ctx* ctx, void* ptr) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaHostUnregister(ptr));
    return VKHR_B200_OK;
}

// ---- density -> AO / opacity / Gaussian prefilter ------------------------------------------------------------
void vkhr_b200_prefilter_defaults(vkhr_b200_prefilter_params* p) {
    if (!p) return;
    p->ao_radius = 2.5f; p->ao_exponent = 10.0f; p->ao_max = 0.16f;      // interface.hh:101-105
    p->strand_alpha = 0.3f; p->thickness = 11.0f;                         // volume.frag:78
    p->gauss_width = 3.0f; p->flags = 0;
}

namespace {
// texels and weights of a LINEAR sample displaced by +-r voxels from a voxel centre (prefilter_oracle.c of the test infrastructure)
AxisTaps axis_taps(float r, bool positive) {
    AxisTaps a;
    const float fl = std::floor(r), fp = r - fl;
    const int ifl = (int)fl;
    if (positive) { a.o0 = ifl; a.o1 = ifl + 1; a.w0 = 1.0f - fp; a.w1 = fp; }
    else if (fp != 0.0f) { a.o0 = -ifl - 1; a.o1 = -ifl; a.w0 = fp; a.w1 = 1.0f - fp; }
    else { a.o0 = -ifl; a.o1 = -ifl + 1; a.w0 = 1.0f; a.w1 = 0.0f; }
    return a;
}

// instantiations of the tiled kernel: [0] row-wise (tap offsets at run time), then the column form for the tap offsets
// (neg.o0, pos.o0) of a non-integer radius with floor f: (-f-1, f), and of an integer radius r: (-r, r), up to radius 4
// (the default is 2.5; wider windows need more than 200 registers per thread and stay row-wise), each with tiles 8 and
// 16 voxels deep.  They are compiled in prefilter_variants.cu, one tap-offset pair per object file.
}  // namespace
extern "C++" {                                       // (this part of the file sits inside the extern "C" block of the ABI)
namespace vkhr_b200 {
int pf_variants_0(PfVariant*); int pf_variants_1(PfVariant*); int pf_variants_2(PfVariant*);
int pf_variants_3(PfVariant*); int pf_variants_4(PfVariant*); int pf_variants_5(PfVariant*);
int pf_variants_6(PfVariant*); int pf_variants_7(PfVariant*); int pf_variants_8(PfVariant*);
}
}  // extern "C++"
namespace {
struct PfTable {
    PfVariant v[kPfVariantCount];
    int n = 0;
    PfTable() {
        int (*groups[])(PfVariant*) = {pf_variants_0, pf_variants_1, pf_variants_2, pf_variants_3, pf_variants_4,
                                       pf_variants_5, pf_variants_6, pf_variants_7, pf_variants_8};
        for (auto g : groups) n += g(v + n);                 // 3 + 8 x 2 = kPfVariantCount entries, [0] = row-wise
    }
};
const PfTable& pf_table() { static const PfTable t; return t; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else (void)cudaGetLastError();
    }
    return fn;
}
}  // namespace

int vkhr_b200_prefilter_dev(vkhr_b200_ctx* ctx, const uint8_t* d_densities, uint32_t W, uint32_t H, uint32_t D,
                            const vkhr_b200_prefilter_params* params, float* d_ao, float* d_opacity, float* d_gauss, void* stream) {
    RET_IF(bind(ctx));
    if (!d_densities) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null densities");
    if (W == 0 || H == 0 || D == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "zero resolution");
    if ((unsigned long long)W * H * D >= (1ull << 32)) return fail(ctx, VKHR_B200_ERR_UNSUPPORTED, "W*H*D must be < 2^32");
    if (!d_ao && !d_opacity && !d_gauss) return VKHR_B200_OK;
    vkhr_b200_prefilter_params P;
    if (params) P = *params; else vkhr_b200_prefilter_defaults(&P);
    PrefilterArgs A{};
    A.dens = d_densities; A.W = (int)W; A.H = (int)H; A.D = (int)D;
    A.ao = d_ao; A.opacity = d_opacity; A.gauss = d_gauss;
    int halo = 0;
    if (d_ao) {
        if (!(P.ao_radius >= 0.0f) || !(P.ao_radius <= 64.0f) || !std::isfinite(P.ao_exponent) || !std::isfinite(P.ao_max))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "ao_radius must be in [0, 64] voxels and ao_exponent / ao_max finite");
        // local_ambient_occlusion.glsl:15-18 with kernel_size = 2: the taps sit kernel_radius * voxel_scaling voxels away
        const float kernel_radius = (2.0f - 1.0f) / 2.0f, voxel_scaling = P.ao_radius / kernel_radius;
        const float r = kernel_radius * voxel_scaling;
        A.neg = axis_taps(r, false); A.pos = axis_taps(r, true);
        A.ao_max = P.ao_max; A.ao_exponent = P.ao_exponent;
        halo = std::max(halo, std::max(A.pos.o1, -A.neg.o0));
    }
    if (d_opacity) {
        if (!std::isfinite(P.strand_alpha) || !std::isfinite(P.thickness))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "strand_alpha and thickness must be finite");
        A.one_minus_alpha = 1.0f - P.strand_alpha; A.thickness = P.thickness;
    }
    if (d_gauss) {
        const int n = (int)P.gauss_width;
        if ((float)n != P.gauss_width || n < 1 || n > kPfMaxGauss || (n & 1) == 0)
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "gauss_width must be an odd integer in [1, 9]");
        // sample_volume.glsl:17-19
        const float sigma_stddev = (P.gauss_width / 2.0f) / 2.4f;
        A.g_range = (n - 1) / 2; A.g_sigma2 = sigma_stddev * sigma_stddev;
        halo = std::max(halo, A.g_range);
    }
    A.halo = halo;
    A.tiles_x = (W + kPfTX - 1) / kPfTX; A.tiles_y = (H + kPfTY - 1) / kPfTY;
    cudaStream_t s = pick(ctx, stream);
    PhaseMark mk(ctx, s, PH_PREFILTER);
    const bool tiled = !(P.flags & VKHR_B200_PREFILTER_GENERIC) && halo <= kPfMaxHalo && (W % 16u) == 0 &&
                       (reinterpret_cast<uintptr_t>(d_densities) & 15u) == 0;
    if (!tiled) {
        const uint64_t n = (uint64_t)W * H * D;
        k_prefilter_generic<<<stride_blocks(ctx, n, 256, 32), 256, 0, s>>>(A);
        ctx->launches++;
        CU_CHECK(ctx, cudaGetLastError());
        return VKHR_B200_OK;
    }
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return fail(ctx, VKHR_B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // AO with tap offsets the column kernel is instantiated for (every radius up to 4 voxels) takes that instantiation --
    // with the deep tile when two CTAs of it fit an SM -- everything else the row-wise one
    int variant = 0;
    if (d_ao && !(P.flags & VKHR_B200_PREFILTER_ROWWISE) && A.neg.o1 == A.neg.o0 + 1 && A.pos.o1 == A.pos.o0 + 1) {
        const bool deep = 2u * (pf_plan(halo, A.g_range, kPfTZDeep).total + 1024u) <= (uint32_t)ctx->smem_per_sm && D > (uint32_t)kPfTZ;
        for (int v = 1; v < pf_table().n; ++v)
            if (pf_table().v[v].no0 == A.neg.o0 && pf_table().v[v].po0 == A.pos.o0 && pf_table().v[v].tz == (deep ? kPfTZDeep : kPfTZ))
                variant = v;
    }
    const PfKernel kernel = pf_table().v[variant].kernel;
    const int tz = pf_table().v[variant].tz;
    A.tiles_z = (D + tz - 1) / tz;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {W, H, D};
    const cuuint64_t gstride[2] = {(cuuint64_t)W, (cuuint64_t)W * H};                 // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)kPfBX, (cuuint32_t)(kPfTY + 2 * halo), (cuuint32_t)(tz + 2 * halo)};
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_densities), gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ctx, VKHR_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)cr));
    const PfSmemPlan plan = pf_plan(halo, A.g_range, tz);
    if (plan.total > ctx->pf_smem_opted[variant]) {
        CU_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total));
        ctx->pf_smem_opted[variant] = plan.total;
    }
    int per_sm = 0;
    CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPfThreads, plan.total));
    if (per_sm < 1) return fail(ctx, VKHR_B200_ERR_CUDA, "prefilter kernel does not fit on an SM");
    const uint64_t n_tiles = (uint64_t)A.tiles_x * A.tiles_y * A.tiles_z;
    // occupancy pre-pass: which tiles (with their halo) hold any hair at all -- the others are never loaded
    if (!(P.flags & VKHR_B200_PREFILTER_DENSE)) {
        const int cx = (int)((W + kPfCellX - 1) / kPfCellX), cy = (int)((H + kPfCellY - 1) / kPfCellY), cz = (int)((D + kPfCellZ - 1) / kPfCellZ);
        const size_t n_cells = (size_t)cx * cy * cz, occ_bytes = (n_cells + 255) & ~size_t(255);
        const size_t tiles_bytes = (n_tiles + 255) & ~size_t(255);
        RET_IF(reserve(ctx, ctx->pf_occ, occ_bytes + tiles_bytes + 256));
        uint8_t* occ = static_cast<uint8_t*>(ctx->pf_occ.p);
        uint8_t* active = occ + occ_bytes;
        uint32_t* counter = reinterpret_cast<uint32_t*>(active + tiles_bytes);
        k_pf_cell_occupancy<<<stride_blocks(ctx, n_cells, 8, 16), 256, 0, s>>>(d_densities, (int)W, (int)H, (int)D, cx, cy, cz, occ);
        k_pf_tile_active<<<(unsigned)((n_tiles + 255) / 256), 256, 0, s>>>(occ, cx, cy, cz, A.tiles_x, A.tiles_y, A.tiles_z, tz, active, counter);
        ctx->launches += 2;
        A.tile_active = active;
        A.tile_counter = counter;
    }
    const unsigned blocks = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)per_sm * ctx->sm_count);
    kernel<<<blocks, kPfThreads, plan.total, s>>>(tmap, A);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_prefilter(vkhr_b200_ctx* ctx, const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
                        const vkhr_b200_prefilter_params* params, float* ao_out, float* opacity_out, float* gauss_out) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null densities");
    const size_t n = (size_t)W * H * D;
    if (n == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "zero resolution");
    float* outs[3] = {ao_out, opacity_out, gauss_out};
    int wanted = 0;
    for (float* o : outs) wanted += o != nullptr;
    if (!wanted) return VKHR_B200_OK;
    RET_IF(stage_in(ctx, ctx->st_dens, densities, n));
    RET_IF(reserve(ctx, ctx->st_tang_out, n * 4 * wanted));
    float* d_out[3] = {nullptr, nullptr, nullptr};
    int slot = 0;
    for (int k = 0; k < 3; ++k) if (outs[k]) d_out[k] = static_cast<float*>(ctx->st_tang_out.p) + n * (slot++);
    RET_IF(vkhr_b200_prefilter_dev(ctx, static_cast<const uint8_t*>(ctx->st_dens.p), W, H, D, params, d_out[0], d_out[1], d_out[2], ctx->stream));
    for (int k = 0; k < 3; ++k)
        if (outs[k]) CU_CHECK(ctx, cudaMemcpyAsync(outs[k], d_out[k], n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

// ---- .hair bytes -> volume (HairStyle::load + SceneGraph::add_style + voxelize_segments) -------------------
// File layout: SURVEY Appendix C / include/vkhr/scene_graph/hair_style.hh:147-174, src/.../hair_style.cc:24-47.
int vkhr_b200_voxelize_hair(vkhr_b200_ctx* ctx, const void* hair_bytes, size_t n_bytes, uint32_t W, uint32_t H, uint32_t D,
                            uint32_t flags, uint8_t* densities_out, int8_t* tangents_out, float aabb_out[6]) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!hair_bytes || !densities_out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null file bytes or densities");
    const unsigned char* f = static_cast<const unsigned char*>(hair_bytes);
    if (n_bytes < 128 || std::memcmp(f, "HAIR", 4) != 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "not a .hair file (signature)");
    uint32_t strands, nverts, bits, dsegs;
    float bbox[6];
    std::memcpy(&strands, f + 4, 4); std::memcpy(&nverts, f + 8, 4); std::memcpy(&bits, f + 12, 4); std::memcpy(&dsegs, f + 16, 4);
    std::memcpy(bbox, f + 104, 24);                            // bounding_box_min[3], bounding_box_max[3]
    if (!(bits & 2u) || nverts == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, ".hair file without vertices");
    // array offsets, in file order: segments, vertices, thickness, transparency, color, tangents, indices
    size_t off = 128, o_seg = 0, o_vert = 0, o_tan = 0, o_idx = 0;
    if (bits & 1u) { o_seg = off; off += (size_t)strands * 2; }
    o_vert = off; off += (size_t)nverts * 12;
    if (bits & 4u) off += (size_t)nverts * 4;
    if (bits & 8u) off += (size_t)nverts * 4;
    if (bits & 16u) off += (size_t)nverts * 12;
    if (bits & 32u) { o_tan = off; off += (size_t)nverts * 12; }
    if (strands > nverts) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, ".hair header: more strands than vertices");
    const uint64_t n_file_indices = 2ull * (nverts - strands);  // read_indices sizes the array from get_segment_count()
    if (bits & 64u) { o_idx = off; off += (size_t)n_file_indices * 4; }
    if (off > n_bytes) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, ".hair file is shorter than its header says");

    cudaStream_t s = ctx->stream;
    RET_IF(stage_in(ctx, ctx->st_vertices, f + o_vert, (size_t)nverts * 12));
    const float* d_vertices = static_cast<const float*>(ctx->st_vertices.p);
    // bounding box: the header's (get_bounding_box, hair_style.cc:236-255), else generate_bounding_box (:215-234)
    if (!(bits & 128u)) {
        float* d_box = reinterpret_cast<float*>(static_cast<char*>(ctx->small.p) + 64);
        RET_IF(vkhr_b200_generate_bounding_box_dev(ctx, d_vertices, nverts, d_box, s));
        CU_CHECK(ctx, cudaMemcpyAsync(bbox, d_box, 24, cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
    }
    const float origin[3] = {bbox[0], bbox[1], bbox[2]};
    const float size[3] = {bbox[3] - bbox[0], bbox[4] - bbox[1], bbox[5] - bbox[2]};
    if (aabb_out) { std::memcpy(aabb_out, origin, 12); std::memcpy(aabb_out + 3, size, 12); }
    // indices: the file's, else generate_indices (hair_style.cc:196-213) from segments[] or default_segment_count
    const uint32_t* d_indices = nullptr;
    uint64_t n_indices = 0;
    uint32_t segs = 0;
    if (bits & 64u) {
        RET_IF(stage_in(ctx, ctx->st_indices, f + o_idx, (size_t)n_file_indices * 4));
        d_indices = static_cast<const uint32_t*>(ctx->st_indices.p);
        n_indices = n_file_indices;
    } else if (bits & 1u) {
        std::vector<uint32_t> prefix((size_t)strands + 1);
        uint64_t total = 0;
        bool uniform = strands > 0;
        uint16_t first_len = 0;
        if (strands) std::memcpy(&first_len, f + o_seg, 2);
        for (uint32_t k = 0; k < strands; ++k) {
            uint16_t len;
            std::memcpy(&len, f + o_seg + (size_t)k * 2, 2);
            prefix[k] = (uint32_t)total;
            total += len;
            uniform = uniform && len == first_len;
        }
        prefix[strands] = (uint32_t)total;
        if (total + strands != nverts) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, ".hair segments[] do not add up to the vertex count");
        if (uniform && first_len > 0) segs = first_len;           // equal strands: no index buffer needed
        else if (total > 0) {
            RET_IF(reserve(ctx, ctx->st_tangents, prefix.size() * 4));         // (free at this point; tangents are staged later)
            CU_CHECK(ctx, cudaMemcpyAsync(ctx->st_tangents.p, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, s));
            RET_IF(reserve(ctx, ctx->st_indices, (size_t)total * 8));
            k_generate_indices<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(static_cast<const uint32_t*>(ctx->st_tangents.p), strands,
                                                                                (uint32_t)total, static_cast<uint2*>(ctx->st_indices.p));
            ctx->launches++;
            CU_CHECK(ctx, cudaGetLastError());
            CU_CHECK(ctx, cudaStreamSynchronize(s));                            // `prefix` dies here; st_tangents is reused below
            d_indices = static_cast<const uint32_t*>(ctx->st_indices.p);
            n_indices = 2 * total;
        }
    } else {
        if (dsegs == 0 || nverts % (dsegs + 1) != 0 || nverts / (dsegs + 1) != strands)
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, ".hair header: strand_count * (default_segment_count + 1) != vertex_count");
        segs = dsegs;
    }
    GridParams g;
    RET_IF(make_grid(ctx, origin, size, W, H, D, flags, g));
    const size_t nv = g.n_voxels;
    RET_IF(reserve(ctx, ctx->st_dens, nv));
    if (tangents_out) RET_IF(reserve(ctx, ctx->st_tang_out, nv * 4));
    if (!d_indices && segs == 0) {                                              // no segments at all: empty volume
        CU_CHECK(ctx, cudaMemsetAsync(ctx->st_dens.p, 0, nv, s));
        if (tangents_out) CU_CHECK(ctx, cudaMemsetAsync(ctx->st_tang_out.p, 0, nv * 4, s));
    } else {
        const float* d_tangents = nullptr;                                      // the file's tangents, else derived on the fly
        if (tangents_out && (bits & 32u)) {
            RET_IF(stage_in(ctx, ctx->st_tangents, f + o_tan, (size_t)nverts * 12));
            d_tangents = static_cast<const float*>(ctx->st_tangents.p);
        }
        RET_IF(vkhr_b200_voxelize_segments_dev(ctx, d_vertices, nverts, d_indices, n_indices, segs, d_tangents, origin, size, W, H, D, flags,
                                               static_cast<uint8_t*>(ctx->st_dens.p),
                                               tangents_out ? static_cast<int8_t*>(ctx->st_tang_out.p) : nullptr, s));
    }
    CU_CHECK(ctx, cudaMemcpyAsync(densities_out, ctx->st_dens.p, nv, cudaMemcpyDeviceToHost, s));
    if (tangents_out) CU_CHECK(ctx, cudaMemcpyAsync(tangents_out, ctx->st_tang_out.p, nv * 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return VKHR_B200_OK;
}

// ---- volumetric ADSM transmittance volume (approximate_deep_shadows.glsl:24-36 at every voxel centre) ----
int vkhr_b200_adsm_dev(vkhr_b200_ctx* ctx, const uint8_t* d_densities, uint32_t W, uint32_t H, uint32_t D,
                       const float origin[3], const float size[3], const vkhr_b200_adsm_params* P, float* d_out, void* stream) {
    RET_IF(bind(ctx));
    if (!d_densities || !d_out || !origin || !size || !P) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null argument");
    const unsigned long long n = (unsigned long long)W * H * D;
    if (n == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "zero resolution");
    if (W > 32768 || H > 32768 || D > 32768 || n >= (1ull << 40)) return fail(ctx, VKHR_B200_ERR_UNSUPPORTED, "resolution too large");
    for (int c = 0; c < 3; ++c)
        if (!std::isfinite(origin[c]) || !std::isfinite(size[c]) || !(size[c] > 0.0f) || !std::isfinite(P->light[c]))
            return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "AABB and light must be finite, size > 0");
    if (!(P->steps >= 1.0f) || !(P->steps <= 65536.0f)) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "steps must be in [1, 65536]");
    cudaStream_t s = pick(ctx, stream);
    // the shader's t sequence: for (t = 0; t < 1; t += 1 / steps), accumulated in fp32
    if (ctx->adsm_steps != P->steps) {
        std::vector<float> table;
        const float step_size = 1.0f / P->steps;
        for (float t = 0.0f; t < 1.0f; t += step_size) { table.push_back(t); if (table.size() > (1u << 20)) break; }
        RET_IF(reserve(ctx, ctx->adsm_table, table.size() * 4));
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->adsm_table.p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));               // `table` dies at the end of this scope
        ctx->adsm_steps = P->steps;
        ctx->adsm_n = (uint32_t)table.size();
    }
    AdsmArgs A;
    A.dens = d_densities; A.W = (int)W; A.H = (int)H; A.D = (int)D;
    A.ox = origin[0]; A.oy = origin[1]; A.oz = origin[2];
    A.sx = size[0]; A.sy = size[1]; A.sz = size[2];
    A.lx = P->light[0]; A.ly = P->light[1]; A.lz = P->light[2];
    A.vsx = size[0] / (float)W; A.vsy = size[1] / (float)H; A.vsz = size[2] / (float)D;
    {   // RN(1 / size) when the FMA division of walk.cuh applies (size in [2^-40, 2^40]), else 0 = IEEE division
        auto recip = [](float v) { return (v >= 9.094947e-13f && v <= 1.0995116e12f) ? 1.0f / v : 0.0f; };
        A.rsx = recip(size[0]); A.rsy = recip(size[1]); A.rsz = recip(size[2]);
    }
    A.t_table = static_cast<const float*>(ctx->adsm_table.p);
    A.n_t = ctx->adsm_n;
    A.step_size = 1.0f / P->steps; A.thickness = P->thickness; A.base = 1.0f - P->strand_alpha;
    A.out = d_out;
    PhaseMark mk(ctx, s, PH_PREFILTER);
    // coarse occupancy bits for empty-space skipping (cells of 4 texels; x0 + 1 ranges over [0, W])
    A.occ_nx32 = (int)((W / 4 + 1 + 31) / 32); A.occ_ny = (int)(H / 4 + 1); A.occ_nz = (int)(D / 4 + 1);
    const size_t occ_words = (size_t)A.occ_nx32 * A.occ_ny * A.occ_nz;
    RET_IF(reserve(ctx, ctx->adsm_occ, occ_words * 4));
    A.occ = static_cast<const uint32_t*>(ctx->adsm_occ.p);
    k_adsm_occupancy<<<stride_blocks(ctx, occ_words * 32, 256, 16), 256, 0, s>>>(d_densities, (int)W, (int)H, (int)D, A.occ_nx32, A.occ_ny, A.occ_nz,
                                                                                 static_cast<uint32_t*>(ctx->adsm_occ.p));
    ctx->launches++;
    k_adsm<<<(unsigned)((n + kAdsmThreads - 1) / kAdsmThreads), kAdsmThreads, 0, s>>>(A);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return VKHR_B200_OK;
}

int vkhr_b200_adsm(vkhr_b200_ctx* ctx, const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
                   const float origin[3], const float size[3], const vkhr_b200_adsm_params* P, float* out) {
    RET_IF(bind(ctx));
    (void)pick(ctx, nullptr);                                   // the context's own stream, ordered behind the previous call
    if (!densities || !out) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null argument");
    const size_t n = (size_t)W * H * D;
    if (n == 0) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "zero resolution");
    RET_IF(stage_in(ctx, ctx->st_dens, densities, n));
    RET_IF(reserve(ctx, ctx->st_tang_out, n * 4));
    RET_IF(vkhr_b200_adsm_dev(ctx, static_cast<const uint8_t*>(ctx->st_dens.p), W, H, D, origin, size, P,
                              static_cast<float*>(ctx->st_tang_out.p), ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(out, ctx->st_tang_out.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return VKHR_B200_OK;
}

// ---- device memory helpers --------------------------------------------------------
int vkhr_b200_malloc(vkhr_b200_ctx* ctx, size_t bytes, void** d_ptr) {
    RET_IF(bind(ctx));
    if (!d_ptr) return fail(ctx, VKHR_B200_ERR_INVALID_ARGUMENT, "null out pointer");
    CU_CHECK(ctx, cudaMalloc(d_ptr, bytes ? bytes : 16));
    return VKHR_B200_OK;
}
int vkhr_b200_free(vkhr_b200_ctx* ctx, void* d_ptr) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaFree(d_ptr));
    return VKHR_B200_OK;
}
int vkhr_b200_memset(vkhr_b200_ctx* ctx, void* d_ptr, int value, size_t bytes, void* stream) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaMemsetAsync(d_ptr, value, bytes, pick(ctx, stream)));
    return VKHR_B200_OK;
}
int vkhr_b200_upload(vkhr_b200_ctx* ctx, void* d_dst, const void* src, size_t bytes, void* stream) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, pick(ctx, stream)));
    return VKHR_B200_OK;
}
int vkhr_b200_download(vkhr_b200_ctx* ctx, void* dst, const void* d_src, size_t bytes, void* stream) {
    RET_IF(bind(ctx));
    CU_CHECK(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, pick(ctx, stream)));
    return VKHR_B200_OK;
}

}  // extern "C"
