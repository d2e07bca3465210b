// tools/microbench2.cu -- B200 micro-benchmarks, part 2: XU-pipe rates (FRND / F2I / MUFU.RCP), and packed
// RED throughput for strand-like address patterns, cold-L2 lines and interleaved arithmetic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
// Prints one JSON object.  Measurement tool only; not part of the product path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// OP 0: floorf (FRND.FLOOR)  1: __float2uint_rz (F2I)  2: MUFU.RCP (__frcp approx)  3: FFMA (reference)  4: magic-add floor (FADD x2 + FSETP + FADD)
template <int OP>
__global__ void k_xu(float* out, uint32_t iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 0.3f, a2 = a0 + 0.6f, a3 = a0 + 0.9f;
    for (uint32_t i = 0; i < iters; ++i) {
        if (OP == 0) { a0 = floorf(a0) + 0.7f; a1 = floorf(a1) + 0.7f; a2 = floorf(a2) + 0.7f; a3 = floorf(a3) + 0.7f; }
        if (OP == 1) { a0 = __uint_as_float(__float2uint_rz(a0)); a1 = __uint_as_float(__float2uint_rz(a1)); a2 = __uint_as_float(__float2uint_rz(a2)); a3 = __uint_as_float(__float2uint_rz(a3)); }
        if (OP == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a1)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a3)); }
        if (OP == 3) { a0 = __fmaf_rn(a0, 1.0001f, 0.5f); a1 = __fmaf_rn(a1, 1.0001f, 0.5f); a2 = __fmaf_rn(a2, 1.0001f, 0.5f); a3 = __fmaf_rn(a3, 1.0001f, 0.5f); }
        if (OP == 4) {
            const float M = 12582912.0f;
            float t0 = __fadd_rn(__fadd_rn(a0, M), -M); if (t0 > a0) t0 -= 1.0f; a0 = t0 + 0.7f;
            float t1 = __fadd_rn(__fadd_rn(a1, M), -M); if (t1 > a1) t1 -= 1.0f; a1 = t1 + 0.7f;
            float t2 = __fadd_rn(__fadd_rn(a2, M), -M); if (t2 > a2) t2 -= 1.0f; a2 = t2 + 0.7f;
            float t3 = __fadd_rn(__fadd_rn(a3, M), -M); if (t3 > a3) t3 -= 1.0f; a3 = t3 + 0.7f;
        }
    }
    if (a0 + a1 + a2 + a3 == 12345.678f) *out = a0;
}

// Packed-byte RED into a 256^3 u8 grid (16 MiB) with a strand-like pattern: a warp is a strand, lane t sits
// t * (dx,dy,dz) voxels along it, and sample k of every lane advances one voxel along the strand's major axis.
// PATTERN 0: uniformly random voxel per lane (reference)   1: strand along mostly -y   2: strand along mostly x
// PAD: dependent FFMAs between reds (models the per-sample arithmetic); XU: floorf per sample as well.
template <int PATTERN, int PAD, int XU>
__global__ void k_red_pattern(uint32_t* grid, uint32_t per_thread, uint32_t active_lanes, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = tid >> 5;
    if (lane >= active_lanes) return;
    float f = (float)tid * 1e-3f;
    for (uint32_t i = 0; i < per_thread; ++i) {
        const uint32_t h = hash32(warp * 0x9E3779B9u + (i >> 1) * 0x85EBCA6Bu);        // a new strand every 2 samples
        uint32_t x = h & 255u, y = (h >> 8) & 255u, z = (h >> 16) & 255u;
        if (PATTERN == 0) { const uint32_t r = hash32(tid * 0x9E3779B9u + i * 0x85EBCA6Bu); x = r & 255u; y = (r >> 8) & 255u; z = (r >> 16) & 255u; }
        if (PATTERN == 1) { x = (x + (lane * 3u) / 4u) & 255u; y = (y + lane * 2u + (i & 1u)) & 255u; z = (z + lane / 2u) & 255u; }
        if (PATTERN == 2) { x = (x + lane * 2u + (i & 1u)) & 255u; y = (y + (lane * 3u) / 4u) & 255u; z = (z + lane / 2u) & 255u; }
#pragma unroll
        for (int k = 0; k < PAD; ++k) f = __fmaf_rn(f, 1.0000001f, 0.5f);
        if (XU) { f = floorf(f) + floorf(f * 0.5f) + floorf(f * 0.25f) + 0.3f; x ^= (__float2uint_rz(f) & 1u); }
        if (PAD) x ^= (__float_as_uint(f) & 1u);
        const uint32_t idx = x + (y << 8) + (z << 16);
        atomicAdd(grid + (idx >> 2), 1u << ((idx & 3u) * 8u));
    }
    if (f == 12345.678f) *out = f;
}

// The same with shared-memory loads / shuffles / global loads inside the arithmetic between two reds, as a real
// walk has them: MIO 1 = 3 dependent LDS, 2 = 3 dependent SHFL, 3 = 1 dependent LDG (L2-resident table), 0 = none.
template <int MIO, int HALF = 48>
__global__ void k_red_mio(uint32_t* grid, const float* table, uint32_t per_thread, float* out) {
    __shared__ float s[256 * 3];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = tid >> 5;
    for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) s[i] = (float)i * 1e-3f;
    __syncthreads();
    float f = (float)tid * 1e-3f;
    for (uint32_t i = 0; i < per_thread; ++i) {
        const uint32_t h = hash32(warp * 0x9E3779B9u + (i >> 1) * 0x85EBCA6Bu);
        uint32_t x = (h + (lane * 3u) / 4u) & 255u, y = ((h >> 8) + lane * 2u + (i & 1u)) & 255u, z = ((h >> 16) + lane / 2u) & 255u;
#pragma unroll
        for (int k = 0; k < HALF; ++k) f = __fmaf_rn(f, 1.0000001f, 0.5f);
        if (MIO == 1) { const uint32_t j = (__float_as_uint(f) >> 3) & 255u; f += s[3 * j] + s[3 * j + 1] + s[3 * j + 2]; }
        if (MIO == 2) { f += __shfl_down_sync(0xFFFFFFFFu, f, 1); f += __shfl_down_sync(0xFFFFFFFFu, f * 0.5f, 1); f += __shfl_down_sync(0xFFFFFFFFu, f * 0.25f, 1); }
        if (MIO == 3) { const uint32_t j = (__float_as_uint(f) >> 3) & 0xFFFFu; f += __ldg(table + j); }
#pragma unroll
        for (int k = 0; k < HALF; ++k) f = __fmaf_rn(f, 1.0000001f, 0.5f);
        x ^= (__float_as_uint(f) & 1u);
        const uint32_t idx = x + (y << 8) + (z << 16);
        atomicAdd(grid + (idx >> 2), 1u << ((idx & 3u) * 8u));
    }
    if (f == 12345.678f) *out = f;
}
// no reds at all: the arithmetic + MIO part alone
template <int MIO, int HALF = 48>
__global__ void k_nored_mio(const float* table, uint32_t per_thread, float* out) {
    __shared__ float s[256 * 3];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) s[i] = (float)i * 1e-3f;
    __syncthreads();
    float f = (float)tid * 1e-3f;
    for (uint32_t i = 0; i < per_thread; ++i) {
#pragma unroll
        for (int k = 0; k < HALF; ++k) f = __fmaf_rn(f, 1.0000001f, 0.5f);
        if (MIO == 1) { const uint32_t j = (__float_as_uint(f) >> 3) & 255u; f += s[3 * j] + s[3 * j + 1] + s[3 * j + 2]; }
        if (MIO == 2) { f += __shfl_down_sync(0xFFFFFFFFu, f, 1); f += __shfl_down_sync(0xFFFFFFFFu, f * 0.5f, 1); f += __shfl_down_sync(0xFFFFFFFFu, f * 0.25f, 1); }
        if (MIO == 3) { const uint32_t j = (__float_as_uint(f) >> 3) & 0xFFFFu; f += __ldg(table + j); }
#pragma unroll
        for (int k = 0; k < HALF; ++k) f = __fmaf_rn(f, 1.0000001f, 0.5f);
    }
    if (f == 12345.678f) *out = f;
}

// A synthetic copy of the walk kernel's control structure: per "tile" a block of arithmetic with LDS + SHFL + VOTE,
// then a per-lane sample loop with a data-dependent trip count (DIVERGE) whose body ends in a predicated red.
// RED 0 = no red (index folded into a register), 1 = red.  UNIFORM trip count = 2 when DIVERGE == 0.
template <int RED, int DIVERGE, int VOTE, int STAGE = 0>
__global__ void k_mimic(uint32_t* grid, uint32_t tiles, float* out, const float* stream = nullptr) {
    __shared__ float s[256 * 3];
    __shared__ float st[8][768];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = tid >> 5;
    for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) s[i] = (float)i * 1e-3f;
    __syncthreads();
    if (STAGE) {                                   // every warp first streams its own 3 KB from HBM (as the walk kernel does)
        float v[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) v[j] = __ldg(stream + (size_t)warp * 768 + lane + 32 * j);
#pragma unroll
        for (int j = 0; j < 24; ++j) st[threadIdx.x >> 5][lane + 32 * j] = v[j];
        __syncwarp();
    }
    float f = (float)tid * 1e-3f;
    uint32_t acc = 0;
    for (uint32_t t = 0; t < tiles; ++t) {
        const uint32_t h = hash32(warp * 0x9E3779B9u + t * 0x85EBCA6Bu);
        const uint32_t j = (h + lane) & 255u;
        float a = s[3 * j], b = s[3 * j + 1], c = s[3 * j + 2];
        if (STAGE) { a += st[threadIdx.x >> 5][93 * (t & 7u) + 3 * lane]; b += st[threadIdx.x >> 5][93 * (t & 7u) + 3 * lane + 1]; c += st[threadIdx.x >> 5][93 * (t & 7u) + 3 * lane + 2]; }
#pragma unroll
        for (int k = 0; k < 20; ++k) { a = __fmaf_rn(a, 1.0000001f, f); b = __fmaf_rn(b, 1.0000001f, a); c = __fmaf_rn(c, 1.0000001f, b); }
        if (VOTE) { if (!__all_sync(0xFFFFFFFFu, a == a)) c += 1.0f; }
        const float ta = __shfl_down_sync(0xFFFFFFFFu, a, 1), tb = __shfl_down_sync(0xFFFFFFFFu, b, 1), tc = __shfl_down_sync(0xFFFFFFFFu, c, 1);
        float da = ta - a, db = tb - b, dc = tc - c;
#pragma unroll
        for (int k = 0; k < 15; ++k) { da = __fmaf_rn(da, 0.999f, dc); db = __fmaf_rn(db, 0.999f, da); dc = __fmaf_rn(dc, 0.999f, db); }
        if (VOTE) { if (!__all_sync(0xFFFFFFFFu, da == da)) dc += 1.0f; }
        const uint32_t hh = hash32(tid * 0x9E3779B9u + t);
        const bool active = (hh & 7u) != 0u;                                   // 28 of 32 lanes
        uint32_t n = DIVERGE ? (1u + ((hh >> 3) & 1u) + (((hh >> 4) & 15u) == 0u ? 1u : 0u) + ((hh >> 8) & 1u)) : 2u;   // mean ~2.06
        uint32_t x = hh & 255u, y = (hh >> 8) & 255u, z = (hh >> 16) & 255u;
        if (active) {
            do {
                const float fx = fminf(floorf(a), 255.0f), fy = fminf(floorf(b), 255.0f), fz = fminf(floorf(c), 255.0f);
                const float fi = __fadd_rn(__fadd_rn(fx, __fmul_rn(fy, 256.0f)), __fmul_rn(__fmul_rn(fz, 256.0f), 256.0f));
                uint32_t idx = (__float2uint_rz(fi) & 1u) ^ (x + (y << 8) + (z << 16));
                if (fi >= 0.0f && idx < (1u << 24)) {
                    if (RED) atomicAdd(grid + (idx >> 2), 1u << ((idx & 3u) * 8u)); else acc += idx;
                }
                a += da; b += db; c += dc; y = (y + 1u) & 255u;
            } while (--n);
        }
        f = a * 1e-9f;
    }
    if (f == 12345.678f || acc == 0xDEADBEEF) *out = f;
}

__global__ void k_fill(uint4* p, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
// one cold run: the grid is zeroed, then L2 is flushed by filling a 512 MiB buffer, then f is timed once
template <class F> float time_cold_ms(F f, uint32_t* grid, size_t grid_bytes, uint4* flush, size_t flush_bytes, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float tot = 0;
    for (int i = 0; i < reps; ++i) {
        k_fill<<<148 * 8, 256>>>((uint4*)grid, grid_bytes / 16);
        k_fill<<<148 * 8, 256>>>(flush, flush_bytes / 16);
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); tot += ms;
    }
    return tot / reps;
}

int main() {
    uint32_t* grid; float* out; uint4* flush;
    const size_t GRID = 16u << 20, FLUSH = 512u << 20;
    CK(cudaMalloc(&grid, GRID)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&flush, FLUSH));
    CK(cudaMemset(grid, 0, GRID));
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const double clk = prop.clockRate * 1e3;   // Hz (max)
    printf("{\n");
    {
        const int blocks = 148 * 8, threads = 256; const uint32_t iters = 4096;
        const double lane_ops = (double)blocks * threads * iters * 4;
        float t0 = time_ms([&] { k_xu<0><<<blocks, threads>>>(out, iters, 1.5f); }, 3);
        float t1 = time_ms([&] { k_xu<1><<<blocks, threads>>>(out, iters, 1.5f); }, 3);
        float t2 = time_ms([&] { k_xu<2><<<blocks, threads>>>(out, iters, 1.5f); }, 3);
        float t3 = time_ms([&] { k_xu<3><<<blocks, threads>>>(out, iters, 1.5f); }, 3);
        float t4 = time_ms([&] { k_xu<4><<<blocks, threads>>>(out, iters, 1.5f); }, 3);
        auto per_clk_sm = [&](float ms) { return lane_ops / (ms * 1e-3) / clk / 148.0; };
        printf(" \"pipe_rates_lanes_per_clk_per_sm_at_max_clock\": {\"floorf_plus_fadd\": %.1f, \"f2i\": %.1f, \"mufu_rcp\": %.1f, \"ffma\": %.1f, \"magic_floor_plus_fadd\": %.1f},\n",
               per_clk_sm(t0), per_clk_sm(t1), per_clk_sm(t2), per_clk_sm(t3), per_clk_sm(t4));
    }
    {
        const int blocks = 148 * 64, threads = 256; const uint32_t per = 16;
        const double nops = (double)blocks * threads * per;
#define RUN(name, P, PAD, XU, act) { \
            float w = time_ms([&] { k_red_pattern<P, PAD, XU><<<blocks, threads>>>(grid, per, act, out); }, 5); \
            float c = time_cold_ms([&] { k_red_pattern<P, PAD, XU><<<blocks, threads>>>(grid, per, act, out); }, grid, GRID, flush, FLUSH, 3); \
            printf(" \"%s\": {\"warm_Gops\": %.1f, \"cold_Gops\": %.1f},\n", name, nops * act / 32 / w / 1e6, nops * act / 32 / c / 1e6); }
        RUN("red_packed_random", 0, 0, 0, 32u)
        RUN("red_packed_random_28lanes", 0, 0, 0, 28u)
        RUN("red_packed_strand_y", 1, 0, 0, 32u)
        RUN("red_packed_strand_y_28lanes", 1, 0, 0, 28u)
        RUN("red_packed_strand_x", 2, 0, 0, 32u)
        RUN("red_packed_strand_y_pad32", 1, 32, 0, 28u)
        RUN("red_packed_strand_y_pad32_xu", 1, 32, 1, 28u)
        RUN("red_packed_strand_y_pad96_xu", 1, 96, 1, 28u)
    }
    {
        const int blocks = 148 * 64, threads = 256; const uint32_t per = 16;
        const double nops = (double)blocks * threads * per;
        const float* table = (const float*)flush;
#define RUNM(name, M) { \
            float w = time_ms([&] { k_red_mio<M><<<blocks, threads>>>(grid, table, per, out); }, 5); \
            float n = time_ms([&] { k_nored_mio<M><<<blocks, threads>>>(table, per, out); }, 5); \
            printf(" \"%s\": {\"with_red_us\": %.1f, \"without_red_us\": %.1f, \"red_alone_at_220G_us\": %.1f, \"red_Gops\": %.1f},\n", name, w * 1e3, n * 1e3, nops / 220e3, nops / w / 1e6); }
        RUNM("red_fma96", 0)
        RUNM("red_fma96_lds3", 1)
        RUNM("red_fma96_shfl3", 2)
        RUNM("red_fma96_ldg1", 3)
#define RUNP(name, H) { \
            float w = time_ms([&] { k_red_mio<1, H><<<blocks, threads>>>(grid, table, per, out); }, 5); \
            float n = time_ms([&] { k_nored_mio<1, H><<<blocks, threads>>>(table, per, out); }, 5); \
            printf(" \"%s\": {\"with_red_us\": %.1f, \"without_red_us\": %.1f, \"red_alone_at_220G_us\": %.1f},\n", name, w * 1e3, n * 1e3, nops / 220e3); }
        RUNP("balance_fma128_lds3", 64)
        RUNP("balance_fma160_lds3", 80)
        RUNP("balance_fma192_lds3", 96)
        RUNP("balance_fma256_lds3", 128)
        RUNP("balance_fma384_lds3", 192)
    }
    {
        const int blocks = 148 * 6 * 4, threads = 256; const uint32_t tiles = 8;     // ~4 waves of 6 CTAs/SM
        const double lanes = (double)blocks * threads * tiles * 0.875 * 2.06;
#define RUNK(name, D, V) { \
            float w = time_ms([&] { k_mimic<1, D, V><<<blocks, threads>>>(grid, tiles, out); }, 5); \
            float n = time_ms([&] { k_mimic<0, D, V><<<blocks, threads>>>(grid, tiles, out); }, 5); \
            printf(" \"%s\": {\"with_red_us\": %.1f, \"without_red_us\": %.1f, \"red_alone_at_220G_us\": %.1f},\n", name, w * 1e3, n * 1e3, lanes / 220e3); }
        RUNK("mimic_uniform_novote", 0, 0)
        RUNK("mimic_uniform_vote", 0, 1)
        RUNK("mimic_diverge_novote", 1, 0)
        RUNK("mimic_diverge_vote", 1, 1)
        {
            const float* stream = (const float*)flush;      // 512 MiB, far larger than what one launch reads (87 MB)
            float w = time_ms([&] { k_mimic<1, 1, 1, 1><<<blocks, threads>>>(grid, tiles, out, stream); }, 5);
            float n = time_ms([&] { k_mimic<0, 1, 1, 1><<<blocks, threads>>>(grid, tiles, out, stream); }, 5);
            printf(" \"mimic_diverge_vote_staged\": {\"with_red_us\": %.1f, \"without_red_us\": %.1f, \"red_alone_at_220G_us\": %.1f},\n", w * 1e3, n * 1e3, lanes / 220e3);
        }
    }
    printf(" \"sm_clock_max_mhz\": %.0f\n}\n", clk / 1e6);
    return 0;
}
