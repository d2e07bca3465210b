// prefilter.cuh -- density -> ambient-occlusion / opacity / Gaussian prefilter: a TMA-staged 3-D stencil pass.
//
// What the reference's fragment shaders compute from the density volume at every shaded point, evaluated
// once per voxel centre (paths relative to the reference tree):
//   local_ambient_occlusion            share/shaders/volumes/local_ambient_occlusion.glsl:9-30
//   volume_approximated_deep_shadows   share/shaders/self-shadowing/approximate_deep_shadows.glsl:24-36 (one step)
//   filter_volume                      share/shaders/volumes/sample_volume.glsl:12-35
//   sample_volume / sampler            sample_volume.glsl:7-9; R8_UNORM, LINEAR, CLAMP_TO_BORDER (black):
//                                      src/vkhr/rasterizer/hair_style.cc:79-85,:94-101
// The arithmetic contract (texel decode, tap geometry, nested x-y-z weighted sums, accumulation order) is the
// one written out in prefilter_oracle.c of the test infrastructure; every fp32 operation below is a separately rounded
// __f*_rn intrinsic in that order, so everything up to the final powf is bit-identical to the oracle.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "walk.cuh"      // div_exact: the correctly rounded quotient without the division instruction sequence

namespace vkhr_b200 {

constexpr int kPfTX = 32, kPfTY = 8;                 // output voxels per tile along x, y
// ... and along z: 8, or 16 where two CTAs of the deeper tile still fit an SM (the column form of AO shares its x/y
// lerps along z, and the halo planes are a smaller share of a deeper tile: 22/16 instead of 14/8 at radius 2.5)
constexpr int kPfTZ = 8, kPfTZDeep = 16;
// The staged box starts kPfLead texels left of the tile: the innermost TMA coordinate must be a multiple of 16 BYTES
// (measured: tools/tma_probe.cu -- x = -16, 16, 48 load, x = -3, 4, 29 raise "illegal instruction"), so the halo
// cannot start at x0 - halo; the box is [x0 - 16, x0 + 48) and the kernel uses [x0 - halo, x0 + 32 + halo) of it.
constexpr int kPfLead = 16;
constexpr int kPfBX = kPfTX + 2 * kPfLead;           // 64 bytes per staged row (TMA inner box extent)
constexpr int kPfMaxHalo = 8;                        // shared-memory budget (the float tile grows with halo^3)
constexpr int kPfThreads = 256;
constexpr int kPfMaxGauss = 9;                       // largest Gaussian kernel width

struct AxisTaps { int o0, o1; float w0, w1; };       // texels i + o0, i + o1 and their weights

struct PrefilterArgs {
    const uint8_t* dens;
    int W, H, D;
    int halo;                  // texels needed on each side of a tile
    float* ao;                 // outputs (nullptr = not wanted)
    float* opacity;
    float* gauss;
    AxisTaps neg, pos;         // LAO taps at -radius / +radius voxels
    float ao_max, ao_exponent;
    float one_minus_alpha, thickness;
    int g_range;               // (kernel_width - 1) / 2
    float g_sigma2;
    uint32_t tiles_x, tiles_y, tiles_z;
    uint32_t* tile_counter;     // dynamic tile scheduling: CTAs draw tiles from this counter (zeroed by k_pf_tile_active); nullptr = static stride
    const uint8_t* tile_active; // one byte per tile (k_pf_tile_active): 0 = the tile and its halo hold no hair -> constants, no load; or nullptr
};

// ---- the arithmetic, shared by the tiled and the generic kernel ---------------------------------------------
// fetch(ox, oy, oz) = tau of the texel displaced by (ox, oy, oz) from this thread's voxel (0 outside the grid).

template <class Fetch>
__device__ __forceinline__ float lao_at(const PrefilterArgs& A, Fetch&& fetch) {
    const AxisTaps ax[2] = {A.neg, A.pos};
    // X-lerps for the 4 x 4 (y,z) rows of the footprint, Y-lerps kept per z row
    float yl[2][2][4];                                   // [sx][sy][z row]
#pragma unroll
    for (int zr = 0; zr < 4; ++zr) {
        const int oz = (zr & 1) ? ax[zr >> 1].o1 : ax[zr >> 1].o0;
        float xl[2][4];                                  // [sx][y row]
#pragma unroll
        for (int yr = 0; yr < 4; ++yr) {
            const int oy = (yr & 1) ? ax[yr >> 1].o1 : ax[yr >> 1].o0;
#pragma unroll
            for (int sx = 0; sx < 2; ++sx) {
                const float t0 = fetch(ax[sx].o0, oy, oz), t1 = fetch(ax[sx].o1, oy, oz);
                xl[sx][yr] = __fadd_rn(__fmul_rn(t0, ax[sx].w0), __fmul_rn(t1, ax[sx].w1));
            }
        }
#pragma unroll
        for (int sx = 0; sx < 2; ++sx)
#pragma unroll
            for (int sy = 0; sy < 2; ++sy)
                yl[sx][sy][zr] = __fadd_rn(__fmul_rn(xl[sx][2 * sy], ax[sy].w0), __fmul_rn(xl[sx][2 * sy + 1], ax[sy].w1));
    }
    float density = 0.0f;
#pragma unroll
    for (int sz = 0; sz < 2; ++sz)
#pragma unroll
        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
            for (int sx = 0; sx < 2; ++sx) {
                const float s = __fadd_rn(__fmul_rn(yl[sx][sy][2 * sz], ax[sz].w0), __fmul_rn(yl[sx][sy][2 * sz + 1], ax[sz].w1));
                density = __fadd_rn(density, (A.ao_max < s) ? A.ao_max : s);
            }
    return powf(__fsub_rn(1.0f, __fdiv_rn(density, 8.0f)), A.ao_exponent);    // pow(kernel_size = 2, 3) = 8
}

// The same sum for a COLUMN of TZ voxels along z (one thread: fixed x, y): the x- and y-lerps of plane z' are
// shared by the four outputs z = z' - o (o = the four z offsets), so they are computed once per plane instead of once
// per output -- 8 x-lerps + 4 y-lerps per plane, 8 z-lerps per output; per voxel 28 instead of 64 texel loads and 29
// instead of 56 lerps at TZ = 8, radius 2.5.  Every lerp is the operation sequence of lao_at, and an output adds its
// eight clamped z-lerps in lao_at's (sz, sy, sx) order, so the result is bit-identical to lao_at.  The tap offsets
// are template parameters (NO0, NO0 + 1 below, PO0, PO0 + 1 above the voxel): the loop over planes unrolls fully and
// the window of live planes (PO0 - NO0 + 2 planes x 4 floats) stays in registers.
// A plane whose four tap rows hold no hair contributes exact zeros (+0 * w + +0 * w, w >= 0) and is not read; an
// output whose four planes are all such planes is the empty-space constant.
template <int NO0, int PO0, int TZ, class Emit>
__device__ __forceinline__ void lao_column(const PrefilterArgs& A, const float* __restrict__ column, const uint32_t* __restrict__ flagcol,
                                           int plane_floats, int row_floats, int BY, bool tile_any, float ao_empty, int n_out, Emit&& emit) {
    // column  = texel (lane, jy, kz = 0) of the float tile; flagcol = row flag of (jy, kz = 0)
    constexpr int SPAN = PO0 - NO0 + 1;                  // planes between the first and the last tap plane of an output
    constexpr int NQ = TZ + SPAN;                        // planes q = 0 .. NQ-1 sit at z offset q + NO0 from output 0
    const int oy[4] = {NO0, NO0 + 1, PO0, PO0 + 1};
    const AxisTaps ax[2] = {A.neg, A.pos};
    float yl[NQ][4];                                     // [q][2 * sx + sy]
    bool live[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        uint32_t any = 0u;
        if (tile_any) {
#pragma unroll
            for (int yr = 0; yr < 4; ++yr) any |= flagcol[(q + NO0) * BY + oy[yr]];
        }
        live[q] = any != 0u;
        if (any) {
            const float* pl = column + (q + NO0) * plane_floats;
            float xl[2][4];
#pragma unroll
            for (int yr = 0; yr < 4; ++yr) {
                const float* row = pl + oy[yr] * row_floats;
                xl[0][yr] = __fadd_rn(__fmul_rn(row[NO0], ax[0].w0), __fmul_rn(row[NO0 + 1], ax[0].w1));
                xl[1][yr] = __fadd_rn(__fmul_rn(row[PO0], ax[1].w0), __fmul_rn(row[PO0 + 1], ax[1].w1));
            }
#pragma unroll
            for (int sx = 0; sx < 2; ++sx)
#pragma unroll
                for (int sy = 0; sy < 2; ++sy)
                    yl[q][2 * sx + sy] = __fadd_rn(__fmul_rn(xl[sx][2 * sy], ax[sy].w0), __fmul_rn(xl[sx][2 * sy + 1], ax[sy].w1));
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) yl[q][c] = 0.0f;
        }
        if (q >= SPAN) {                                 // output kz = q - SPAN has its last plane
            const int kz = q - SPAN;
            if (kz < n_out) {                            // warp-uniform
                float r = ao_empty;
                if (live[kz] || live[kz + 1] || live[kz + SPAN - 1] || live[kz + SPAN]) {
                    float density = 0.0f;
#pragma unroll
                    for (int sz = 0; sz < 2; ++sz) {
                        const int qa = kz + (sz ? SPAN - 1 : 0);
#pragma unroll
                        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                            for (int sx = 0; sx < 2; ++sx) {
                                const float s = __fadd_rn(__fmul_rn(yl[qa][2 * sx + sy], ax[sz].w0), __fmul_rn(yl[qa + 1][2 * sx + sy], ax[sz].w1));
                                density = __fadd_rn(density, (A.ao_max < s) ? A.ao_max : s);
                            }
                    }
                    r = powf(__fsub_rn(1.0f, __fdiv_rn(density, 8.0f)), A.ao_exponent);
                }
                emit(kz, r);
            }
        }
    }
}

// lao_column with the plane loop unrolled by the PERIOD of the register window (SPAN + 1 planes) instead of fully:
// plane q lives in slot q mod (SPAN + 1), so inside one period every slot index is a compile-time constant and the code
// is (SPAN + 1) / (TZ + SPAN) of the fully unrolled form (a third at radius 2.5, TZ = 16) -- the fully unrolled kernel
// spends its third largest stall waiting for instructions.  Same operations in the same order:
// bit-identical (tests/host_emulation on the CPU, the column-form tests on the GPU).  Measured on B200: AO at 512^3
// 1.11 -> 0.94 ms (5040 instead of 8304 SASS instructions for <-3, 2, 16>); this is the form the kernel calls.
template <int NO0, int PO0, int TZ, class Emit>
__device__ __forceinline__ void lao_column_rolled(const PrefilterArgs& A, const float* __restrict__ column, const uint32_t* __restrict__ flagcol,
                                                  int plane_floats, int row_floats, int BY, bool tile_any, float ao_empty, int n_out, Emit&& emit) {
    constexpr int SPAN = PO0 - NO0 + 1;
    constexpr int P = SPAN + 1;                          // window period
    constexpr int NQ = TZ + SPAN;
    const int oy[4] = {NO0, NO0 + 1, PO0, PO0 + 1};
    const AxisTaps ax[2] = {A.neg, A.pos};
    float yl[P][4];                                      // [q mod P][2 * sx + sy]
    bool live[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        live[j] = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) yl[j][c] = 0.0f;
    }
#pragma unroll 1
    for (int q0 = 0; q0 < NQ; q0 += P) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int q = q0 + j;                        // q mod P == j: q0 is a multiple of P
            if (q >= NQ) break;                          // warp-uniform
            uint32_t any = 0u;
            if (tile_any) {
#pragma unroll
                for (int yr = 0; yr < 4; ++yr) any |= flagcol[(q + NO0) * BY + oy[yr]];
            }
            live[j] = any != 0u;
            if (any) {
                const float* pl = column + (q + NO0) * plane_floats;
                float xl[2][4];
#pragma unroll
                for (int yr = 0; yr < 4; ++yr) {
                    const float* row = pl + oy[yr] * row_floats;
                    xl[0][yr] = __fadd_rn(__fmul_rn(row[NO0], ax[0].w0), __fmul_rn(row[NO0 + 1], ax[0].w1));
                    xl[1][yr] = __fadd_rn(__fmul_rn(row[PO0], ax[1].w0), __fmul_rn(row[PO0 + 1], ax[1].w1));
                }
#pragma unroll
                for (int sx = 0; sx < 2; ++sx)
#pragma unroll
                    for (int sy = 0; sy < 2; ++sy)
                        yl[j][2 * sx + sy] = __fadd_rn(__fmul_rn(xl[sx][2 * sy], ax[sy].w0), __fmul_rn(xl[sx][2 * sy + 1], ax[sy].w1));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) yl[j][c] = 0.0f;
            }
            const int kz = q - SPAN;                     // the output whose last plane this is
            if (kz >= 0 && kz < n_out) {                 // warp-uniform
                // its planes kz, kz + 1, kz + SPAN - 1, kz + SPAN sit in slots j + 1, j + 2, j - 1, j (mod P)
                const int s0 = (j + 1) % P, s1 = (j + 2) % P, s2 = (j + P - 1) % P, s3 = j;
                float r = ao_empty;
                if (live[s0] || live[s1] || live[s2] || live[s3]) {
                    float density = 0.0f;
#pragma unroll
                    for (int sz = 0; sz < 2; ++sz) {
                        const int qa = sz ? s2 : s0, qb = sz ? s3 : s1;
#pragma unroll
                        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                            for (int sx = 0; sx < 2; ++sx) {
                                const float s = __fadd_rn(__fmul_rn(yl[qa][2 * sx + sy], ax[sz].w0), __fmul_rn(yl[qb][2 * sx + sy], ax[sz].w1));
                                density = __fadd_rn(density, (A.ao_max < s) ? A.ao_max : s);
                            }
                    }
                    r = powf(__fsub_rn(1.0f, __fdiv_rn(density, 8.0f)), A.ao_exponent);
                }
                emit(kz, r);
            }
        }
    }
}

// Gaussian weight of tap (x,y,z) (sample_volume.glsl:26-27, precedence quirk kept).
__device__ __forceinline__ float gauss_weight(float x, float y, float z, float sigma2) {
    const float e = __fmul_rn(__fdiv_rn(__fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))), 2.0f), sigma2);
    const float pi = 3.14159265358979323846f, euler = 2.71828182845904523536f;
    return __fmul_rn(__fdiv_rn(1.0f, __fmul_rn(__fmul_rn(2.0f, pi), sigma2)), powf(euler, e));
}

// total_weight (sample_volume.glsl:31) is the same sum for every voxel -- the weights in loop order -- so a thread
// adds it up once (gauss_total) instead of once per voxel; the quotient is the same operation on the same operands.
__device__ __forceinline__ float gauss_total(const PrefilterArgs& A, const float* __restrict__ w) {
    const int N = 2 * A.g_range + 1;
    float total = 0.0f;
    for (int t = 0; t < N * N * N; ++t) total = __fadd_rn(total, w[t]);
    return total;
}

template <class Fetch>
__device__ __forceinline__ float gauss_at(const PrefilterArgs& A, const float* __restrict__ w, float total, Fetch&& fetch) {
    const int R = A.g_range, N = 2 * R + 1;
    float density = 0.0f;
    for (int z = -R; z <= R; ++z)
        for (int y = -R; y <= R; ++y)
            for (int x = -R; x <= R; ++x)
                density = __fadd_rn(density, __fmul_rn(fetch(x, y, z), w[((z + R) * N + (y + R)) * N + (x + R)]));
    return __fdiv_rn(density, total);
}

// the same sum with the kernel width as a compile-time constant: the tap loop unrolls, tap addresses become
// immediate offsets off one row pointer per (y, z) -- same operations, same order, a third of the instructions
template <int R, class Fetch>
__device__ __forceinline__ float gauss_at_fixed(const float* __restrict__ w, float total, Fetch&& fetch) {
    constexpr int N = 2 * R + 1;
    float density = 0.0f;
#pragma unroll
    for (int z = -R; z <= R; ++z)
#pragma unroll
        for (int y = -R; y <= R; ++y)
#pragma unroll
            for (int x = -R; x <= R; ++x)
                density = __fadd_rn(density, __fmul_rn(fetch(x, y, z), w[((z + R) * N + (y + R)) * N + (x + R)]));
    return __fdiv_rn(density, total);
}

__device__ __forceinline__ float opacity_of(const PrefilterArgs& A, uint32_t d) {
    return powf(A.one_minus_alpha, __fmul_rn(__fdiv_rn((float)d, 255.0f), A.thickness));
}

#ifndef VKHR_PF_TILED_ONLY      // the non-template kernels live in ONE translation unit (vkhr_b200.cu)
// ---- occupancy pre-pass of the tiled kernel ------------------------------------------------------------------
// Cells of 32 x 8 x 8 voxels (a quarter or half of a tile): occ[cell] = 1 when any voxel of the cell is non-zero.  One
// warp per cell, four 16-byte loads per lane (W % 16 == 0 on the tiled path).  Reads the volume once: N^3 bytes.
constexpr int kPfCellX = 32, kPfCellY = 8, kPfCellZ = 8;
__global__ void __launch_bounds__(256)
k_pf_cell_occupancy(const uint8_t* __restrict__ dens, int W, int H, int D, int cx, int cy, int cz, uint8_t* __restrict__ occ) {
    const uint32_t n_cells = (uint32_t)cx * cy * cz;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cell < n_cells; cell += gridDim.x * (blockDim.x >> 5)) {
        const int x0 = (int)(cell % cx) * kPfCellX, y0 = (int)((cell / cx) % cy) * kPfCellY, z0 = (int)(cell / ((uint32_t)cx * cy)) * kPfCellZ;
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int chunk = (int)lane + 32 * k;                               // 128 chunks: 64 rows x 2 halves
            const int X = x0 + 16 * (chunk & 1), Y = y0 + ((chunk >> 1) & 7), Z = z0 + (chunk >> 4);
            if (X < W && Y < H && Z < D) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(dens + (size_t)X + (size_t)Y * W + (size_t)Z * W * H));
                any |= q.x | q.y | q.z | q.w;
            }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, any != 0u);
        if (lane == 0) occ[cell] = b ? 1 : 0;
    }
}
// active[tile] = 1 when a cell overlapping the tile's box (tile + halo, halo <= 8) is occupied.  One thread per tile.
__global__ void __launch_bounds__(256)
k_pf_tile_active(const uint8_t* __restrict__ occ, int cx, int cy, int cz, uint32_t tiles_x, uint32_t tiles_y, uint32_t tiles_z, int tz,
                 uint8_t* __restrict__ active, uint32_t* __restrict__ counter) {
    const uint32_t n = tiles_x * tiles_y * tiles_z;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *counter = 0u;                                     // the tiled kernel draws its tiles from here
    if (t >= n) return;
    const int tx = (int)(t % tiles_x), ty = (int)((t / tiles_x) % tiles_y), tzi = (int)(t / (tiles_x * tiles_y));
    // tiles are kPfTX x kPfTY x tz voxels = 1 x 1 x (tz / 8) cells
    const int zc0 = tzi * (tz / kPfCellZ) - 1, zc1 = (tzi + 1) * (tz / kPfCellZ);
    uint32_t any = 0;
    for (int z = max(zc0, 0); z <= min(zc1, cz - 1); ++z)
        for (int y = max(ty - 1, 0); y <= min(ty + 1, cy - 1); ++y)
            for (int x = max(tx - 1, 0); x <= min(tx + 1, cx - 1); ++x) any |= occ[((size_t)z * cy + y) * cx + x];
    active[t] = any ? 1 : 0;
}

// ---- generic kernel: any grid size / alignment / radius; one thread per voxel, texels straight from global ----
__global__ void __launch_bounds__(256)
k_prefilter_generic(const __grid_constant__ PrefilterArgs A) {
    __shared__ float s_gw[kPfMaxGauss * kPfMaxGauss * kPfMaxGauss];
    if (A.gauss) {
        const int N = 2 * A.g_range + 1;
        for (int t = threadIdx.x; t < N * N * N; t += blockDim.x)
            s_gw[t] = gauss_weight((float)(t % N - A.g_range), (float)((t / N) % N - A.g_range), (float)(t / (N * N) - A.g_range), A.g_sigma2);
        __syncthreads();
    }
    const float g_total = A.gauss ? gauss_total(A, s_gw) : 1.0f;
    const uint64_t n = (uint64_t)A.W * A.H * A.D;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (uint64_t)gridDim.x * blockDim.x) {
        const int i = (int)(v % A.W), j = (int)((v / A.W) % A.H), k = (int)(v / ((uint64_t)A.W * A.H));
        auto fetch = [&](int ox, int oy, int oz) -> float {
            const int x = i + ox, y = j + oy, z = k + oz;
            if (x < 0 || y < 0 || z < 0 || x >= A.W || y >= A.H || z >= A.D) return 0.0f;
            return __fdiv_rn((float)__ldg(A.dens + (size_t)x + (size_t)y * A.W + (size_t)z * A.W * A.H), 255.0f);
        };
        if (A.ao) A.ao[v] = lao_at(A, fetch);
        if (A.opacity) A.opacity[v] = opacity_of(A, __ldg(A.dens + v));
        if (A.gauss) A.gauss[v] = gauss_at(A, s_gw, g_total, fetch);
    }
}

#endif  // VKHR_PF_TILED_ONLY

// ---- tiled kernel: persistent CTAs, 3-D TMA box loads (zero fill outside the grid = the black border) ----------
__device__ __forceinline__ uint32_t pf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pf_tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void pf_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();                 // a lost copy must fail, not hang the device
    }
}

// Shared-memory plan (dynamic): [stage 0][stage 1][float tile][row flags][decode LUT][opacity LUT][gauss weights][2 mbarriers]
struct PfSmemPlan {
    uint32_t stage_bytes, stage1, ftile, flags, lut, oplut, gw, bars, total;
};
// The float tile keeps whole 32-bit words of the staged rows: its x halo is the halo rounded up to 4 texels, so a
// staged word is either converted as a whole (one 16-byte store) or not at all.
__host__ __device__ inline int pf_xhalo(int halo) { return (halo + 3) & ~3; }
__host__ __device__ inline PfSmemPlan pf_plan(int halo, int g_range, int tz) {
    PfSmemPlan p;
    const uint32_t BY = kPfTY + 2 * halo, BZ = tz + 2 * halo, FX = kPfTX + 2 * pf_xhalo(halo);
    p.stage_bytes = (kPfBX * BY * BZ + 127u) & ~127u;
    p.stage1 = p.stage_bytes;
    p.ftile = 2 * p.stage_bytes;
    p.flags = p.ftile + FX * BY * BZ * 4u;
    p.lut = p.flags + ((BY * BZ * 4u + 15u) & ~15u);
    p.oplut = p.lut + 1024u;
    p.gw = p.oplut + 1024u;
    const uint32_t N = 2 * g_range + 1;
    p.bars = p.gw + ((N * N * N * 4u + 15u) & ~15u);
    p.total = p.bars + 16u;
    return p;
}

// NO0 / PO0: the AO tap offsets as compile-time constants (lao_column: one warp per y row, the tile's TZ outputs of a
// lane as one register-tiled column), or kPfRowWise: offsets read from A, one lao_at per voxel (any radius, and the
// launches that do not ask for AO).
constexpr int kPfRowWise = 99;
// One instantiation of the tiled kernel, as the host's dispatch table sees it.  The instantiations are spread over
// translation units (prefilter_variants.cu, one tap-offset pair each) and collected through pf_variants_<g>().
typedef void (*PfKernel)(const CUtensorMap, const PrefilterArgs);
struct PfVariant { int no0, po0, tz; PfKernel kernel; };
constexpr int kPfVariantCount = 19;                  // row-wise + 9 column instantiations x 2 tile depths (the host's dispatch table)
static_assert(kPfThreads / 32 == kPfTY, "one warp per y row of the tile");
static_assert(kPfThreads % (kPfBX / 4) == 0 && kPfLead % 4 == 0, "a thread converts the same word of every staged row it visits");

template <int NO0, int PO0, int TZ>
__global__ void __launch_bounds__(kPfThreads, TZ == kPfTZ ? 3 : 2)        // measured: 2.09 ms (one CTA of 166 registers) -> 1.41 ms at 512^3
k_prefilter_tiled(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PrefilterArgs A) {
    extern __shared__ __align__(128) unsigned char pf_smem[];
    const int h = A.halo, hx = pf_xhalo(h), BY = kPfTY + 2 * h, BZ = TZ + 2 * h, FX = kPfTX + 2 * hx;
    const PfSmemPlan P = pf_plan(h, A.g_range, TZ);
    float* ftile = reinterpret_cast<float*>(pf_smem + P.ftile);
    uint32_t* rowflag = reinterpret_cast<uint32_t*>(pf_smem + P.flags);
    float* lut = reinterpret_cast<float*>(pf_smem + P.lut);
    float* oplut = reinterpret_cast<float*>(pf_smem + P.oplut);
    float* gw = reinterpret_cast<float*>(pf_smem + P.gw);
    const uint32_t bar0 = pf_smem_u32(pf_smem + P.bars);
    const uint32_t box_bytes = (uint32_t)(kPfBX * BY * BZ);
    const uint32_t n_tiles = A.tiles_x * A.tiles_y * A.tiles_z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    auto tile_origin = [&](uint32_t t, int& x0, int& y0, int& z0) {
        x0 = (int)(t % A.tiles_x) * kPfTX;
        y0 = (int)((t / A.tiles_x) % A.tiles_y) * kPfTY;
        z0 = (int)(t / (A.tiles_x * A.tiles_y)) * TZ;
    };
    auto issue = [&](uint32_t t, uint32_t buf) {
        int x0, y0, z0;
        tile_origin(t, x0, y0, z0);
        pf_tma_load_3d(pf_smem_u32(pf_smem + buf * P.stage_bytes), &tmap, x0 - kPfLead, y0 - h, z0 - h, bar0 + 8u * buf, box_bytes);
    };

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // per-CTA tables: R8_UNORM decode, opacity of each of the 256 densities, Gaussian weights
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    if (A.opacity) oplut[tid] = opacity_of(A, (uint32_t)tid);
    if (A.gauss) {
        const int N = 2 * A.g_range + 1;
        for (int t = tid; t < N * N * N; t += kPfThreads)
            gw[t] = gauss_weight((float)(t % N - A.g_range), (float)((t / N) % N - A.g_range), (float)(t / (N * N) - A.g_range), A.g_sigma2);
    }
    for (int r = tid; r < BY * BZ; r += kPfThreads) rowflag[r] = 0u;
    // constants of empty space
    const float ao_empty = lao_at(A, [](int, int, int) -> float { return 0.0f; });
    __syncthreads();
    const float g_total = A.gauss ? gauss_total(A, gw) : 1.0f;
    // Sparse volumes (a 1024^3 hair volume is 99.5 % empty): tiles whose box (tile + halo) lies in empty occupancy cells
    // are never loaded -- they get the constants of empty space straight away (k_pf_cell_occupancy / k_pf_tile_active).
    auto is_active = [&](uint32_t t) -> bool { return !A.tile_active || A.tile_active[t] != 0; };
    // the constants of empty space for one tile
    auto fill_empty = [&](int x0, int y0, int z0) {
        float* outs[3] = {A.ao, A.opacity, A.gauss};
        const float vals[3] = {ao_empty, A.opacity ? oplut[0] : 0.0f, 0.0f};
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            float* out = outs[o];
            if (!out) continue;
            const bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0u;
            const float4 v4 = make_float4(vals[o], vals[o], vals[o], vals[o]);
            for (int q = tid; q < (kPfTX / 4) * kPfTY * TZ; q += kPfThreads) {
                const int X = x0 + 4 * (q % (kPfTX / 4)), Y = y0 + (q / (kPfTX / 4)) % kPfTY, Z = z0 + q / ((kPfTX / 4) * kPfTY);
                if (Y >= A.H || Z >= A.D || X >= A.W) continue;
                float* dst = out + ((size_t)X + (size_t)Y * A.W + (size_t)Z * A.W * A.H);
                if (vec && X + 3 < A.W) __stcs(reinterpret_cast<float4*>(dst), v4);
                else for (int e = 0; e < 4 && X + e < A.W; ++e) __stcs(dst + e, vals[o]);
            }
        }
    };
    // Tiles are drawn from a counter, two ahead (the tile being worked on, the one being loaded, the ticket in flight):
    // loaded tiles cost an order of magnitude more than tiles of empty space and come in runs, so a fixed stride leaves
    // some CTAs with all the work (half of the tiles of a 1024^3 hair volume are loaded).
    __shared__ uint32_t s_next[2];
    const bool dynamic = A.tile_counter != nullptr;
    if (dynamic && tid == 0) { s_next[0] = atomicAdd(A.tile_counter, 1u); s_next[1] = atomicAdd(A.tile_counter, 1u); }
    __syncthreads();
    uint32_t tile = dynamic ? s_next[0] : blockIdx.x, nxt = dynamic ? s_next[1] : blockIdx.x + gridDim.x, it = 0;
    auto advance = [&]() {
        __syncthreads();                                   // thread 0's ticket of this iteration is in s_next[it & 1]
        tile = nxt;
        nxt = dynamic ? s_next[it & 1u] : nxt + gridDim.x;
        ++it;
    };
    bool act_cur = tile < n_tiles && is_active(tile);
    __syncthreads();                                       // (s_next[0] has been read by everyone before it is overwritten)
    if (act_cur && tid == 0) issue(tile, 0);

    uint32_t na = 0;                                       // loaded tiles consumed so far: the next one arrives in buffer na & 1
    for (; tile < n_tiles; advance()) {
        if (dynamic && tid == 0) s_next[it & 1u] = atomicAdd(A.tile_counter, 1u);
        const bool act = act_cur;
        const bool act_nxt = nxt < n_tiles && is_active(nxt);
        act_cur = act_nxt;
        int x0, y0, z0;
        tile_origin(tile, x0, y0, z0);
        if (!act) {
            // (buffer na & 1 was last read two loaded tiles ago, with CTA barriers since)
            if (tid == 0 && act_nxt) issue(nxt, na & 1u);
            fill_empty(x0, y0, z0);
            continue;
        }
        const uint32_t buf = na & 1u;
        if (tid == 0 && act_nxt) issue(nxt, buf ^ 1u);   // prefetch the next tile
        pf_mbar_wait(bar0 + 8u * buf, (na >> 1) & 1u);
        ++na;
        const unsigned char* stage = pf_smem + buf * P.stage_bytes;

        // ---- empty space: a staged box (tile + halo) without a single hair writes the constants of empty space ----
        // (the barrier below also orders every thread's last read of this stage buffer before thread 0 re-arms it)
        {
            int nz = 0;
            const uint4* s16 = reinterpret_cast<const uint4*>(stage);
            for (int c = tid; c < (int)(box_bytes / 16u); c += kPfThreads) {
                const uint4 q = s16[c];
                nz |= (int)((q.x | q.y | q.z | q.w) != 0u);
            }
            if (!__syncthreads_or(nz)) {
                fill_empty(x0, y0, z0);
                continue;
            }
        }

        // ---- u8 rows -> float rows (+ which rows hold anything) --------------------------------------------
        // one thread per 32-bit word of a staged row: 4 texels -> 4 floats (decode table) -> one 16-byte store; the
        // word a thread handles is the same in every iteration (kPfThreads is a multiple of the 16 words of a row), so
        // threads on words outside the float tile's span sit the pass out
        int mine = 0;
        constexpr int kWords = kPfBX / 4;
        const bool need_ftile = A.ao != nullptr || A.gauss != nullptr;           // opacity reads the staged bytes directly
        {
            const int word = tid & (kWords - 1);
            const int fw = word - (kPfLead - hx) / 4;                            // word of the float row
            if (need_ftile && fw >= 0 && fw < FX / 4)
                for (int row = tid / kWords; row < BY * BZ; row += kPfThreads / kWords) {
                    const uint32_t q = *reinterpret_cast<const uint32_t*>(stage + row * kPfBX + word * 4);
                    float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (q != 0u) {
                        f = make_float4(lut[q & 0xFFu], lut[(q >> 8) & 0xFFu], lut[(q >> 16) & 0xFFu], lut[q >> 24]);
                        rowflag[row] = 1u; mine = 1;
                    }
                    *reinterpret_cast<float4*>(ftile + row * FX + fw * 4) = f;
                }
        }
        const int tile_any = __syncthreads_or(mine);

        // ---- AO, column form: warp = y row jy, lane = x, the TZ outputs along z from one pass over the planes ----
        if constexpr (NO0 != kPfRowWise) {
            const int X = x0 + lane, Y = y0 + warp;
            if (A.ao && Y < A.H) {                                            // warp-uniform
                const size_t zs = (size_t)A.W * A.H;
                float* out = A.ao + ((size_t)X + (size_t)Y * A.W + (size_t)z0 * zs);
                const bool inside = X < A.W;
#ifdef VKHR_B200_PF_UNROLLED                                         // A/B: the fully unrolled plane loop (1.11 ms at 512^3; rolled: 0.94 ms)
                lao_column<NO0, PO0, TZ>(
#else
                lao_column_rolled<NO0, PO0, TZ>(
#endif
                                     A, ftile + (h * BY + (warp + h)) * FX + (lane + hx), rowflag + h * BY + (warp + h),
                                     BY * FX, FX, BY, tile_any != 0, ao_empty, min(TZ, A.D - z0),
                                     [&](int kz, float r) { if (inside) __stcs(out + (size_t)kz * zs, r); });
            }
        }

        // ---- one output row (32 voxels along x) per warp-iteration ---------------------------------------------
        const bool row_work = A.opacity != nullptr || A.gauss != nullptr || (NO0 == kPfRowWise && A.ao != nullptr);
        for (int rr = warp; row_work && rr < kPfTY * TZ; rr += kPfThreads / 32) {
            const int jy = rr % kPfTY, kz = rr / kPfTY;
            const int X = x0 + lane, Y = y0 + jy, Z = z0 + kz;
            if (Y >= A.H || Z >= A.D) continue;                               // warp-uniform
            const size_t v = (size_t)X + (size_t)Y * A.W + (size_t)Z * A.W * A.H;
            const bool inside = X < A.W;
            const float* centre = ftile + ((kz + h) * BY + (jy + h)) * FX + (lane + hx);
            auto fetch = [&](int ox, int oy, int oz) -> float { return centre[(oz * BY + oy) * FX + ox]; };
            auto flag = [&](int oy, int oz) -> uint32_t { return rowflag[(kz + h + oz) * BY + (jy + h + oy)]; };
            if (NO0 == kPfRowWise && A.ao) {
                float r = ao_empty;
                uint32_t any = 0u;
                if (tile_any) {
                    const int o[4] = {A.neg.o0, A.neg.o1, A.pos.o0, A.pos.o1};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) any |= flag(o[a], o[b]);
                }
                if (any) r = lao_at(A, fetch);
                if (inside) __stcs(A.ao + v, r);
            }
            if (A.opacity) {
                const uint32_t d = stage[((kz + h) * BY + (jy + h)) * kPfBX + (lane + kPfLead)];
                if (inside) __stcs(A.opacity + v, oplut[d]);
            }
            if (A.gauss) {
                float r = 0.0f;                                               // 0 / total_weight
                uint32_t any = 0u;
                if (tile_any)
                    for (int b = -A.g_range; b <= A.g_range; ++b)
                        for (int a = -A.g_range; a <= A.g_range; ++a) any |= flag(a, b);
                if (any) r = A.g_range == 1 ? gauss_at_fixed<1>(gw, g_total, fetch) : A.g_range == 2 ? gauss_at_fixed<2>(gw, g_total, fetch)
                                                                                                     : gauss_at(A, gw, g_total, fetch);
                if (inside) __stcs(A.gauss + v, r);
            }
        }
        __syncthreads();                                   // everyone is done with ftile / flags / this stage buffer
        if (tile_any) for (int r = tid; r < BY * BZ; r += kPfThreads) rowflag[r] = 0u;
        __syncthreads();                                   // flags are clean before the next tile's conversion sets them
    }
}

#ifndef VKHR_PF_TILED_ONLY
// ---------------------------------------------------------------------------------------------------------
// Volumetric ADSM transmittance volume: volume_approximated_deep_shadows (share/shaders/self-shadowing/
// approximate_deep_shadows.glsl:24-36, call site volumes/volume.frag:72-78) evaluated at every voxel centre --
// the visibility of the light from that voxel through the strand density volume.  One thread per voxel marches
// the shader's own sample sequence t = 0, step, 2 step, ... (t accumulated in fp32: the table is that sequence)
// from the centre to the light; every fp32 operation is a separately rounded intrinsic in the order of the GLSL
// text (the arithmetic contract is written out in prefilter_oracle.c of the test infrastructure), so the strand
// sum is bit-identical and only the final powf may differ in the last place.
// Samples whose footprint lies outside the grid fetch the border colour (0) and add +0.0f, which leaves the sum
// unchanged: the march is clipped to the parameter range in which the ray can touch [-1, res) (plus a margin of
// two steps for the rounding of the clip itself), and every sample inside that range is still bounds-checked.
// ---------------------------------------------------------------------------------------------------------
struct AdsmArgs {
    const uint8_t* dens;
    int W, H, D;
    float ox, oy, oz, sx, sy, sz, lx, ly, lz;
    float vsx, vsy, vsz;                 // size / resolution
    float rsx, rsy, rsz;                 // RN(1 / size) for div_exact, or 0 (plain IEEE division)
    const uint32_t* occ;                 // coarse occupancy bits (k_adsm_occupancy), or nullptr
    int occ_nx32, occ_ny, occ_nz;        // words per row, rows, slices
    const float* t_table;                // the accumulated t sequence (t < 1)
    uint32_t n_t;
    float step_size, thickness, base;    // 1 / steps, thickness, 1 - strand_alpha
    float* out;
};

constexpr int kAdsmThreads = 256;

// Coarse occupancy for empty-space skipping.  A LINEAR sample whose lower texel is (x0, y0, z0), x0 in [-1, W-1],
// reads texels x0 and x0 + 1 per axis; cell c = (x0 + 1) >> 2 therefore covers texels [4c - 1, 4c + 3].  Bit
// (cx, cy, cz) is set when any texel of that 5 x 5 x 5 block is non-zero: a clear bit proves the sample is exactly
// 0.0f (all eight texels zero), and adding +0.0f leaves the strand sum unchanged, so skipping it is exact.
__global__ void __launch_bounds__(256)
k_adsm_occupancy(const uint8_t* __restrict__ dens, int W, int H, int D, int nx32, int ny, int nz, uint32_t* __restrict__ occ) {
    const uint64_t n = (uint64_t)nx32 * 32u * ny * nz;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
        const int cx = (int)(c % (uint64_t)(nx32 * 32)), cy = (int)((c / (uint64_t)(nx32 * 32)) % (uint64_t)ny), cz = (int)(c / ((uint64_t)nx32 * 32u * ny));
        uint32_t any = 0;
        for (int z = max(4 * cz - 1, 0); z <= min(4 * cz + 3, D - 1); ++z)
            for (int y = max(4 * cy - 1, 0); y <= min(4 * cy + 3, H - 1); ++y) {
                const uint8_t* row = dens + (size_t)y * W + (size_t)z * W * H;
                for (int x = max(4 * cx - 1, 0); x <= min(4 * cx + 3, W - 1); ++x) any |= __ldg(row + x);
            }
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, any != 0u);      // 32 consecutive cx = one word (n is a multiple of 32)
        if ((threadIdx.x & 31) == 0) occ[c >> 5] = bits;
    }
}

__global__ void __launch_bounds__(kAdsmThreads)
k_adsm(const __grid_constant__ AdsmArgs A) {
    __shared__ float s_tau[256];                                     // R8_UNORM decode: (float)v / 255.0f
    for (int v = threadIdx.x; v < 256; v += blockDim.x) s_tau[v] = __fdiv_rn((float)v, 255.0f);
    __syncthreads();
    const uint64_t n = (uint64_t)A.W * A.H * A.D;
    const uint64_t lin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lin >= n) return;
    const int i = (int)(lin % (uint64_t)A.W), j = (int)((lin / (uint64_t)A.W) % (uint64_t)A.H), k = (int)(lin / ((uint64_t)A.W * A.H));
    const float resx = (float)A.W, resy = (float)A.H, resz = (float)A.D;
    const float px = __fadd_rn(A.ox, __fmul_rn(__fadd_rn((float)i, 0.5f), A.vsx));
    const float py = __fadd_rn(A.oy, __fmul_rn(__fadd_rn((float)j, 0.5f), A.vsy));
    const float pz = __fadd_rn(A.oz, __fmul_rn(__fadd_rn((float)k, 0.5f), A.vsz));

    // ---- clip: parameter range in which the texel coordinate can lie in [-1, res) on every axis ----------------
    float t_lo = 0.0f, t_hi = 1.0f;
    {
        const float c0[3] = {(float)i, (float)j, (float)k};
        const float cl[3] = {(A.lx - A.ox) / A.sx * resx - 0.5f, (A.ly - A.oy) / A.sy * resy - 0.5f, (A.lz - A.oz) / A.sz * resz - 0.5f};
        const float hi[3] = {resx, resy, resz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float d = cl[a] - c0[a];
            if (fabsf(d) > 1e-20f) {
                const float ta = (-1.0f - c0[a]) / d, tb = (hi[a] - c0[a]) / d;
                t_lo = fmaxf(t_lo, fminf(ta, tb));
                t_hi = fminf(t_hi, fmaxf(ta, tb));
            }                                                        // d == 0: the centre itself is inside on this axis
        }
    }
    const float margin = 2.0f * A.step_size + 1e-5f;
    t_lo -= margin; t_hi += margin;
    // first k with t_k >= t_lo, first k with t_k > t_hi (the table is increasing)
    uint32_t k_lo = 0, k_hi = A.n_t;
    {
        uint32_t lo = 0, hi = A.n_t;
        while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(A.t_table + m) < t_lo) lo = m + 1; else hi = m; }
        k_lo = lo;
        hi = A.n_t;
        while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(A.t_table + m) <= t_hi) lo = m + 1; else hi = m; }
        k_hi = lo;
    }

    const size_t sy = (size_t)A.W, sz = (size_t)A.W * A.H;
    float strands = 0.0f;
    for (uint32_t q = k_lo; q < k_hi; ++q) {
        const float t = __ldg(A.t_table + q);
        const float omt = __fsub_rn(1.0f, t);
        // point = mix(p, light, t); u = (point - origin) / size; c = u * res - 0.5
        const float cx = __fsub_rn(__fmul_rn(div_exact(__fsub_rn(__fadd_rn(__fmul_rn(px, omt), __fmul_rn(A.lx, t)), A.ox), A.sx, A.rsx), resx), 0.5f);
        const float cy = __fsub_rn(__fmul_rn(div_exact(__fsub_rn(__fadd_rn(__fmul_rn(py, omt), __fmul_rn(A.ly, t)), A.oy), A.sy, A.rsy), resy), 0.5f);
        const float cz = __fsub_rn(__fmul_rn(div_exact(__fsub_rn(__fadd_rn(__fmul_rn(pz, omt), __fmul_rn(A.lz, t)), A.oz), A.sz, A.rsz), resz), 0.5f);
        const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
        if (!(fx0 >= -1.0f && fx0 < resx && fy0 >= -1.0f && fy0 < resy && fz0 >= -1.0f && fz0 < resz)) continue;   // all border: +0
        const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
        if (A.occ) {                                                 // empty neighbourhood: the sample is exactly +0
            const int ccx = (x0 + 1) >> 2, ccy = (y0 + 1) >> 2, ccz = (z0 + 1) >> 2;
            if (!((__ldg(A.occ + ((size_t)ccz * A.occ_ny + ccy) * A.occ_nx32 + (ccx >> 5)) >> (ccx & 31)) & 1u)) continue;
        }
        const float fx = __fsub_rn(cx, fx0), fy = __fsub_rn(cy, fy0), fz = __fsub_rn(cz, fz0);
        const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy), wz0 = __fsub_rn(1.0f, fz);
        float zz[2];
        if (x0 >= 0 && x0 + 1 < A.W && y0 >= 0 && y0 + 1 < A.H && z0 >= 0 && z0 + 1 < A.D) {
            // interior footprint (almost every sample): eight loads off one base pointer, no per-texel bounds tests
            const uint8_t* p = A.dens + ((size_t)z0 * sz + (size_t)y0 * sy + (size_t)x0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float yy[2];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint8_t* row = p + (size_t)c * sz + (size_t)b * sy;
                    yy[b] = __fadd_rn(__fmul_rn(s_tau[__ldg(row)], wx0), __fmul_rn(s_tau[__ldg(row + 1)], fx));
                }
                zz[c] = __fadd_rn(__fmul_rn(yy[0], wy0), __fmul_rn(yy[1], fy));
            }
        } else {
            const bool xin0 = x0 >= 0, xin1 = x0 + 1 < A.W;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float yy[2];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int y = y0 + b, z = z0 + c;
                    float t0 = 0.0f, t1 = 0.0f;
                    if (y >= 0 && y < A.H && z >= 0 && z < A.D) {
                        const uint8_t* row = A.dens + (size_t)y * sy + (size_t)z * sz;
                        if (xin0) t0 = s_tau[__ldg(row + x0)];
                        if (xin1) t1 = s_tau[__ldg(row + x0 + 1)];
                    }
                    yy[b] = __fadd_rn(__fmul_rn(t0, wx0), __fmul_rn(t1, fx));
                }
                zz[c] = __fadd_rn(__fmul_rn(yy[0], wy0), __fmul_rn(yy[1], fy));
            }
        }
        const float s = __fadd_rn(__fmul_rn(zz[0], wz0), __fmul_rn(zz[1], fz));
        strands = __fadd_rn(strands, __fmul_rn(s, A.thickness));
    }
    A.out[lin] = powf(A.base, strands);
}

#endif  // VKHR_PF_TILED_ONLY

}  // namespace vkhr_b200
