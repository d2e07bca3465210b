// walk.cuh -- device-side numerics of the strand walk, shared by every kernel.
//
// One thread reproduces, for one segment (or one vertex), the exact fp32
// operation sequence of the reference CPU voxeliser
//   HairStyle::voxelize_segments   src/vkhr/scene_graph/hair_style.cc:311-329
//   HairStyle::voxelize_vertices   src/vkhr/scene_graph/hair_style.cc:272-281
// so that every sample lands in the same voxel as on the CPU.  All float
// arithmetic goes through the round-to-nearest intrinsics (__fadd_rn, ...):
// they are never contracted into FMAs and never flushed, which is what a
// strict-IEEE x86-64 build of the reference computes.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace vkhr_b200 {

// Per-volume constants, computed once on the host exactly as the reference
// does (hair_style.cc:297-307): resolution as floats, voxel_size = size / res.
struct GridParams {
    float ox, oy, oz;          // bounds.origin
    float vsx, vsy, vsz;       // voxel_size
    float rx1, ry1, rz1;       // resolution - 1.0f   (hair_style.cc:320)
    float Wf, Hf;              // float(width), float(height)
    uint32_t W, H, D;
    uint32_t n_voxels;         // W*H*D  (< 2^32, checked on the host)
    uint32_t index_exact;      // VKHR_B200_INDEX_EXACT
    float rvx, rvy, rvz;       // RN(1 / voxel_size), or 0 when the fast exact division must not be used
    uint32_t fast_div;         // all three reciprocals are usable (warp-cooperative fast path)
    float pos_limit;           // fast walk only while |x|+|y|+|z| of every end point (voxel space) stays below this
    uint32_t small_grid;       // W*H*D <= 2^24: the fp32 index expression is exact, so an int32 index equals it
};

// Correctly rounded a / d for d > 0, given y = RN(1/d) (0 = not available).
//
// Markstein's sequence: q0 = RN(a*y); r0 = a - q0*d (FMA); q1 = RN(q0 + r0*y)
// is a faithful quotient, and one more exact-remainder correction
// q2 = RN(q1 + (a - q1*d)*y) is then the correctly rounded quotient RN(a/d)
// (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic,
// "division with an FMA") whenever y is the correctly rounded reciprocal and
// nothing over/underflows.  The guard keeps |a| in [2^-60, 2^60] (the host
// keeps d in [2^-40, 2^40]), zero keeps its sign (d > 0), and everything else
// -- NaN, infinities, tiny values -- takes the IEEE division instruction
// sequence.  tests/test_parity_gpu.py::test_fast_division_is_exact compares
// this bit for bit with __fdiv_rn on billions of operands.
__device__ __forceinline__ float div_exact(float a, float d, float y) {
    const float aa = fabsf(a);
    if (y != 0.0f && aa >= 8.6736174e-19f && aa <= 1.1529215e18f) {
        const float q0 = __fmul_rn(a, y);
        const float r0 = __fmaf_rn(-q0, d, a);
        const float q1 = __fmaf_rn(r0, y, q0);
        const float r1 = __fmaf_rn(-q1, d, a);
        return __fmaf_rn(r1, y, q1);
    }
    if (a == 0.0f) return a;
    return __fdiv_rn(a, d);
}

// glm::max(a,b) = (a < b) ? b : a ; glm::min(a,b) = (b < a) ? b : a
// (foreign/glm/glm/detail/func_common.inl:16-29) -- NaN behaviour included.
__device__ __forceinline__ float glm_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float glm_min(float a, float b) { return (b < a) ? b : a; }

// Voxel of a point already in voxel space: min(floor(p), res-1), then the
// linear index.  Returns false when the sample must be dropped (index NaN,
// negative or >= W*H*D: undefined behaviour in the reference).
template <int EXACT = -1>      // -1: decided at run time from g.index_exact
__device__ __forceinline__ bool voxel_index(const GridParams& g, float px, float py, float pz,
                                            uint32_t& idx) {
    float vx = glm_min(floorf(px), g.rx1);
    float vy = glm_min(floorf(py), g.ry1);
    float vz = glm_min(floorf(pz), g.rz1);
    if (EXACT == 0 || EXACT == 2 || (EXACT < 0 && !g.index_exact)) {
        // voxel.x + voxel.y*width + voxel.z*width*height, all in fp32 (hair_style.cc:276,:321)
        float a = __fmul_rn(vy, g.Wf);
        float b = __fmul_rn(vz, g.Wf);
        float c = __fmul_rn(b, g.Hf);
        float d = __fadd_rn(vx, a);
        float f = __fadd_rn(d, c);
        if (!(f >= 0.0f) || !(f < 4294967296.0f)) return false;
        idx = __float2uint_rz(f);
    } else {
        const float lim = 2147483648.0f;
        if (!(vx >= -lim && vx < lim) || !(vy >= -lim && vy < lim) || !(vz >= -lim && vz < lim)) return false;
        long long li = (long long)vx + (long long)vy * (long long)g.W +
                       (long long)vz * ((long long)g.W * (long long)g.H);
        if (li < 0 || li >= (long long)g.n_voxels) return false;
        idx = (uint32_t)li;
    }
    return idx < g.n_voxels;
}

// (v - origin) / voxel_size, per component (hair_style.cc:274, :312-313).
__device__ __forceinline__ float to_voxel_space(float v, float o, float vs, float rvs) {
    return div_exact(__fsub_rn(v, o), vs, rvs);
}

// The sampled line walk of one segment whose end points are already in voxel
// space (hair_style.cc:315-328).  `sink(idx)` is called once per sample that
// lands inside the grid.
template <int EXACT = -1, class Sink>
__device__ __forceinline__ void walk_voxel_space(const GridParams& g,
                                                 float rx, float ry, float rz,
                                                 float tx, float ty, float tz, Sink&& sink) {
    float dx = __fsub_rn(tx, rx);
    float dy = __fsub_rn(ty, ry);
    float dz = __fsub_rn(tz, rz);
    float steps = glm_max(glm_max(fabsf(dx), fabsf(dy)), fabsf(dz));     // compMax(abs(direction))
    if (!(steps > 0.0f) || !(steps < 16777216.0f)) return;               // 0 / NaN: no samples; >= 2^24: reference never ends
    // direction /= steps: three IEEE divisions by the same divisor (hair_style.cc:317)
    const float y = (steps >= 9.094947e-13f) ? __frcp_rn(steps) : 0.0f;   // RN(1/steps); tiny steps: plain division
    dx = div_exact(dx, steps, y);
    dy = div_exact(dy, steps, y);
    dz = div_exact(dz, steps, y);
    // All six finite (the only case in-AABB data produces): every later position is finite as well
    // (|root| + 2^24 * |dir|), so glm::min(floor(p), res-1) == fminf(floor(p), res-1) and the loop can
    // use the single-instruction min.  Anything else takes the literal loop below.
    if (fabsf(__fadd_rn(__fadd_rn(fabsf(rx), fabsf(ry)), fabsf(rz))) < 3.0e38f &&
        fabsf(__fadd_rn(__fadd_rn(fabsf(dx), fabsf(dy)), fabsf(dz))) < 3.0e38f) {
        // one sample; SLOT names the sink's pending-result register for this position of the unrolled loop
        auto sample = [&](auto slot) {
            const float vx = fminf(floorf(rx), g.rx1);
            const float vy = fminf(floorf(ry), g.ry1);
            const float vz = fminf(floorf(rz), g.rz1);
            uint32_t idx;
            bool ok;
            if (EXACT == 0 || EXACT == 2 || (EXACT < 0 && !g.index_exact)) {
                const float f = __fadd_rn(__fadd_rn(vx, __fmul_rn(vy, g.Wf)), __fmul_rn(__fmul_rn(vz, g.Wf), g.Hf));
                idx = __float2uint_rz(f);                                 // negative -> 0, checked through f
                ok = (f >= 0.0f) && (f < 4294967296.0f) && idx < g.n_voxels;
            } else {
                const long long li = (long long)__float2int_rz(vx) + (long long)__float2int_rz(vy) * (long long)g.W +
                                     (long long)__float2int_rz(vz) * ((long long)g.W * (long long)g.H);
                idx = (uint32_t)li;
                ok = (fabsf(vx) < 2147483648.0f) && (fabsf(vy) < 2147483648.0f) && (fabsf(vz) < 2147483648.0f) &&
                     li >= 0 && li < (long long)g.n_voxels;
            }
            if (ok) sink.template put<decltype(slot)::value>(idx);
            rx = __fadd_rn(rx, dx);
            ry = __fadd_rn(ry, dy);
            rz = __fadd_rn(rz, dz);
            steps = __fsub_rn(steps, 1.0f);
            return steps > 0.0f;
        };
        for (;;) {
            if (!sample(std::integral_constant<int, 0>{})) break;
            if (!sample(std::integral_constant<int, 1>{})) break;
            if (!sample(std::integral_constant<int, 2>{})) break;
            if (!sample(std::integral_constant<int, 3>{})) break;
        }
        return;
    }
    do {                                                                  // while (steps-- > 0.0f)
        uint32_t idx;
        if (voxel_index<EXACT>(g, rx, ry, rz, idx)) sink.template put<0>(idx);
        rx = __fadd_rn(rx, dx);
        ry = __fadd_rn(ry, dy);
        rz = __fadd_rn(rz, dz);
        steps = __fsub_rn(steps, 1.0f);
    } while (steps > 0.0f);
}

// The same from world-space vertices (hair_style.cc:312-313 first).
template <int EXACT = -1, class Sink>
__device__ __forceinline__ void walk_segment(const GridParams& g,
                                             float ax, float ay, float az,
                                             float bx, float by, float bz, Sink&& sink) {
    walk_voxel_space<EXACT>(g,
        to_voxel_space(ax, g.ox, g.vsx, g.rvx), to_voxel_space(ay, g.oy, g.vsy, g.rvy), to_voxel_space(az, g.oz, g.vsz, g.rvz),
        to_voxel_space(bx, g.ox, g.vsx, g.rvx), to_voxel_space(by, g.oy, g.vsy, g.rvy), to_voxel_space(bz, g.oz, g.vsz, g.rvz),
        sink);
}

// ---------------------------------------------------------------------------
// Warp-cooperative fast path.
//
// The per-operand range guards of div_exact cost two compares and a divergent
// branch per division, six divisions per segment.  The hot kernels instead
// evaluate the guards of all operands of a warp at once, vote, and take ONE
// warp-uniform branch: either every lane runs the branch-free FMA division
// (div_fast) and the single-instruction fminf clamp, or the whole warp runs the
// literal IEEE code above.  Both sides compute the same bits; only the cost
// differs.  All 32 lanes must call these functions together.
// ---------------------------------------------------------------------------
constexpr unsigned kFullWarp = 0xFFFFFFFFu;

// Pin a kernel-parameter value into a register.  The batch is a dynamically indexed __grid_constant__
// array, so every use of a per-instance constant is otherwise an indexed constant load (LDC c[0][R+imm]),
// re-issued inside the loops and tracked on the long scoreboard like a memory operation.
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ GridParams pin(const GridParams& c) {
    GridParams g;
    g.ox = pin(c.ox); g.oy = pin(c.oy); g.oz = pin(c.oz);
    g.vsx = pin(c.vsx); g.vsy = pin(c.vsy); g.vsz = pin(c.vsz);
    g.rx1 = pin(c.rx1); g.ry1 = pin(c.ry1); g.rz1 = pin(c.rz1);
    g.Wf = pin(c.Wf); g.Hf = pin(c.Hf);
    g.W = pin(c.W); g.H = pin(c.H); g.D = pin(c.D);        // integer index modes only (dead otherwise)
    g.n_voxels = pin(c.n_voxels);
    g.index_exact = c.index_exact;
    g.rvx = pin(c.rvx); g.rvy = pin(c.rvy); g.rvz = pin(c.rvz);
    g.fast_div = pin(c.fast_div);
    g.pos_limit = pin(c.pos_limit);
    g.small_grid = c.small_grid;
    return g;
}

// |a| in [2^-60, 2^60], or zero.  (A zero numerator gives +-0 on either path; the sign of a
// zero coordinate never reaches the voxel index: floor, min and the index sum map both to 0.)
__device__ __forceinline__ bool div_fast_ok(float a) {
    const float aa = fabsf(a);
    return (aa <= 1.1529215e18f) && (aa >= 8.6736174e-19f || aa == 0.0f);
}
// The same test for three numerators at once, in nine instructions instead of twenty-odd:
//   lower bound  w = 2 * bits - 1 drops the sign and sends +-0 to 0xFFFFFFFF, so min(w) >= 2 * bits(2^-60) - 1 says
//                "every numerator is zero or at least 2^-60" (NaN and infinities pass this half);
//   upper bound  max(|a|) <= 2^60 with the single-instruction maximum, which ignores a NaN operand -- a NaN
//                numerator therefore takes the FMA sequence, which returns NaN like the division does, and every NaN
//                is dropped alike by the walk (its range vote fails on a NaN position; the index test on a NaN index).
__device__ __forceinline__ bool div_fast_ok3(float a, float b, float c) {
    const uint32_t wa = 2u * __float_as_uint(a) - 1u, wb = 2u * __float_as_uint(b) - 1u, wc = 2u * __float_as_uint(c) - 1u;
    const float hi = fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(c));
    return min(min(wa, wb), wc) >= 2u * 0x21800000u - 1u && hi <= 1.1529215e18f;
}
// The lower half alone, for numerators whose magnitude is already known to be below 2^60.
__device__ __forceinline__ bool div_fast_ok3_lower(float a, float b, float c) {
    const uint32_t wa = 2u * __float_as_uint(a) - 1u, wb = 2u * __float_as_uint(b) - 1u, wc = 2u * __float_as_uint(c) - 1u;
    return min(min(wa, wb), wc) >= 2u * 0x21800000u - 1u;
}
// The Markstein sequence of div_exact without its guard.
__device__ __forceinline__ float div_fast(float a, float d, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r0 = __fmaf_rn(-q0, d, a);
    const float q1 = __fmaf_rn(r0, y, q0);
    const float r1 = __fmaf_rn(-q1, d, a);
    return __fmaf_rn(r1, y, q1);
}
// RN(1 / s) for a normal s well inside the exponent range (the walk calls it for steps in [2^-40, 2^24)): the
// hardware approximation and one Newton step -- the in-range path of __frcp_rn without its range test and branch.
// harness/selftest.cu compares it with __frcp_rn for EVERY float of that interval.
__device__ __forceinline__ float rcp_steps(float s) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    const float e = __fmaf_rn(s, r, -1.0f);
    return __fmaf_rn(r, -e, r);
}

// (v - origin) / voxel_size for one vertex per lane (hair_style.cc:274, :312-313).
__device__ __forceinline__ void to_voxel_space_warp(const GridParams& g, float wx, float wy, float wz,
                                                    float& px, float& py, float& pz) {
    const float ax = __fsub_rn(wx, g.ox), ay = __fsub_rn(wy, g.oy), az = __fsub_rn(wz, g.oz);
    if (g.fast_div && __all_sync(kFullWarp, div_fast_ok3(ax, ay, az))) {
        px = div_fast(ax, g.vsx, g.rvx);
        py = div_fast(ay, g.vsy, g.rvy);
        pz = div_fast(az, g.vsz, g.rvz);
    } else {
        px = __fdiv_rn(ax, g.vsx);
        py = __fdiv_rn(ay, g.vsy);
        pz = __fdiv_rn(az, g.vsz);
    }
}

// One sample of the walk for finite positions: voxel, fp32 (or exact) linear index, range test.
template <int EXACT>
__device__ __forceinline__ bool sample_index(const GridParams& g, float rx, float ry, float rz, uint32_t& idx) {
    if (EXACT == 2) {
        // Small grids (W*H*D <= 2^24) with bounded positions (|p| < pos_limit, voted by the caller): every
        // term of the fp32 expression is an exactly representable integer, so the int32 expression is the
        // same number; a negative index (the fp32 path's f < 0) wraps above n_voxels and is dropped.
        const int ix = min(__float2int_rd(rx), (int)g.W - 1);
        const int iy = min(__float2int_rd(ry), (int)g.H - 1);
        const int iz = min(__float2int_rd(rz), (int)g.D - 1);
        idx = (uint32_t)((iz * (int)g.H + iy) * (int)g.W + ix);
        return idx < g.n_voxels;
    }
    const float vx = fminf(floorf(rx), g.rx1);
    const float vy = fminf(floorf(ry), g.ry1);
    const float vz = fminf(floorf(rz), g.rz1);
    if (EXACT == 0 || (EXACT < 0 && !g.index_exact)) {
        const float f = __fadd_rn(__fadd_rn(vx, __fmul_rn(vy, g.Wf)), __fmul_rn(__fmul_rn(vz, g.Wf), g.Hf));
        idx = __float2uint_rz(f);                       // NaN and negatives -> 0 (caught by f >= 0), >= 2^32 -> 0xFFFFFFFF
        return (f >= 0.0f) && idx < g.n_voxels;
    } else {
        const long long li = (long long)__float2int_rz(vx) + (long long)__float2int_rz(vy) * (long long)g.W +
                             (long long)__float2int_rz(vz) * ((long long)g.W * (long long)g.H);
        idx = (uint32_t)li;
        return (fabsf(vx) < 2147483648.0f) && (fabsf(vy) < 2147483648.0f) && (fabsf(vz) < 2147483648.0f) &&
               li >= 0 && li < (long long)g.n_voxels;
    }
}

// Brick layout of the BRICK8 scratch volume: a brick is 4 x 4 x 2 voxels (x, y, z) = eight 32-bit words = one 32-byte
// sector; word ((z & 1) << 2) | (y & 3) of the brick holds the four voxels x & ~3 .. x | 3 of that row -- the same
// four voxels, in the same byte order, as word idx >> 2 of the x-fastest volume (W % 4 == 0).
__device__ __forceinline__ uint32_t brick_word(const GridParams& g, int ix, int iy, int iz) {
    const uint32_t brick = ((uint32_t)(iz >> 1) * (g.H >> 2) + (uint32_t)(iy >> 2)) * (g.W >> 2) + (uint32_t)(ix >> 2);
    return (brick << 3) | ((uint32_t)(iz & 1) << 2) | (uint32_t)(iy & 3);
}

// The brick word of a LINEAR index when W and H are powers of two (lw = log2 W, lh = log2 H): bit fields only.
__device__ __forceinline__ uint32_t brick_word_pow2(uint32_t lin, uint32_t lw, uint32_t lh) {
    const uint32_t x = lin & ((1u << lw) - 1u), y = (lin >> lw) & ((1u << lh) - 1u), z = lin >> (lw + lh);
    const uint32_t brick = ((((z >> 1) << (lh - 2u)) | (y >> 2)) << (lw - 2u)) | (x >> 2);
    return (brick << 3) | ((z & 1u) << 2) | (y & 3u);
}

// The sampled walk of one segment per lane, end points in voxel space (hair_style.cc:315-328).
// EXACT == 3: the int32 index of EXACT == 2 for a sink that counts in the brick layout (put_brick).
// EXACT == 4: the reference's fp32 index (it rounds above 2^24 voxels, hair_style.cc:321) for such a sink on a grid
//             whose W and H are powers of two: the ROUNDED index is what names the voxel, so the brick word is taken
//             from its bit fields (sink.put_linear).
// `active` = this lane has a segment.
template <int EXACT, bool CHECK_TIP = true, class Sink>
__device__ __forceinline__ void walk_voxel_space_warp(const GridParams& g, bool active,
                                                      float rx, float ry, float rz,
                                                      float tx, float ty, float tz, Sink& sink) {
    float dx = __fsub_rn(tx, rx), dy = __fsub_rn(ty, ry), dz = __fsub_rn(tz, rz);
    float steps = glm_max(glm_max(fabsf(dx), fabsf(dy)), fabsf(dz));          // compMax(abs(direction))
    const bool go = active && (steps > 0.0f) && (steps < 16777216.0f);        // 0 / NaN: no samples; >= 2^24: never ends
    // fast when the root is finite (then every later position is finite too: |root| + 2^24 |dir|, so
    // glm::min == fminf) and the three divisions by `steps` are inside the FMA division's range
    // (EXACT == 2 needs the bound on EVERY end point the warp holds, idle lanes included: lane t's tip is lane
    // t+1's root in the uniform kernel; CHECK_TIP adds the tip for kernels whose tips are not another lane's root).
    // |direction| <= steps < 2^24, so only the lower bound of the numerators is left to test.
    bool bounded = __fadd_rn(__fadd_rn(fabsf(rx), fabsf(ry)), fabsf(rz)) < g.pos_limit;
    if (CHECK_TIP && EXACT >= 2) bounded = bounded && (__fadd_rn(__fadd_rn(fabsf(tx), fabsf(ty)), fabsf(tz)) < g.pos_limit);
    const bool fast = (EXACT >= 2 || go ? bounded : true) &&
                      (!go || ((steps >= 9.094947e-13f) && div_fast_ok3_lower(dx, dy, dz)));
    if (__all_sync(kFullWarp, fast)) {
        if (go) {
            const float y = rcp_steps(steps);                                 // RN(1/steps)
            dx = div_fast(dx, steps, y);
            dy = div_fast(dy, steps, y);
            dz = div_fast(dz, steps, y);
            // while (steps-- > 0.0f); unrolled by four so that a sink may keep one pending result per position
            auto sample = [&](auto slot) {
                if constexpr (EXACT == 3) {
                    const int ix = min(__float2int_rd(rx), (int)g.W - 1);
                    const int iy = min(__float2int_rd(ry), (int)g.H - 1);
                    const int iz = min(__float2int_rd(rz), (int)g.D - 1);
                    if ((ix | iy | iz) >= 0) sink.template put_xyz<decltype(slot)::value>(ix, iy, iz);   // inside the grid
                    else {                                                   // a negative coordinate that still indexes a voxel
                        const uint32_t lin = (uint32_t)((iz * (int)g.H + iy) * (int)g.W + ix);   // the reference's index, as EXACT == 2
                        if (lin < g.n_voxels) sink.template put<decltype(slot)::value>(lin);
                    }
                } else if constexpr (EXACT == 4) {
                    uint32_t idx;
                    if (sample_index<0>(g, rx, ry, rz, idx)) sink.template put_linear<decltype(slot)::value>(idx);
                } else {
                    uint32_t idx;
                    if (sample_index<EXACT>(g, rx, ry, rz, idx)) sink.template put<decltype(slot)::value>(idx);
                }
                rx = __fadd_rn(rx, dx);
                ry = __fadd_rn(ry, dy);
                rz = __fadd_rn(rz, dz);
                steps = __fsub_rn(steps, 1.0f);
                return steps > 0.0f;
            };
            for (;;) {
                if (!sample(std::integral_constant<int, 0>{})) break;
                if (!sample(std::integral_constant<int, 1>{})) break;
                if (!sample(std::integral_constant<int, 2>{})) break;
                if (!sample(std::integral_constant<int, 3>{})) break;
            }
        }
    } else if (go) {
        walk_voxel_space<(EXACT == 3 ? 2 : EXACT == 4 ? 0 : EXACT)>(g, rx, ry, rz, tx, ty, tz, sink);   // the literal code, any input
    }
}

// ---------------------------------------------------------------------------
// The interior walk (small grids, EXACT == 3 sinks): a tile whose 32 vertices ALL lie in [1, res - 1] on every axis.
//
// Then, for a segment of fewer than 2048 steps between two such vertices,
//   * every sample lies in [0, res): the exact point root + k * dir is between the end points, and the accumulated
//     `root += dir` is at most k half-ulps of a position below 8192 away from it (k * 2^-11 < 1), so neither the
//     reference's upper clamp min(floor(p), res - 1) nor a negative coordinate can occur -- the loop needs no clamp and
//     no sign test, and every sample is inside the grid (put_xyz);
//   * the FMA division needs no range test: the numerators of (v - origin) / voxel_size were validated by the result
//     (a numerator outside [2^-60, 2^60] gives a quotient above 2^20 or below 2^-20, which is not in [1, res - 1];
//     NaN fails the comparison), and `direction / steps` divides differences of floats that are at least 1 by their
//     largest magnitude: a component is zero or at least 2^-23, and so is `steps`.
// The caller votes on the condition (vertex_is_interior on every lane's own vertex, idle lanes included, and
// steps < 2048 on every lane with a segment) and runs the guarded walk_voxel_space_warp when it fails.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool vertex_is_interior(const GridParams& g, float px, float py, float pz) {
    return fminf(fminf(px, py), pz) >= 1.0f && px <= g.rx1 && py <= g.ry1 && pz <= g.rz1;
}
template <class Sink>
__device__ __forceinline__ void walk_interior_lane(float rx, float ry, float rz, float dx, float dy, float dz, float steps, Sink& sink) {
    const float y = rcp_steps(steps);                                         // RN(1/steps)
    dx = div_fast(dx, steps, y);
    dy = div_fast(dy, steps, y);
    dz = div_fast(dz, steps, y);
    for (;;) {                                                                // while (steps-- > 0.0f)
        sink.template put_xyz<0>(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
        steps = __fsub_rn(steps, 1.0f);
        if (!(steps > 0.0f)) break;
        rx = __fadd_rn(rx, dx);
        ry = __fadd_rn(ry, dy);
        rz = __fadd_rn(rz, dz);
    }
}

// The interior walk of a whole tile with the samples DEALT OUT AGAIN before they are added (all 32 lanes call it).
//
// What a sample costs is its 32-byte request packet on the SM -> L2 path, and one warp instruction sends one packet per
// distinct sector its lanes touch.  With one segment per lane, instruction k adds sample k of 31
// consecutive segments -- points spread along two and a half strands: 0.57 sectors per sample in the brick layout.
// Here every lane first computes the packed address (sink.pack_xyz) of its segment's first TWO samples; then two
// instructions add them, each taking BOTH samples of 16 consecutive segments (lane j: sample j & 1 of segment
// base + j / 2, fetched by shuffle): the same points, but an instruction's lanes now lie along one strand's length --
// 0.34 sectors per sample on the bench's strands (tools/locality_stats.py), 40 % fewer packets.  Samples beyond the
// second (segments longer than two voxels) are added by their own lane afterwards, as walk_interior_lane does.
// The positions are the reference's: root, root + dir, (root + dir) + dir, ... accumulated in that order; `steps - k > 0`
// with the exact float decrements of `while (steps-- > 0.0f)` (hair_style.cc:318) is `steps > k` for k = 1, 2.
template <class Sink>
__device__ __forceinline__ void walk_interior_tile_dealt(bool go, float rx, float ry, float rz, float dx, float dy, float dz,
                                                         float steps, Sink& sink) {
    constexpr uint32_t kNone = 0xFFFFFFFFu;                                   // (a packed sample is below 2^24)
    uint32_t a0 = kNone, a1 = kNone;
    bool more = false;
    if (go) {
        const float y = rcp_steps(steps);                                     // RN(1/steps)
        dx = div_fast(dx, steps, y);
        dy = div_fast(dy, steps, y);
        dz = div_fast(dz, steps, y);
        a0 = sink.pack_xyz(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
        if (steps > 1.0f) {
            rx = __fadd_rn(rx, dx);
            ry = __fadd_rn(ry, dy);
            rz = __fadd_rn(rz, dz);
            a1 = sink.pack_xyz(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
            more = steps > 2.0f;
        }
    }
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u, half = lane >> 1;
    const bool second = (lane & 1u) != 0u;
#pragma unroll
    for (uint32_t base = 0; base < 32u; base += 16u) {
        const uint32_t v0 = __shfl_sync(kFullWarp, a0, base + half), v1 = __shfl_sync(kFullWarp, a1, base + half);
        const uint32_t a = second ? v1 : v0;
        if (a != kNone) sink.put_packed(a);
    }
    if (__any_sync(kFullWarp, more)) {
        if (more) {
            steps = __fsub_rn(__fsub_rn(steps, 1.0f), 1.0f);
            do {
                rx = __fadd_rn(rx, dx);
                ry = __fadd_rn(ry, dy);
                rz = __fadd_rn(rz, dz);
                sink.template put_xyz<0>(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
                steps = __fsub_rn(steps, 1.0f);
            } while (steps > 0.0f);
        }
        __syncwarp();
    }
}

// The interior walk of a whole tile with NEIGHBOUR ABSORPTION (all 32 lanes call it).
//
// Lanes that add to the same 32-bit word in one warp instruction are serialised into separate packets, lanes that add
// to different words of one sector share a packet (measured: profiles/README, round 2).  The second sample of segment
// t and the first sample of segment t + 1 lie a quarter of a voxel apart -- the same word six times out of ten on the
// bench's strands -- but sit in different instructions (sample 1 of lane t, sample 0 of lane t + 1) and cost a packet
// each.  Here lane t + 1 adds its neighbour's second sample to its own first one (one red of 1 << 8 b0 + 1 << 8 b1)
// whenever the words agree, and lane t drops it: 0.57 -> 0.46 packets per sample.  The byte-sum verdict is unaffected:
// a red still raises the volume's byte sum by the number of samples it carries unless a byte carries out.
template <class Sink>
__device__ __forceinline__ void walk_interior_tile_absorb(bool go, float rx, float ry, float rz, float dx, float dy, float dz,
                                                          float steps, Sink& sink) {
    constexpr uint32_t kNone = 0xFFFFFFFFu;                                   // (a packed sample is below 2^24)
    uint32_t a0 = kNone, a1 = kNone;
    bool more = false;
    if (go) {
        const float y = rcp_steps(steps);                                     // RN(1/steps)
        dx = div_fast(dx, steps, y);
        dy = div_fast(dy, steps, y);
        dz = div_fast(dz, steps, y);
        a0 = sink.pack_xyz(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
        if (steps > 1.0f) {
            rx = __fadd_rn(rx, dx);
            ry = __fadd_rn(ry, dy);
            rz = __fadd_rn(rz, dz);
            a1 = sink.pack_xyz(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
            more = steps > 2.0f;
        }
    }
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nxt0 = __shfl_down_sync(kFullWarp, a0, 1), prv1 = __shfl_up_sync(kFullWarp, a1, 1);
    // the pair (second sample of lane t, first sample of lane t + 1): both lanes evaluate the same test
    const bool give = lane != 31u && (a1 ^ nxt0) < 4u && (a1 | nxt0) < 0x80000000u;    // same word, both present
    const bool take = lane != 0u && (prv1 ^ a0) < 4u && (prv1 | a0) < 0x80000000u;
    if (a0 != kNone) sink.put_packed_value(a0, (1u << ((a0 & 3u) * 8u)) + (take ? 1u << ((prv1 & 3u) * 8u) : 0u), 1u);
    if (a1 != kNone) { if (give) sink.count(1u); else sink.put_packed(a1); }
    if (__any_sync(kFullWarp, more)) {
        if (more) {
            steps = __fsub_rn(__fsub_rn(steps, 1.0f), 1.0f);
            do {
                rx = __fadd_rn(rx, dx);
                ry = __fadd_rn(ry, dy);
                rz = __fadd_rn(rz, dz);
                sink.template put_xyz<0>(__float2int_rd(rx), __float2int_rd(ry), __float2int_rd(rz));
                steps = __fsub_rn(steps, 1.0f);
            } while (steps > 0.0f);
        }
        __syncwarp();
    }
}

// Vertex pair of segment `s`.  indices == nullptr => uniform strands of
// `segs` segments: the pairs HairStyle::generate_indices (hair_style.cc:196-213)
// would emit, without reading an index buffer.
__device__ __forceinline__ void segment_vertices(const uint32_t* __restrict__ indices, uint32_t segs,
                                                 uint64_t s, uint32_t& i0, uint32_t& i1) {
    if (indices) {
        uint2 p = __ldg(reinterpret_cast<const uint2*>(indices) + s);
        i0 = p.x; i1 = p.y;
    } else {
        uint32_t strand = (uint32_t)(s / segs);
        i0 = (uint32_t)s + strand;
        i1 = i0 + 1;
    }
}

}  // namespace vkhr_b200
