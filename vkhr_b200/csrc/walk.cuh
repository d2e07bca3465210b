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
};

// glm::max(a,b) = (a < b) ? b : a ; glm::min(a,b) = (b < a) ? b : a
// (foreign/glm/glm/detail/func_common.inl:16-29) -- NaN behaviour included.
__device__ __forceinline__ float glm_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float glm_min(float a, float b) { return (b < a) ? b : a; }

// Voxel of a point already in voxel space: min(floor(p), res-1), then the
// linear index.  Returns false when the sample must be dropped (index NaN,
// negative or >= W*H*D: undefined behaviour in the reference).
__device__ __forceinline__ bool voxel_index(const GridParams& g, float px, float py, float pz,
                                            uint32_t& idx) {
    float vx = glm_min(floorf(px), g.rx1);
    float vy = glm_min(floorf(py), g.ry1);
    float vz = glm_min(floorf(pz), g.rz1);
    if (!g.index_exact) {
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
__device__ __forceinline__ float to_voxel_space(float v, float o, float vs) {
    return __fdiv_rn(__fsub_rn(v, o), vs);
}

// The sampled line walk of one segment (hair_style.cc:312-328).  `sink(idx)` is
// called once per sample that lands inside the grid.
template <class Sink>
__device__ __forceinline__ void walk_segment(const GridParams& g,
                                             float ax, float ay, float az,
                                             float bx, float by, float bz, Sink&& sink) {
    float rx = to_voxel_space(ax, g.ox, g.vsx);
    float ry = to_voxel_space(ay, g.oy, g.vsy);
    float rz = to_voxel_space(az, g.oz, g.vsz);
    float dx = __fsub_rn(to_voxel_space(bx, g.ox, g.vsx), rx);
    float dy = __fsub_rn(to_voxel_space(by, g.oy, g.vsy), ry);
    float dz = __fsub_rn(to_voxel_space(bz, g.oz, g.vsz), rz);
    float steps = glm_max(glm_max(fabsf(dx), fabsf(dy)), fabsf(dz));     // compMax(abs(direction))
    if (!(steps > 0.0f) || !(steps < 16777216.0f)) return;               // 0 / NaN: no samples; >= 2^24: reference never ends
    dx = __fdiv_rn(dx, steps);
    dy = __fdiv_rn(dy, steps);
    dz = __fdiv_rn(dz, steps);
    do {                                                                  // while (steps-- > 0.0f)
        uint32_t idx;
        if (voxel_index(g, rx, ry, rz, idx)) sink(idx);
        rx = __fadd_rn(rx, dx);
        ry = __fadd_rn(ry, dy);
        rz = __fadd_rn(rz, dz);
        steps = __fsub_rn(steps, 1.0f);
    } while (steps > 0.0f);
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
