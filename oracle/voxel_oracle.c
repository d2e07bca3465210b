/* oracle/voxel_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's strand-voxelisation path
 * (CaffeineViking/vkhr @ 6d26f5b, src/vkhr/scene_graph/hair_style.cc and
 * include/vkhr/scene_graph/hair_style.hh; GLM 0.9.9.2 arithmetic).  Every
 * function cites the reference lines it follows.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check / time
 * the CUDA path; nothing in vkhr_b200/ may call it.
 *
 * PINNING.  The reference ships no tests or golden vectors for this path
 * (SURVEY.md F11).  This restatement is pinned instead against outputs of the
 * reference itself: oracle/_ref/libvkhr_ref.so (the unmodified hair_style.cc,
 * built by oracle/Makefile) on the hand-verified 4^3 known-answer cases of
 * SURVEY.md Appendix B and on seeded bulk sets; the resulting vectors are
 * committed under tests/golden/ (generator: tests/golden/make_golden.py) and
 * re-checked by tests/test_oracle.py on every run.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off (see oracle/Makefile); the
 * arithmetic below is IEEE-754 binary32, round-to-nearest-even, one rounding
 * per written operation.
 *
 * Where the reference is undefined we fix a rule (the same one the CUDA path
 * implements), and keep those inputs out of the reference-vs-oracle checks:
 *   - a sample whose fp32 linear index is NaN, negative or >= W*H*D (the
 *     reference indexes out of bounds) is dropped;
 *   - a segment whose step count is not < 2^24 (the reference's
 *     `while (steps-- > 0)` never terminates) is skipped;
 *   - normalize() with max == min (division by zero) leaves the grid as is;
 *   - fewer than two indices (size()-1 underflow) gives an empty volume.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* glm::min(a,b) = (b < a) ? b : a   -- glm/detail/func_common.inl:16-20 */
static inline float glm_min(float a, float b) { return (b < a) ? b : a; }
/* glm::max(a,b) = (a < b) ? b : a   -- glm/detail/func_common.inl:24-29 */
static inline float glm_max(float a, float b) { return (a < b) ? b : a; }

/* HairStyle::generate_bounding_box, hair_style.cc:215-234.  min/max start at
 * (0,0,0), so the box always contains the world origin. */
void oracle_generate_bounding_box(const float* xyz, uint64_t n_vertices,
                                  float aabb_min[3], float aabb_max[3]) {
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {0.0f, 0.0f, 0.0f};
    for (uint64_t i = 0; i < n_vertices; ++i)
        for (int c = 0; c < 3; ++c) {
            float p = xyz[3 * i + c];
            lo[c] = glm_min(p, lo[c]);
            hi[c] = glm_max(p, hi[c]);
        }
    memcpy(aabb_min, lo, 12);
    memcpy(aabb_max, hi, 12);
}

/* HairStyle::get_bounding_box, hair_style.cc:236-255: origin = min,
 * size = max - min, radius = |size|, volume = sx*sy*sz.
 * out8 = origin.xyz, radius, size.xyz, volume (struct AABB, hair_style.hh:16-21). */
void oracle_get_bounding_box(const float aabb_min[3], const float aabb_max[3], float out8[8]) {
    float sx = aabb_max[0] - aabb_min[0];
    float sy = aabb_max[1] - aabb_min[1];
    float sz = aabb_max[2] - aabb_min[2];
    out8[0] = aabb_min[0]; out8[1] = aabb_min[1]; out8[2] = aabb_min[2];
    out8[3] = sqrtf((sx * sx + sy * sy) + sz * sz);      /* glm::length = sqrt(dot) */
    out8[4] = sx; out8[5] = sy; out8[6] = sz;
    out8[7] = (sx * sy) * sz;
}

/* HairStyle::generate_indices, hair_style.cc:196-213: pairs (k,k+1) within a
 * strand, skipping the strand's last vertex.  segments == NULL => every strand
 * has default_segments.  Returns the number of indices written. */
uint64_t oracle_generate_indices(uint32_t n_strands, uint32_t default_segments,
                                 const uint16_t* segments, uint32_t* indices_out) {
    uint64_t n = 0;
    uint32_t vertex = 0;
    for (uint32_t s = 0; s < n_strands; ++s) {
        uint32_t count = segments ? segments[s] : default_segments;
        for (uint32_t k = 0; k < count; ++k) {
            indices_out[n++] = vertex++;
            indices_out[n++] = vertex;
        }
        ++vertex;
    }
    return n;
}

/* HairStyle::generate_tangents, hair_style.cc:171-194: normalize(v[k+1]-v[k])
 * (glm::normalize = v * (1/sqrt(dot(v,v)))); a strand's last vertex repeats
 * the previous tangent. */
void oracle_generate_tangents(const float* xyz, uint32_t n_strands, uint32_t default_segments,
                              const uint16_t* segments, float* tangents_out) {
    uint64_t vertex = 0;
    for (uint32_t s = 0; s < n_strands; ++s) {
        uint32_t count = segments ? segments[s] : default_segments;
        for (uint32_t k = 0; k < count; ++k) {
            float tx = xyz[3 * (vertex + 1) + 0] - xyz[3 * vertex + 0];
            float ty = xyz[3 * (vertex + 1) + 1] - xyz[3 * vertex + 1];
            float tz = xyz[3 * (vertex + 1) + 2] - xyz[3 * vertex + 2];
            float d = (tx * tx + ty * ty) + tz * tz;
            float inv = 1.0f / sqrtf(d);
            tangents_out[3 * vertex + 0] = tx * inv;
            tangents_out[3 * vertex + 1] = ty * inv;
            tangents_out[3 * vertex + 2] = tz * inv;
            ++vertex;
        }
        if (vertex > 0) memcpy(&tangents_out[3 * vertex], &tangents_out[3 * (vertex - 1)], 12);
        ++vertex;
    }
}

/* fp32 linear index of hair_style.cc:276 / :321:
 *   voxel.x + voxel.y*width + voxel.z*width*height
 * with `voxel` a glm::vec3 and width/height size_t => every operand is
 * converted to float and every operation rounds to binary32. */
static inline float index_f32(float vx, float vy, float vz, float Wf, float Hf) {
    float a = vy * Wf;
    float b = vz * Wf;
    float c = b * Hf;
    float d = vx + a;
    return d + c;
}

typedef struct {
    float origin[3], voxel_size[3], res_m1[3], Wf, Hf;
    uint64_t W, H, n_voxels;
    int index_exact;          /* flags bit 0: OUR extension, not the reference's behaviour */
} grid_t;

#define ORACLE_INDEX_EXACT 1u

static void grid_init(grid_t* g, const float origin[3], const float size[3],
                      uint64_t W, uint64_t H, uint64_t D, uint32_t flags) {
    g->W = W; g->H = H;
    g->index_exact = (flags & ORACLE_INDEX_EXACT) != 0;
    float res[3] = {(float)W, (float)H, (float)D};       /* hair_style.cc:297-304 */
    for (int c = 0; c < 3; ++c) {
        g->origin[c] = origin[c];
        g->voxel_size[c] = size[c] / res[c];              /* :307 */
        g->res_m1[c] = res[c] - 1.0f;                     /* resolution - 1.0f, :320 */
    }
    g->Wf = res[0];
    g->Hf = res[1];
    g->n_voxels = W * H * D;
}

/* One voxel hit.  counts == NULL: reference semantics (saturating u8 + first-255
 * tangent accumulation, hair_style.cc:322-325).  counts != NULL: plain u32
 * hit count (what a multi-GPU partial holds before the clamp). */
static inline void hit(const grid_t* g, float vx, float vy, float vz, uint8_t* dens, uint32_t* counts,
                       float* tsum, const float* tangent) {
    uint64_t idx;
    if (!g->index_exact) {
        float idx_f = index_f32(vx, vy, vz, g->Wf, g->Hf);                  /* :276 / :321 */
        /* `int voxel_index = <float>` (:321) / `size_t pos = <float>` (:276):
         * NaN, negative or >= 2^31 is UB there; >= n_voxels indexes out of
         * bounds.  All of those are dropped. */
        if (!(idx_f >= 0.0f) || !(idx_f < 4294967296.0f)) return;
        idx = (uint64_t)idx_f;
    } else {
        /* exact-integer extension: x + y*W + z*W*H in 64-bit integers */
        if (!(vx >= -2147483648.0f && vx < 2147483648.0f)) return;
        if (!(vy >= -2147483648.0f && vy < 2147483648.0f)) return;
        if (!(vz >= -2147483648.0f && vz < 2147483648.0f)) return;
        int64_t li = (int64_t)vx + (int64_t)vy * (int64_t)g->W + (int64_t)vz * (int64_t)(g->W * g->H);
        if (li < 0) return;
        idx = (uint64_t)li;
    }
    if (idx >= g->n_voxels) return;
    if (counts) { counts[idx] += 1; return; }
    if (dens[idx] != 255) {
        if (tsum) {
            tsum[3 * idx + 0] += tangent[0];
            tsum[3 * idx + 1] += tangent[1];
            tsum[3 * idx + 2] += tangent[2];
        }
        dens[idx] += 1;
    }
}

/* hair_style.cc:331-339 (and :283-291): tangent_sum / float(density) * 127.0f
 * truncated to int8; 0/0 = NaN converts to 0 on x86 (cvttss2si -> INT_MIN,
 * low byte 0). */
static void quantise(const grid_t* g, const uint8_t* dens, const float* tsum, int8_t* tangents_out) {
    for (uint64_t i = 0; i < g->n_voxels; ++i) {
        float dn = (float)dens[i];
        for (int c = 0; c < 3; ++c) {
            float q = tsum[3 * i + c] / dn * 127.0f;
            int32_t qi = (q != q) ? INT32_MIN : (int32_t)q;
            tangents_out[4 * i + c] = (int8_t)qi;
        }
        tangents_out[4 * i + 3] = 0;
    }
}

/* The sampled line walk of HairStyle::voxelize_segments, hair_style.cc:311-329. */
static void walk_segments(const grid_t* g, const float* xyz, const uint32_t* indices,
                          uint64_t n_indices, const float* tangents,
                          uint8_t* dens, uint32_t* counts, float* tsum) {
    if (n_indices < 2) return;
    for (uint64_t i = 0; i < n_indices - 1; i += 2) {                       /* :311 */
        const float* a = &xyz[3 * (uint64_t)indices[i]];
        const float* b = &xyz[3 * (uint64_t)indices[i + 1]];
        float root[3], dir[3];
        for (int c = 0; c < 3; ++c) {
            float r = (a[c] - g->origin[c]) / g->voxel_size[c];            /* :312 */
            float t = (b[c] - g->origin[c]) / g->voxel_size[c];            /* :313 */
            root[c] = r;
            dir[c] = t - r;                                                 /* :315 */
        }
        float steps = glm_max(glm_max(fabsf(dir[0]), fabsf(dir[1])), fabsf(dir[2]));   /* :316 compMax(abs) */
        dir[0] /= steps; dir[1] /= steps; dir[2] /= steps;                  /* :317 */
        if (!(steps < 16777216.0f)) continue;                               /* reference would not terminate */
        const float* tangent = tangents ? &tangents[3 * (uint64_t)indices[i]] : NULL;
        while (steps-- > 0.0f) {                                            /* :319 */
            float vx = glm_min(floorf(root[0]), g->res_m1[0]);             /* :320 */
            float vy = glm_min(floorf(root[1]), g->res_m1[1]);
            float vz = glm_min(floorf(root[2]), g->res_m1[2]);
            hit(g, vx, vy, vz, dens, counts, tsum, tangent);                /* :321-325 */
            root[0] += dir[0]; root[1] += dir[1]; root[2] += dir[2];        /* :327 */
        }
    }
}

/* HairStyle::voxelize_vertices inner loop, hair_style.cc:272-281. */
static void walk_vertices(const grid_t* g, const float* xyz, uint64_t n_vertices,
                          const float* tangents, uint8_t* dens, uint32_t* counts, float* tsum) {
    for (uint64_t i = 0; i < n_vertices; ++i) {
        float v[3];
        for (int c = 0; c < 3; ++c) {
            float p = (xyz[3 * i + c] - g->origin[c]) / g->voxel_size[c];  /* :274 */
            v[c] = glm_min(floorf(p), g->res_m1[c]);                        /* :275 */
        }
        hit(g, v[0], v[1], v[2], dens, counts, tsum, tangents ? &tangents[3 * i] : NULL);   /* :276-280 */
    }
}

/* HairStyle::voxelize_segments, hair_style.cc:296-342.
 * densities_out: W*H*D u8, x fastest then y then z (zeroed here).
 * tangents_in / tangents_out may both be NULL for the density-only walk (no
 * 12*W*H*D temporary); otherwise tangents_out is W*H*D x int8[4]. */
int oracle_voxelize_segments(const float* xyz, const uint32_t* indices, uint64_t n_indices,
                             const float* tangents_in, const float origin[3], const float size[3],
                             uint64_t W, uint64_t H, uint64_t D, uint32_t flags,
                             uint8_t* densities_out, int8_t* tangents_out) {
    grid_t g;
    grid_init(&g, origin, size, W, H, D, flags);
    memset(densities_out, 0, g.n_voxels);
    float* tsum = NULL;
    if (tangents_in && tangents_out) {
        tsum = (float*)calloc(g.n_voxels * 3, sizeof(float));               /* precise_tangents, :309 */
        if (!tsum) return -1;
    }
    walk_segments(&g, xyz, indices, n_indices, tsum ? tangents_in : NULL, densities_out, NULL, tsum);
    if (tsum) { quantise(&g, densities_out, tsum, tangents_out); free(tsum); }
    return 0;
}

/* HairStyle::voxelize_vertices, hair_style.cc:257-294. */
int oracle_voxelize_vertices(const float* xyz, uint64_t n_vertices,
                             const float* tangents_in, const float origin[3], const float size[3],
                             uint64_t W, uint64_t H, uint64_t D, uint32_t flags,
                             uint8_t* densities_out, int8_t* tangents_out) {
    grid_t g;
    grid_init(&g, origin, size, W, H, D, flags);
    memset(densities_out, 0, g.n_voxels);
    float* tsum = NULL;
    if (tangents_in && tangents_out) {
        tsum = (float*)calloc(g.n_voxels * 3, sizeof(float));
        if (!tsum) return -1;
    }
    walk_vertices(&g, xyz, n_vertices, tsum ? tangents_in : NULL, densities_out, NULL, tsum);
    if (tsum) { quantise(&g, densities_out, tsum, tangents_out); free(tsum); }
    return 0;
}

/* Unclamped u32 hit counts (ADDED to counts_inout): the quantity a rank holds
 * before the multi-GPU sum; density = min(sum of counts, 255). */
void oracle_count_segments(const float* xyz, const uint32_t* indices, uint64_t n_indices,
                           const float origin[3], const float size[3],
                           uint64_t W, uint64_t H, uint64_t D, uint32_t flags, uint32_t* counts_inout) {
    grid_t g;
    grid_init(&g, origin, size, W, H, D, flags);
    walk_segments(&g, xyz, indices, n_indices, NULL, NULL, counts_inout, NULL);
}

void oracle_count_vertices(const float* xyz, uint64_t n_vertices,
                           const float origin[3], const float size[3],
                           uint64_t W, uint64_t H, uint64_t D, uint32_t flags, uint32_t* counts_inout) {
    grid_t g;
    grid_init(&g, origin, size, W, H, D, flags);
    walk_vertices(&g, xyz, n_vertices, NULL, NULL, counts_inout, NULL);
}

/* Number of samples the walk of hair_style.cc:319 takes (ceil(steps) per
 * segment) -- used to report sigma = samples / segment. */
uint64_t oracle_count_samples(const float* xyz, const uint32_t* indices, uint64_t n_indices,
                              const float origin[3], const float size[3],
                              uint64_t W, uint64_t H, uint64_t D) {
    const uint32_t flags = 0;
    grid_t g;
    grid_init(&g, origin, size, W, H, D, flags);
    uint64_t n = 0;
    if (n_indices < 2) return 0;
    for (uint64_t i = 0; i < n_indices - 1; i += 2) {
        const float* a = &xyz[3 * (uint64_t)indices[i]];
        const float* b = &xyz[3 * (uint64_t)indices[i + 1]];
        float m = 0.0f;
        for (int c = 0; c < 3; ++c) {
            float r = (a[c] - g.origin[c]) / g.voxel_size[c];
            float t = (b[c] - g.origin[c]) / g.voxel_size[c];
            float d = fabsf(t - r);
            if (c == 0) m = d; else m = glm_max(m, d);
        }
        if (!(m < 16777216.0f)) continue;
        while (m-- > 0.0f) ++n;
    }
    return n;
}

/* HairStyle::Volume::normalize, hair_style.cc:344-357. */
void oracle_normalize(uint8_t* densities, uint64_t n) {
    unsigned char lo = 255, hi = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (densities[i] > hi) hi = densities[i];
        if (densities[i] < lo) lo = densities[i];
    }
    if (hi == lo) return;                                  /* 255/0 in the reference */
    float scaling = 255.0f / (float)(hi - lo);             /* :351 */
    for (uint64_t i = 0; i < n; ++i) {
        unsigned char d = (unsigned char)(densities[i] - lo);   /* :354 */
        densities[i] = (unsigned char)((float)d * scaling);     /* :355 */
    }
}

/* HairStyle::Volume::downsample, hair_style.hh:228-257, with the functor
 * fixed to one of: 0 max, 1 sum/8, 2 (u8)sum, 3 min over the 2x2x2 block
 * ordered x + 2y + 4z.  Output resolution is floor(res / 2). */
void oracle_downsample(const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
                       int filter, uint8_t* out) {
    uint32_t w = W / 2, h = H / 2, d = D / 2;
    for (uint32_t k = 0; k < d; ++k)
    for (uint32_t j = 0; j < h; ++j)
    for (uint32_t i = 0; i < w; ++i) {
        unsigned char n[8];
        for (int z = 0; z < 2; ++z)
        for (int y = 0; y < 2; ++y)
        for (int x = 0; x < 2; ++x)
            n[x + 2 * y + 4 * z] =
                densities[(2 * i + x) + (uint64_t)(2 * j + y) * W + (uint64_t)(2 * k + z) * W * H];
        unsigned s = 0, mx = 0, mn = 255;
        for (int t = 0; t < 8; ++t) { s += n[t]; if (n[t] > mx) mx = n[t]; if (n[t] < mn) mn = n[t]; }
        unsigned char r = filter == 0 ? (unsigned char)mx
                        : filter == 1 ? (unsigned char)(s / 8)
                        : filter == 2 ? (unsigned char)s : (unsigned char)mn;
        out[i + (uint64_t)j * w + (uint64_t)k * w * h] = r;
    }
}

/* FNV-1a 64 over a byte buffer: the fingerprint used by tests/golden. */
uint64_t oracle_fnv1a64(const uint8_t* p, uint64_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
