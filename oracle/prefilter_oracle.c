/* prefilter_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of what the reference's fragment shaders compute from the strand
 * density volume -- evaluated at every voxel CENTRE, which is what a precomputed (prefiltered)
 * volume stores.  Paths are relative to the reference tree:
 *
 *   sample_volume                      share/shaders/volumes/sample_volume.glsl:7-9
 *   filter_volume (Gaussian)           share/shaders/volumes/sample_volume.glsl:12-35
 *   local_ambient_occlusion            share/shaders/volumes/local_ambient_occlusion.glsl:9-30
 *   volume_approximated_deep_shadows   share/shaders/self-shadowing/approximate_deep_shadows.glsl:24-36
 *   sampler / image format             src/vkhr/rasterizer/hair_style.cc:79-85,:94-101 (R8_UNORM, LINEAR,
 *                                      CLAMP_TO_BORDER) ; border colour opaque black (src/vkpp/sampler.cc:51)
 *   call sites and constants           share/shaders/strands/strand.frag:69-74, volumes/volume.frag:72-87
 *                                      (kernel_size 2, thickness 11.0), include/vkhr/rasterizer/interface.hh:101-105
 *
 * PARITY UNPINNED BY THE REFERENCE: there is no GLSL toolchain or Vulkan device here, the reference
 * has no tests for these functions, and GPU texture filtering is not bit-specified (fixed-point
 * weights).  This file IS the definition the CUDA prefilter is held to (<= 1e-6 relative):
 *
 *   - texel value tau(x,y,z) = (float)u8 / 255.0f inside the grid, 0.0f outside (border);
 *   - a sample at a voxel centre displaced by `o` voxels along an axis has unnormalised texel
 *     coordinate i + o, exactly (the voxel-size factors of the shader cancel); with r = |o|,
 *     fl = floor(r), fp = r - fl the two texels and weights are
 *         o > 0:  (i + fl    , 1 - fp) and (i + fl + 1, fp)
 *         o < 0:  (i - fl - 1, fp)     and (i - fl    , 1 - fp)          [fp == 0: (i - fl, 1) alone]
 *     i.e. the weights are the same for every voxel of the grid;
 *   - trilinear = nested weighted sums, x first, then y, then z, each `a*wa + b*wb` with separately
 *     rounded fp32 operations (compiled with -ffp-contract=off);
 *   - loops, accumulation order and the remaining arithmetic follow the GLSL text literally.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct {
    const uint8_t* d;
    int W, H, D;
} vol_t;

/* texel fetch with CLAMP_TO_BORDER / opaque black, R8_UNORM decode */
static float tau(const vol_t* v, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= v->W || y >= v->H || z >= v->D) return 0.0f;
    return (float)v->d[(size_t)x + (size_t)y * v->W + (size_t)z * v->W * v->H] / 255.0f;
}

/* The two texels and weights of one axis of a LINEAR sample displaced by +-r voxels from centre i. */
typedef struct { int o0, o1; float w0, w1; } axis_t;   /* texel = i + o0 / i + o1 */

static axis_t axis_taps(float r, int positive) {
    axis_t a;
    const float fl = floorf(r), fp = r - fl;
    const int ifl = (int)fl;
    if (positive) { a.o0 = ifl; a.o1 = ifl + 1; a.w0 = 1.0f - fp; a.w1 = fp; }
    else if (fp != 0.0f) { a.o0 = -ifl - 1; a.o1 = -ifl; a.w0 = fp; a.w1 = 1.0f - fp; }
    else { a.o0 = -ifl; a.o1 = -ifl + 1; a.w0 = 1.0f; a.w1 = 0.0f; }
    return a;
}

static float trilinear(const vol_t* v, int i, int j, int k, const axis_t* ax, const axis_t* ay, const axis_t* az) {
    float zz[2];
    for (int c = 0; c < 2; ++c) {
        const int z = k + (c ? az->o1 : az->o0);
        float yy[2];
        for (int b = 0; b < 2; ++b) {
            const int y = j + (b ? ay->o1 : ay->o0);
            const float t0 = tau(v, i + ax->o0, y, z), t1 = tau(v, i + ax->o1, y, z);
            yy[b] = t0 * ax->w0 + t1 * ax->w1;
        }
        zz[c] = yy[0] * ay->w0 + yy[1] * ay->w1;
    }
    return zz[0] * az->w0 + zz[1] * az->w1;
}

/* local_ambient_occlusion(volume, centre(i,j,k), origin, size, kernel_size = 2, radius, intensity, min_intensity)
 * (local_ambient_occlusion.glsl:9-30).  kernel_size is the literal 2 of both call sites:
 * kernel_radius = 0.5, voxel_scaling = radius / 0.5, the loops visit x,y,z in {-0.5,+0.5}, so a tap sits
 * 0.5 * voxel_scaling = radius voxels from the centre along each axis. */
void oracle_prefilter_ao(const uint8_t* dens, uint32_t W, uint32_t H, uint32_t D,
                         float radius, float intensity, float min_intensity, float* out) {
    const vol_t v = {dens, (int)W, (int)H, (int)D};
    const float kernel_size = 2.0f;
    const float kernel_radius = (kernel_size - 1.0f) / 2.0f;
    const float voxel_scaling = radius / kernel_radius;
    const float r = kernel_radius * voxel_scaling;                 /* |offset| in voxels */
    const axis_t neg = axis_taps(r, 0), pos = axis_taps(r, 1);
    const float norm = powf(kernel_size, 3.0f);
    for (int k = 0; k < v.D; ++k)
        for (int j = 0; j < v.H; ++j)
            for (int i = 0; i < v.W; ++i) {
                float density = 0.0f;
                for (int z = 0; z < 2; ++z)
                    for (int y = 0; y < 2; ++y)
                        for (int x = 0; x < 2; ++x) {
                            const float s = trilinear(&v, i, j, k, x ? &pos : &neg, y ? &pos : &neg, z ? &pos : &neg);
                            density += (min_intensity < s) ? min_intensity : s;      /* GLSL min(s, min_intensity) */
                        }
                out[(size_t)i + (size_t)j * W + (size_t)k * W * H] = powf(1.0f - density / norm, intensity);
            }
}

/* One raymarch step of volume_approximated_deep_shadows (approximate_deep_shadows.glsl:24-36) through a voxel
 * centre: strands = tau * thickness; visibility factor pow(1 - strand_alpha, strands).  The shader's result
 * over a ray is the product of these factors (pow(a, s1 + s2) = pow(a, s1) pow(a, s2)). */
void oracle_prefilter_opacity(const uint8_t* dens, uint64_t n, float strand_alpha, float thickness, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float strands = ((float)dens[i] / 255.0f) * thickness;
        out[i] = powf(1.0f - strand_alpha, strands);
    }
}

/* filter_volume(volume, kernel_width, centre(i,j,k), ...).r  (sample_volume.glsl:12-35), including the
 * `/ 2.0f*sigma_squared` precedence of :26 (divides by 2, then MULTIPLIES by sigma^2).  Offsets are whole
 * voxels, so every tap is a texel centre (weight 1).  The float loop counters follow the GLSL text. */
void oracle_prefilter_gauss(const uint8_t* dens, uint32_t W, uint32_t H, uint32_t D, float kernel_width, float* out) {
    const vol_t v = {dens, (int)W, (int)H, (int)D};
    const float kernel_range = (kernel_width - 1.0f) / 2.0f;
    const float sigma_stddev = (kernel_width / 2.0f) / 2.4f;
    const float sigma_squared = sigma_stddev * sigma_stddev;
    const float M_PI_F = (float)3.14159265358979323846, M_E_F = (float)2.71828182845904523536;
    for (int k = 0; k < v.D; ++k)
        for (int j = 0; j < v.H; ++j)
            for (int i = 0; i < v.W; ++i) {
                float density = 0.0f, total_weight = 0.0f;
                for (float z = -kernel_range; z <= +kernel_range; z += 1.0f)
                    for (float y = -kernel_range; y <= +kernel_range; y += 1.0f)
                        for (float x = -kernel_range; x <= +kernel_range; x += 1.0f) {
                            const float exponent = -1.0f * (x * x + y * y + z * z) / 2.0f * sigma_squared;
                            const float local_weight = 1.0f / (2.0f * M_PI_F * sigma_squared) * powf(M_E_F, exponent);
                            /* centre + (x,y,z) voxels: for odd widths a texel centre */
                            const float fx = floorf((float)i + x + 0.5f), fy = floorf((float)j + y + 0.5f), fz = floorf((float)k + z + 0.5f);
                            density += tau(&v, (int)fx, (int)fy, (int)fz) * local_weight;
                            total_weight += local_weight;
                        }
                out[(size_t)i + (size_t)j * W + (size_t)k * W * H] = density / total_weight;
            }
}

/* ---------------------------------------------------------------------------------------------------------
 * Volumetric ADSM transmittance volume: volume_approximated_deep_shadows(volume, centre(i,j,k), light, steps,
 * strand_alpha, origin, size, thickness) (approximate_deep_shadows.glsl:24-36; call site volume.frag:72-78 with
 * steps = raycast_steps = 1024 (interface.hh:101-105), thickness = 11.0) evaluated at every voxel centre: the
 * visibility of the light from that voxel through the density volume.
 *
 * Arithmetic contract (fp32, every operation rounded separately, GLSL text order):
 *   voxel_size = size / (W,H,D);  p = origin + ((float)i + 0.5f) * voxel_size        (the voxel centre, world space)
 *   step_size = 1.0f / steps;  for (t = 0.0f; t < 1.0f; t += step_size)               (t accumulated in fp32)
 *     point = p * (1.0f - t) + light * t                                               (GLSL mix)
 *     u = (point - origin) / size                                                      (sample_volume.glsl:8)
 *     LINEAR / CLAMP_TO_BORDER fetch at unnormalised coordinate c = u * res - 0.5f per axis:
 *       i0 = floor(c), f = c - i0, texels i0 and i0 + 1 with weights (1.0f - f) and f, nested x, y, z as above
 *     strands += sample * thickness
 *   result = powf(1.0f - strand_alpha, strands)
 * t_table (n_t entries) is the accumulated t sequence; the caller passes the same table to the GPU. */
static float trilinear_at(const vol_t* v, float cx, float cy, float cz) {
    const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
    /* outside [-1, res) on any axis every tap is border: 0 (also keeps the int conversions in range) */
    if (!(fx0 >= -1.0f && fx0 < (float)v->W && fy0 >= -1.0f && fy0 < (float)v->H && fz0 >= -1.0f && fz0 < (float)v->D)) return 0.0f;
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
    const float fx = cx - fx0, fy = cy - fy0, fz = cz - fz0;
    const float wx0 = 1.0f - fx, wy0 = 1.0f - fy, wz0 = 1.0f - fz;
    float zz[2];
    for (int c = 0; c < 2; ++c) {
        float yy[2];
        for (int b = 0; b < 2; ++b) {
            const float t0 = tau(v, x0, y0 + b, z0 + c), t1 = tau(v, x0 + 1, y0 + b, z0 + c);
            yy[b] = t0 * wx0 + t1 * fx;
        }
        zz[c] = yy[0] * wy0 + yy[1] * fy;
    }
    return zz[0] * wz0 + zz[1] * fz;
}

uint32_t oracle_adsm_t_table(float steps, float* table, uint32_t cap) {
    const float step_size = 1.0f / steps;
    uint32_t n = 0;
    for (float t = 0.0f; t < 1.0f; t += step_size) {
        if (n < cap && table) table[n] = t;
        ++n;
        if (n > (1u << 20)) break;                   /* steps so large that t stops advancing: not a usable setting */
    }
    return n;
}

void oracle_prefilter_adsm(const uint8_t* dens, uint32_t W, uint32_t H, uint32_t D,
                           const float origin[3], const float size[3], const float light[3],
                           float steps, float strand_alpha, float thickness, float* out) {
    const vol_t v = {dens, (int)W, (int)H, (int)D};
    const float res[3] = {(float)W, (float)H, (float)D};
    float vs[3];
    for (int c = 0; c < 3; ++c) vs[c] = size[c] / res[c];
    const float step_size = 1.0f / steps;
    const float base = 1.0f - strand_alpha;
    for (int k = 0; k < v.D; ++k)
        for (int j = 0; j < v.H; ++j)
            for (int i = 0; i < v.W; ++i) {
                const float p[3] = {origin[0] + ((float)i + 0.5f) * vs[0], origin[1] + ((float)j + 0.5f) * vs[1],
                                    origin[2] + ((float)k + 0.5f) * vs[2]};
                float strands = 0.0f;
                for (float t = 0.0f; t < 1.0f; t += step_size) {
                    float c[3];
                    for (int a = 0; a < 3; ++a) {
                        const float point = p[a] * (1.0f - t) + light[a] * t;
                        const float u = (point - origin[a]) / size[a];
                        c[a] = u * res[a] - 0.5f;
                    }
                    strands += trilinear_at(&v, c[0], c[1], c[2]) * thickness;
                }
                out[(size_t)i + (size_t)j * W + (size_t)k * W * H] = powf(base, strands);
            }
}
