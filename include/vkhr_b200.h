/* vkhr_b200.h -- C ABI of the B200-native strand voxeliser.
 *
 * Drop-in boundary for ONE path of CaffeineViking/vkhr: voxelising hair strand
 * vertices/segments into the 8-bit density volume (reference
 * `vkhr::HairStyle::voxelize_vertices` / `voxelize_segments` returning a
 * `vkhr::HairStyle::Volume`).  Plain C types only: no STL, GLM or torch types
 * cross this boundary.  Paths below are relative to the reference tree.
 *
 * Data contract (identical to the reference):
 *   vertices    V x 3 float32, 12-byte stride          (HairStyle::vertices, hair_style.hh:126)
 *   indices     2 per segment, uint32 vertex ids       (HairStyle::indices,  hair_style.hh:139)
 *   AABB        origin = bbox_min, size = max - min    (HairStyle::get_bounding_box, hair_style.cc:236-255)
 *   densities   W*H*D uint8, x fastest, then y, then z (Volume::densities, hair_style.hh:93;
 *               uploaded as VK_FORMAT_R8_UNORM by rasterizer/hair_style.cc:94-101)
 *   tangents    W*H*D x int8[4], w = 0                 (Volume::tangents, hair_style.hh:94; R8G8B8A8_SNORM)
 *   density     = min(number of samples landing in the voxel, 255)  (hair_style.cc:277-280, :322-325)
 *
 * Numerics: fp32 IEEE, the exact operation order of hair_style.cc:257-281 and
 * :296-329, including the fp32 linear index of :276/:321 (which rounds for
 * grids above 2^24 voxels).  Densities are bit-exact with the reference built
 * with strict IEEE flags.  Where the reference is undefined this library
 * defines: samples whose fp32 index is NaN/negative/>= W*H*D are dropped;
 * segments whose step count is not < 2^24 are skipped; normalize with
 * max == min leaves the grid unchanged; fewer than 2 indices => empty volume.
 *
 * Threading: one context = one device + one internal stream.  Calls on one
 * context must be serialised by the caller; distinct contexts are independent.
 * Host-pointer entry points are synchronous.  `_dev` entry points take device
 * pointers, enqueue on the given cudaStream_t (NULL = the context's own stream;
 * pass cudaStreamLegacy, (void*)1, for the legacy default stream) and return
 * without synchronising.  All per-call scratch (counters, flags, staging) is
 * owned by the context and shared between calls: when a call arrives on a
 * different stream than the previous one, the library orders it behind the
 * previous call with an event (record on the old stream, wait on the new), so
 * switching streams is safe; two calls may still not be ISSUED concurrently
 * from two host threads on one context.
 *
 * Errors: every call returns 0 or a negative vkhr_b200_status; nothing throws.
 * There is no CPU fallback: without a usable sm_100 device vkhr_b200_create fails.
 */
#ifndef VKHR_B200_H
#define VKHR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKHR_B200_API __declspec(dllexport)
#else
#define VKHR_B200_API __attribute__((visibility("default")))
#endif

typedef struct vkhr_b200_ctx vkhr_b200_ctx;

typedef enum vkhr_b200_status {
    VKHR_B200_OK = 0,
    VKHR_B200_ERR_INVALID_ARGUMENT = -1,
    VKHR_B200_ERR_CUDA = -2,
    VKHR_B200_ERR_OUT_OF_MEMORY = -3,
    VKHR_B200_ERR_NO_DEVICE = -4,
    VKHR_B200_ERR_UNSUPPORTED = -5
} vkhr_b200_status;

/* flags (bitwise or) */
enum {
    /* Linear voxel index computed in exact integers instead of the reference's
     * fp32 expression (hair_style.cc:276,:321).  Differs from the reference
     * only for grids above 2^24 voxels.  Default (0) = reference behaviour. */
    VKHR_B200_INDEX_EXACT = 1u << 0,
    /* Apply Volume::normalize() (hair_style.cc:344-357) to the densities, as
     * the only caller does (rasterizer/hair_style.cc:77). */
    VKHR_B200_NORMALIZE = 1u << 1,
    /* Kernel strategy override (default: BRICK8 where it can run, else PACKED8, else COUNT32).
     * COUNT32: u32 hit counts in context scratch, then clamp to u8.
     * PACKED8: saturating counts kept directly in the u8 output grid
     *          (byte-packed 32-bit atomics, exact overflow repair). */
    VKHR_B200_STRATEGY_COUNT32 = 1u << 8,
    VKHR_B200_STRATEGY_PACKED8 = 1u << 9,
    /* BRICK8: PACKED8 counted in a context-owned scratch volume stored as 4x4x2-voxel bricks (one
     *          32-byte sector each, so neighbouring samples of a strand share atomic request packets),
     *          then copied out to the x-fastest output layout.  Segments only, W % 4 == H % 4 == D % 2 == 0,
     *          and either at most 2^24 voxels or W and H powers of two (above 2^24 voxels the reference's fp32
     *          index rounds, and the brick is taken from the bit fields of the rounded index); anything else
     *          falls back to PACKED8.  The default picks it from about one segment per 100 voxels upwards. */
    VKHR_B200_STRATEGY_BRICK8 = 1u << 10,
    /* BRICK8 normally runs as ONE persistent launch per batch (walk and copy-out of consecutive instances side by side,
     * through a ring of scratch volumes that stays in L2).  This flag runs it as the separate kernels of round 1 instead
     * (clear, walk, copy-out, verdict): the cross-check of the frame kernel in the tests, and the A/B timing. */
    VKHR_B200_BRICK8_SPLIT = 1u << 11
};

/* 2x2x2 reduction functors for vkhr_b200_downsample (Volume::downsample takes
 * an arbitrary functor over the 8 texels ordered x + 2y + 4z, hair_style.hh:228-257). */
enum { VKHR_B200_DOWNSAMPLE_MAX = 0, VKHR_B200_DOWNSAMPLE_MEAN = 1,
       VKHR_B200_DOWNSAMPLE_SUM = 2, VKHR_B200_DOWNSAMPLE_MIN = 3 };

/* ---- context ---------------------------------------------------------- */
VKHR_B200_API int vkhr_b200_create(int device, vkhr_b200_ctx** out);
VKHR_B200_API void vkhr_b200_destroy(vkhr_b200_ctx* ctx);
/* Message of the last failing call on this context (ctx may be NULL for create failures). */
VKHR_B200_API const char* vkhr_b200_last_error(const vkhr_b200_ctx* ctx);
VKHR_B200_API const char* vkhr_b200_version(void);
/* The context's internal stream, as a cudaStream_t. */
VKHR_B200_API void* vkhr_b200_stream(vkhr_b200_ctx* ctx);
VKHR_B200_API int vkhr_b200_synchronize(vkhr_b200_ctx* ctx);
/* Number of kernels this context has launched so far. */
VKHR_B200_API uint64_t vkhr_b200_launch_count(const vkhr_b200_ctx* ctx);
/* Strategy the last voxelisation of this context ran with: VKHR_B200_STRATEGY_COUNT32 / _PACKED8 / _BRICK8
 * (0 before the first call).  Introspection for tests and benchmarks; no reference counterpart. */
VKHR_B200_API uint32_t vkhr_b200_last_strategy(const vkhr_b200_ctx* ctx);

/* Bytes of BRICK8 scratch the frame kernel keeps in flight: a ring of floor(bytes / (W*H*D)) volumes (at most eight;
 * below two volumes the separate kernels run instead), sized to sit in L2 beside the streams of strands and output
 * volumes (default 64 MiB of the B200's 126 MB: four 256^3 volumes; three and five were measured slower).  Tuning
 * knob; results do not depend on it. */
VKHR_B200_API int vkhr_b200_set_scratch_ring_bytes(vkhr_b200_ctx* ctx, size_t bytes);

/* Per-phase device timing.  While enabled, every voxelize call records CUDA
 * events on its launching stream around its phases; profile_read waits for
 * them and returns, since the previous read, the summed milliseconds and the
 * number of spans of: [0] grid clear, [1] strand walk (the dominant kernel),
 * [2] finish (overflow repair or u32->u8 clamp), [3] normalize.
 * profile_read_ex also returns [4] prefilter (n_phases = 5). */
VKHR_B200_API int vkhr_b200_profile_enable(vkhr_b200_ctx* ctx, int enable);
VKHR_B200_API int vkhr_b200_profile_read(vkhr_b200_ctx* ctx, double ms_out[4], uint32_t spans_out[4]);
VKHR_B200_API int vkhr_b200_profile_read_ex(vkhr_b200_ctx* ctx, double* ms_out, uint32_t* spans_out, uint32_t n_phases);

/* ---- host-pointer API: replaces HairStyle::voxelize_segments ----------
 * (hair_style.hh:104, hair_style.cc:296-342).
 * indices == NULL with segs_per_strand > 0 means uniform strands: the implicit
 * pairs generate_indices() (hair_style.cc:196-213) would produce.
 * tangents_in / tangents_out may be NULL (density only). */
VKHR_B200_API int vkhr_b200_voxelize_segments(
    vkhr_b200_ctx* ctx, const float* vertices, uint32_t n_vertices,
    const uint32_t* indices, uint64_t n_indices, uint32_t segs_per_strand,
    const float* tangents_in, const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint8_t* densities_out, int8_t* tangents_out);

/* Replaces HairStyle::voxelize_vertices (hair_style.hh:103, hair_style.cc:257-294). */
VKHR_B200_API int vkhr_b200_voxelize_vertices(
    vkhr_b200_ctx* ctx, const float* vertices, uint32_t n_vertices,
    const float* tangents_in, const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint8_t* densities_out, int8_t* tangents_out);

/* ---- device-pointer API (per-frame path: no host<->device copies) ----- */
VKHR_B200_API int vkhr_b200_voxelize_segments_dev(
    vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
    const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
    const float* d_tangents_in, const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint8_t* d_densities_out, int8_t* d_tangents_out, void* stream);

VKHR_B200_API int vkhr_b200_voxelize_vertices_dev(
    vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
    const float* d_tangents_in, const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint8_t* d_densities_out, int8_t* d_tangents_out, void* stream);

/* One crowd member: its own strands, AABB and output volume
 * (the reference keeps one Volume per vulkan::HairStyle, rasterizer.cc:148-151). */
typedef struct vkhr_b200_instance {
    const float*    d_vertices;       /* device, V x 3 float32 */
    const uint32_t* d_indices;        /* device or NULL (uniform strands) */
    uint64_t        n_indices;
    uint32_t        n_vertices;
    uint32_t        segs_per_strand;
    float           aabb_origin[3];
    float           aabb_size[3];
    uint8_t*        d_densities_out;  /* device, W*H*D */
} vkhr_b200_instance;

/* voxelize_segments for `n` independent instances (host array of descriptors,
 * device pointers inside), all at the same resolution. */
VKHR_B200_API int vkhr_b200_voxelize_segments_batch_dev(
    vkhr_b200_ctx* ctx, const vkhr_b200_instance* instances, uint32_t n,
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags, void* stream);

/* ---- host-pointer crowd API (pipelined) ----------------------------------- *
 * voxelize_segments for `n` independent instances whose strands and output
 * volumes live in HOST memory (one vkhr::HairStyle + one Volume each, as the
 * reference keeps them: rasterizer.cc:148-151).  The upload of instance k+1,
 * the kernels of instance k and the download of instance k-1 run concurrently
 * on separate streams / copy engines (PCIe is full duplex), so a crowd costs
 * about max(H2D, D2H) instead of their sum.  Synchronous: returns when every
 * volume is in host memory.  Host buffers should be page-locked
 * (vkhr_b200_host_register) -- pageable memory works but serialises the copies. */
typedef struct vkhr_b200_host_instance {
    const float*    vertices;         /* host, V x 3 float32 */
    const uint32_t* indices;          /* host or NULL (uniform strands) */
    uint64_t        n_indices;
    uint32_t        n_vertices;
    uint32_t        segs_per_strand;
    float           aabb_origin[3];
    float           aabb_size[3];
    uint8_t*        densities_out;    /* host, W*H*D */
} vkhr_b200_host_instance;

VKHR_B200_API int vkhr_b200_voxelize_segments_batch(
    vkhr_b200_ctx* ctx, const vkhr_b200_host_instance* instances, uint32_t n,
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags);

/* Page-lock / unlock a caller-owned host range (e.g. the storage of a std::vector)
 * so that copies to and from it are asynchronous DMA transfers. */
VKHR_B200_API int vkhr_b200_host_register(vkhr_b200_ctx* ctx, void* ptr, size_t bytes);
VKHR_B200_API int vkhr_b200_host_unregister(vkhr_b200_ctx* ctx, void* ptr);

/* ---- multi-GPU building blocks ---------------------------------------- *
 * A rank ADDS the hits of its shard of segments (vertices) into a W*H*D u32
 * grid it owns (not cleared here); the shards' grids are summed (NCCL
 * allreduce, integer) and clamped: density = min(sum, 255) -- exact for any
 * partition because the reference counter saturates (hair_style.cc:322-325). */
VKHR_B200_API int vkhr_b200_count_segments_dev(
    vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
    const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
    const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint32_t* d_counts_inout, void* stream);

VKHR_B200_API int vkhr_b200_count_vertices_dev(
    vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
    const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    uint32_t* d_counts_inout, void* stream);

/* Multi-GPU combine on saturated u8 partials (each a complete voxelisation of one strand shard):
 * out[i] = min(sum over the n_slabs consecutive arrays of slab_bytes bytes, 255).  Exact, because
 * min(sum_r min(c_r, 255), 255) == min(sum_r c_r, 255) (hair_style.cc:322-325 only saturates). */
VKHR_B200_API int vkhr_b200_saturating_sum_u8_dev(
    vkhr_b200_ctx* ctx, const uint8_t* d_slabs, uint32_t n_slabs, uint64_t slab_bytes,
    uint8_t* d_out, void* stream);

/* The same combine fused with its exchange, over NVLink peer memory: d_partials[r] / d_outs[r] are rank r's
 * partial and output volumes as mapped into THIS process (symmetric memory / cudaIpc / cuMem peer mappings, up to 16).
 * The calling rank sums bytes [slab_offset, slab_offset + slab_bytes) of every partial with saturation and stores the
 * result into that range of every output: reduce-scatter + clamp + all-gather in one kernel.  The caller orders it
 * between two barriers of the group (all partials written before; all outputs written after). */
VKHR_B200_API int vkhr_b200_combine_peer_u8_dev(
    vkhr_b200_ctx* ctx, const void* const* d_partials, void* const* d_outs, uint32_t n_peers,
    uint64_t slab_offset_bytes, uint64_t slab_bytes, void* stream);

/* Sparse form of the fused combine: vkhr_b200_chunk_bitmap_dev writes one bit per 16-byte chunk of a volume (set =
 * the chunk holds a non-zero byte; n_bytes a multiple of 512); the combine reads the peers' bitmaps, fetches a chunk
 * only from the peers that have something in it and stores only non-zero results.  Every rank must have zeroed its
 * OUTPUT before the first barrier.  Slab offset and size: multiples of 512 bytes. */
VKHR_B200_API int vkhr_b200_chunk_bitmap_dev(
    vkhr_b200_ctx* ctx, const uint8_t* d_volume, uint64_t n_bytes, uint32_t* d_bitmap_out, void* stream);
VKHR_B200_API int vkhr_b200_combine_peer_u8_sparse_dev(
    vkhr_b200_ctx* ctx, const void* const* d_partials, const void* const* d_bitmaps, void* const* d_outs,
    uint32_t n_peers, uint64_t slab_offset_bytes, uint64_t slab_bytes, void* stream);

/* ---- strand-sharded voxelisation over the GPUs of one box, in ONE call per rank ------------------------------ *
 * BASELINE configs[2]: the strands of one hair style are split into contiguous ranges, one per GPU (rank); every rank
 * calls this with ITS strands and the SHARED AABB, and returns with the complete W*H*D volume in its own output
 * buffer -- the multi-GPU form of HairStyle::voxelize_segments (hair_style.cc:296-342).  Exact for every partition:
 * the reference's counter only saturates (hair_style.cc:322-325), so
 *   min(sum_r min(count_r, 255), 255) == min(sum_r count_r, 255).
 * Steps, all enqueued on `stream`: the rank's shard is voxelised into its partial volume (the single-GPU path, u8);
 * one bit per 16-byte chunk of the partial is published; a device-side barrier over the signal pads; ONE kernel reads
 * the rank's slab of every peer's partial straight from the peers' memory over NVLink (only the chunks the bitmaps mark),
 * adds with saturation and stores the slab into every peer's output; a second barrier.  No NCCL, no host round trip.
 *
 * The caller provides peer-mapped buffers (cudaIpc / cuMem / torch symmetric memory; on one device -- "fake ranks",
 * one context and stream per rank -- plain device pointers): for every rank r, as mapped into THIS process,
 *   partials[r]  padded volume bytes (vkhr_b200_sharded_volume_bytes), zero-filled once before the first call
 *   bitmaps[r]   padded bytes / 512 uint32 words
 *   outs[r]      padded volume bytes; the result of rank r
 *   signals[r]   2 x 16 uint32 words, zero-filled once before the first call
 * All ranks must make the same sequence of sharded calls (the barriers pair up by call count).  Ranks that share ONE
 * device (tests) must have made a plain voxelisation at this resolution before: a context's first call allocates its
 * scratch, and a device memory allocation serialises the device's streams -- behind another rank's waiting barrier.
 * (For the same reason vkhr_b200_create loads the library's kernels eagerly: CUDA's lazy loading can synchronise the context.) */
typedef struct vkhr_b200_shard_peers {
    uint32_t rank, world;                 /* world <= 16 */
    void* const* partials;
    void* const* bitmaps;
    void* const* outs;
    void* const* signals;
} vkhr_b200_shard_peers;
/* W*H*D rounded up so that every rank owns a slab of whole bitmap words (a multiple of 512 * world bytes). */
VKHR_B200_API uint64_t vkhr_b200_sharded_volume_bytes(uint32_t W, uint32_t H, uint32_t D, uint32_t world);
VKHR_B200_API int vkhr_b200_voxelize_segments_sharded_dev(
    vkhr_b200_ctx* ctx, const float* d_vertices, uint32_t n_vertices,
    const uint32_t* d_indices, uint64_t n_indices, uint32_t segs_per_strand,
    const float aabb_origin[3], const float aabb_size[3],
    uint32_t W, uint32_t H, uint32_t D, uint32_t flags,
    const vkhr_b200_shard_peers* peers, void* stream);

/* densities = min(counts, 255), optionally followed by normalize (flags). */
VKHR_B200_API int vkhr_b200_clamp_counts_dev(
    vkhr_b200_ctx* ctx, const uint32_t* d_counts, uint64_t n_voxels, uint32_t flags,
    uint8_t* d_densities_out, void* stream);

/* ---- Volume operations -------------------------------------------------- */
/* Volume::normalize (hair_style.cc:344-357), in place. */
VKHR_B200_API int vkhr_b200_normalize_dev(vkhr_b200_ctx* ctx, uint8_t* d_densities, uint64_t n_voxels, void* stream);
VKHR_B200_API int vkhr_b200_normalize(vkhr_b200_ctx* ctx, uint8_t* densities, uint64_t n_voxels);
/* Volume::downsample (hair_style.hh:228-257): (W/2)*(H/2)*(D/2) texels out. */
VKHR_B200_API int vkhr_b200_downsample_dev(vkhr_b200_ctx* ctx, const uint8_t* d_densities,
    uint32_t W, uint32_t H, uint32_t D, int filter, uint8_t* d_out, void* stream);
VKHR_B200_API int vkhr_b200_downsample(vkhr_b200_ctx* ctx, const uint8_t* densities,
    uint32_t W, uint32_t H, uint32_t D, int filter, uint8_t* out);
/* HairStyle::generate_bounding_box (hair_style.cc:215-234): min/max folded from
 * (0,0,0).  d_aabb_out / aabb_out = min.xyz, max.xyz (6 floats). */
VKHR_B200_API int vkhr_b200_generate_bounding_box_dev(vkhr_b200_ctx* ctx, const float* d_vertices,
    uint32_t n_vertices, float* d_aabb_out, void* stream);
VKHR_B200_API int vkhr_b200_generate_bounding_box(vkhr_b200_ctx* ctx, const float* vertices,
    uint32_t n_vertices, float aabb_out[6]);

/* Volume::save (hair_style.cc:359-369): the raw u8 densities, no header, to `path` (host memory in, host file out).
 * Returns VKHR_B200_OK, or VKHR_B200_ERR_INVALID_ARGUMENT when the file cannot be written (the reference returns false). */
VKHR_B200_API int vkhr_b200_volume_save(const char* path, const uint8_t* densities, uint64_t n_voxels);

/* ---- density -> AO / opacity / Gaussian prefilter -------------------------- *
 * What the reference's fragment shaders derive from the density volume at every
 * shaded point, precomputed once per voxel centre into float32 W*H*D volumes
 * (x fastest, like the densities):
 *   ao       local_ambient_occlusion(density, centre, ..., 2, ao_radius, ao_exponent, ao_max)
 *            (share/shaders/volumes/local_ambient_occlusion.glsl:9-30; call sites
 *            strands/strand.frag:69-74, volumes/volume.frag:82-87)
 *   opacity  pow(1 - strand_alpha, density * thickness): one raymarch step of
 *            volume_approximated_deep_shadows (self-shadowing/approximate_deep_shadows.glsl:24-36,
 *            thickness 11.0 at volumes/volume.frag:72-78); a ray's visibility is the product
 *   gauss    filter_volume(density, gauss_width, centre, ...).r (volumes/sample_volume.glsl:12-35)
 * with the density sampled as the reference samples it: R8_UNORM, LINEAR,
 * CLAMP_TO_BORDER / opaque black (rasterizer/hair_style.cc:79-85).  Any output
 * pointer may be NULL.  Results are within 1e-6 relative of the CPU restatement
 * of the GLSL (oracle/prefilter_oracle.c holds the arithmetic contract).
 * Defaults = include/vkhr/rasterizer/interface.hh:101-105. */
typedef struct vkhr_b200_prefilter_params {
    float ao_radius;       /* occlusion_radius, in voxels      (2.5)  */
    float ao_exponent;     /* ao_exponent                      (10)   */
    float ao_max;          /* ao_clamp                         (0.16) */
    float strand_alpha;    /* hair_alpha = the style's default_transparency (caller supplied; 0.3) */
    float thickness;       /* volume.frag:78                   (11)   */
    float gauss_width;     /* kernel_width of filter_volume, odd, 1..9 (3) */
    uint32_t flags;        /* VKHR_B200_PREFILTER_* */
} vkhr_b200_prefilter_params;
enum {
    VKHR_B200_PREFILTER_GENERIC = 1u << 0,   /* force the untiled kernel (testing) */
    VKHR_B200_PREFILTER_ROWWISE = 1u << 1,   /* tiled kernel: one AO evaluation per voxel instead of the register-tiled z column (testing, A/B timing) */
    VKHR_B200_PREFILTER_DENSE = 1u << 2      /* tiled kernel: load every tile, without the occupancy pre-pass that skips tiles of empty space (testing, A/B timing) */
};

VKHR_B200_API void vkhr_b200_prefilter_defaults(vkhr_b200_prefilter_params* params);
VKHR_B200_API int vkhr_b200_prefilter_dev(
    vkhr_b200_ctx* ctx, const uint8_t* d_densities, uint32_t W, uint32_t H, uint32_t D,
    const vkhr_b200_prefilter_params* params /* NULL = defaults */,
    float* d_ao_out, float* d_opacity_out, float* d_gauss_out, void* stream);
VKHR_B200_API int vkhr_b200_prefilter(
    vkhr_b200_ctx* ctx, const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
    const vkhr_b200_prefilter_params* params,
    float* ao_out, float* opacity_out, float* gauss_out);

/* ---- .hair bytes -> volume -------------------------------------------------- *
 * The load-time path of the reference in one call: HairStyle::load (hair_style.cc:24-47;
 * 128-byte header + raw arrays, hair_style.hh:147-174), what SceneGraph::add_style adds
 * when the file lacks it (scene_graph.cc:235-242: generate_indices from segments[] or
 * default_segment_count, generate_bounding_box; the strand shuffle does not change
 * densities), then voxelize_segments(W, H, D) over get_bounding_box()
 * (rasterizer/hair_style.cc:65-66,75).  The arrays are used in place from the file image:
 * vertices (and indices / tangents when present) are uploaded as they lie, indices of
 * strands with different lengths are generated on the device.  aabb_out (optional) receives
 * origin[3], size[3] -- the caller's volume_bounds.  tangents_out may be NULL. */
VKHR_B200_API int vkhr_b200_voxelize_hair(
    vkhr_b200_ctx* ctx, const void* hair_bytes, size_t n_bytes, uint32_t W, uint32_t H, uint32_t D,
    uint32_t flags, uint8_t* densities_out, int8_t* tangents_out, float aabb_out[6]);

/* ---- volumetric ADSM transmittance volume ---------------------------------- *
 * volume_approximated_deep_shadows(density, centre, light, steps, strand_alpha,
 * origin, size, thickness) of share/shaders/self-shadowing/approximate_deep_shadows.glsl:24-36
 * (call site volumes/volume.frag:72-78: steps = raycast_steps = 1024, thickness 11.0)
 * evaluated at every voxel centre: out[x + y*W + z*W*H] = visibility of the light from
 * that voxel, pow(1 - strand_alpha, sum over the shader's own samples of density * thickness).
 * The density is sampled as the reference samples it (R8_UNORM, LINEAR, CLAMP_TO_BORDER /
 * opaque black).  Within 1e-6 relative of the CPU restatement (oracle/prefilter_oracle.c). */
typedef struct vkhr_b200_adsm_params {
    float light[3];        /* lights[0].origin, world space */
    float steps;           /* raycast_steps                    (1024) */
    float strand_alpha;    /* hair_alpha                       (0.3)  */
    float thickness;       /* volume.frag:78                   (11)   */
} vkhr_b200_adsm_params;
VKHR_B200_API int vkhr_b200_adsm_dev(
    vkhr_b200_ctx* ctx, const uint8_t* d_densities, uint32_t W, uint32_t H, uint32_t D,
    const float aabb_origin[3], const float aabb_size[3], const vkhr_b200_adsm_params* params,
    float* d_visibility_out, void* stream);
VKHR_B200_API int vkhr_b200_adsm(
    vkhr_b200_ctx* ctx, const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
    const float aabb_origin[3], const float aabb_size[3], const vkhr_b200_adsm_params* params,
    float* visibility_out);

/* ---- device memory helpers (so a C/C++ caller needs no CUDA headers) ----- */
VKHR_B200_API int vkhr_b200_malloc(vkhr_b200_ctx* ctx, size_t bytes, void** d_ptr);
VKHR_B200_API int vkhr_b200_free(vkhr_b200_ctx* ctx, void* d_ptr);
VKHR_B200_API int vkhr_b200_memset(vkhr_b200_ctx* ctx, void* d_ptr, int value, size_t bytes, void* stream);
VKHR_B200_API int vkhr_b200_upload(vkhr_b200_ctx* ctx, void* d_dst, const void* src, size_t bytes, void* stream);
VKHR_B200_API int vkhr_b200_download(vkhr_b200_ctx* ctx, void* dst, const void* d_src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VKHR_B200_H */
