// sharded_test.cc -- the multi-GPU form of HairStyle::voxelize_segments through the plain C ABI, from C++.
//
// What a vkhr-side caller would write to split one hair style over the GPUs of a box:
// one context per rank, peer-mapped partial / bitmap / output / signal buffers, ONE call per rank
// (vkhr_b200_voxelize_segments_sharded_dev) -- and nothing else: no CUDA headers, no NCCL, no torch.
// Here the k ranks are k contexts on device 0 with their own streams ("fake ranks": plain device pointers stand in for
// the peer mappings), so the test runs on one GPU; on a box with several GPUs the same sequence runs with cudaIpc /
// cuMem mappings.  The result of every rank is compared byte for byte with the single-context volume of the whole set.
// Test infrastructure: needs a B200; prints one line per case and exits non-zero on any mismatch.
#include "../include/vkhr_b200.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#define CHECK(ctx, expr)                                                                             \
    do {                                                                                             \
        int rc_ = (expr);                                                                            \
        if (rc_ != VKHR_B200_OK) { std::printf("FAILED %s: %d %s\n", #expr, rc_, vkhr_b200_last_error(ctx)); return 2; } \
    } while (0)

int main() {
    const uint32_t strands = 6000, segs = 12, W = 128, H = 96, D = 64;
    const uint32_t vps = segs + 1;
    std::vector<float> xyz(size_t(strands) * vps * 3);
    std::mt19937 rng(17);
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};                     // generate_bounding_box folds from (0,0,0)
    for (uint32_t s = 0; s < strands; ++s) {
        float p[3] = {-25.0f + 50.0f * u(rng), 60.0f + 40.0f * u(rng), -25.0f + 50.0f * u(rng)};
        for (uint32_t k = 0; k < vps; ++k) {
            for (int c = 0; c < 3; ++c) {
                xyz[(size_t(s) * vps + k) * 3 + c] = p[c];
                lo[c] = p[c] < lo[c] ? p[c] : lo[c];
                hi[c] = p[c] > hi[c] ? p[c] : hi[c];
            }
            p[0] += 1.2f * (u(rng) - 0.5f); p[1] -= 0.9f * u(rng); p[2] += 1.2f * (u(rng) - 0.5f);
        }
    }
    const float size[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const size_t nv = size_t(W) * H * D;

    // the whole set on one context: the volume every rank must end up with
    vkhr_b200_ctx* one = nullptr;
    CHECK(nullptr, vkhr_b200_create(0, &one));
    std::vector<uint8_t> want(nv);
    CHECK(one, vkhr_b200_voxelize_segments(one, xyz.data(), strands * vps, nullptr, 0, segs, nullptr, lo, size, W, H, D, 0, want.data(), nullptr));

    int bad = 0;
    // (more fake ranks than hardware work queues would put one rank's kernels behind another rank's waiting barrier kernel:
    // CUDA_DEVICE_MAX_CONNECTIONS, 8 by default -- real ranks have a device each)
    for (uint32_t world : {2u, 3u, 4u}) {
        const uint64_t nvp = vkhr_b200_sharded_volume_bytes(W, H, D, world);
        std::vector<vkhr_b200_ctx*> ctx(world, nullptr);
        std::vector<void*> partials(world), bitmaps(world), outs(world), signals(world), verts(world);
        std::vector<uint32_t> first(world + 1);
        for (uint32_t r = 0; r <= world; ++r) first[r] = uint32_t(uint64_t(strands) * r / world);   // contiguous strand ranges
        for (uint32_t r = 0; r < world; ++r) {
            CHECK(nullptr, vkhr_b200_create(0, &ctx[r]));
            CHECK(ctx[r], vkhr_b200_malloc(ctx[r], nvp, &partials[r]));
            CHECK(ctx[r], vkhr_b200_malloc(ctx[r], nvp / 512 * 4, &bitmaps[r]));
            CHECK(ctx[r], vkhr_b200_malloc(ctx[r], nvp, &outs[r]));
            CHECK(ctx[r], vkhr_b200_malloc(ctx[r], 2 * 16 * 4, &signals[r]));
            CHECK(ctx[r], vkhr_b200_memset(ctx[r], partials[r], 0, nvp, nullptr));
            CHECK(ctx[r], vkhr_b200_memset(ctx[r], signals[r], 0, 2 * 16 * 4, nullptr));
            const size_t nverts = size_t(first[r + 1] - first[r]) * vps;
            CHECK(ctx[r], vkhr_b200_malloc(ctx[r], nverts * 12 + 16, &verts[r]));
            CHECK(ctx[r], vkhr_b200_upload(ctx[r], verts[r], xyz.data() + size_t(first[r]) * vps * 3, nverts * 12, nullptr));
            // ONE DEVICE ONLY: the first voxelisation of a context allocates its scratch, and a device memory allocation
            // serialises the device's streams (CUDA's implicit synchronisation) -- inside the sharded call that would put
            // rank r + 1's kernels behind rank r's waiting barrier kernel.  So every context allocates here, up front.
            // (Ranks on different GPUs do not share a device and need none of this.)
            CHECK(ctx[r], vkhr_b200_voxelize_segments_dev(ctx[r], static_cast<const float*>(verts[r]), (first[r + 1] - first[r]) * vps, nullptr, 0, segs,
                                                          nullptr, lo, size, W, H, D, 0, static_cast<uint8_t*>(outs[r]), nullptr, nullptr));
            CHECK(ctx[r], vkhr_b200_synchronize(ctx[r]));
        }
        std::printf("world %u: buffers ready\n", world); std::fflush(stdout);
        for (int frame = 0; frame < 2; ++frame) {
            for (uint32_t r = 0; r < world; ++r) {                    // asynchronous: every rank's chain is enqueued, then all are awaited
                vkhr_b200_shard_peers peers{r, world, partials.data(), bitmaps.data(), outs.data(), signals.data()};
                CHECK(ctx[r], vkhr_b200_voxelize_segments_sharded_dev(ctx[r], static_cast<const float*>(verts[r]), (first[r + 1] - first[r]) * vps,
                                                                     nullptr, 0, segs, lo, size, W, H, D, 0, &peers, nullptr));
            }
            std::printf("world %u frame %d: enqueued\n", world, frame); std::fflush(stdout);
            for (uint32_t r = 0; r < world; ++r) {
                CHECK(ctx[r], vkhr_b200_synchronize(ctx[r]));
                std::vector<uint8_t> got(nv);
                CHECK(ctx[r], vkhr_b200_download(ctx[r], got.data(), outs[r], nv, nullptr));
                CHECK(ctx[r], vkhr_b200_synchronize(ctx[r]));
                size_t diff = 0;
                for (size_t i = 0; i < nv; ++i) diff += got[i] != want[i];
                if (diff) { std::printf("world %u frame %d rank %u: %zu voxels differ\n", world, frame, r, diff); ++bad; }
            }
        }
        std::printf("sharded over %u ranks (one device, C ABI only): %s\n", world, bad ? "MISMATCH" : "every rank byte-identical to the one-context volume");
        for (uint32_t r = 0; r < world; ++r) {
            void* bufs[] = {partials[r], bitmaps[r], outs[r], signals[r], verts[r]};
            for (void* b : bufs) vkhr_b200_free(ctx[r], b);
            vkhr_b200_destroy(ctx[r]);
        }
    }
    vkhr_b200_destroy(one);
    std::printf(bad ? "FAILED\n" : "OK\n");
    return bad ? 1 : 0;
}
