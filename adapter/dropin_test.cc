// dropin_test.cc -- the reference's own call sequence, once through the reference and once through the drop-in.
//
// Links the UNMODIFIED reference translation unit src/vkhr/scene_graph/hair_style.cc (compiled where it lies, by
// adapter/Makefile) and libvkhr_b200.so.  Builds a HairStyle the way SceneGraph::add_style prepares one
// (src/vkhr/scene_graph.cc:222-245, minus the random shuffle), then compares
//     hs.voxelize_segments(W,H,D) [+ normalize()]      vs      vkhr_b200::voxelize_segments(hs, W,H,D) [+ normalize]
// byte for byte on the densities and within 1 LSB on the tangents of unsaturated voxels.
// Test infrastructure: needs a B200; prints one line per case and exits non-zero on any mismatch.
#include "vkhr_b200_adapter.hh"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// seeded random-walk strands (an input generator; its arithmetic is not part of parity)
static vkhr::HairStyle make_style(unsigned strands, unsigned segs, unsigned seed, float seg_len, bool explicit_aabb_from_file) {
    vkhr::HairStyle hs;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    hs.set_strand_count(strands);
    hs.set_default_segment_count(segs);
    hs.vertices.reserve(size_t(strands) * (segs + 1));
    for (unsigned s = 0; s < strands; ++s) {
        glm::vec3 p(-25.0f + 50.0f * u(rng), 60.0f + 40.0f * u(rng), -25.0f + 50.0f * u(rng));
        glm::vec3 dir(0.0f, -1.0f, 0.0f);
        hs.vertices.push_back(p);
        for (unsigned k = 0; k < segs; ++k) {
            dir = glm::normalize(dir + glm::vec3(0.6f * (u(rng) - 0.5f), 0.6f * (u(rng) - 0.5f) - 0.35f, 0.6f * (u(rng) - 0.5f)));
            p += seg_len * dir;
            hs.vertices.push_back(p);
        }
    }
    hs.generate_tangents();
    hs.generate_indices();
    hs.generate_bounding_box();
    (void)explicit_aabb_from_file;
    return hs;
}

static int compare(const char* name, const vkhr::HairStyle::Volume& ref, const vkhr::HairStyle::Volume& got, double t_ref, double t_got) {
    int bad = 0;
    if (ref.resolution != got.resolution) { std::printf("  %s: resolution differs\n", name); ++bad; }
    if (std::memcmp(&ref.bounds, &got.bounds, sizeof ref.bounds) != 0) { std::printf("  %s: bounds differ\n", name); ++bad; }
    size_t dens_bad = 0, tan_bad = 0, tan_checked = 0;
    if (ref.densities.size() != got.densities.size()) { std::printf("  %s: density size differs\n", name); return bad + 1; }
    for (size_t i = 0; i < ref.densities.size(); ++i) dens_bad += ref.densities[i] != got.densities[i];
    if (got.tangents.size() == ref.tangents.size())
        for (size_t i = 0; i < ref.tangents.size(); ++i) {
            if (ref.densities[i] == 0 || ref.densities[i] == 255) continue;   // 0/0 in the reference; order-dependent when saturated
            ++tan_checked;
            for (int c = 0; c < 4; ++c) tan_bad += std::abs(int(ref.tangents[i][c]) - int(got.tangents[i][c])) > 1;
        }
    std::printf("%-28s densities %s (%zu of %zu differ)  tangents %s (%zu comps off by > 1 LSB in %zu voxels)  reference %.3f s  b200 %.4f s\n",
                name, dens_bad ? "MISMATCH" : "bit-exact", dens_bad, ref.densities.size(),
                got.tangents.size() != ref.tangents.size() ? "not produced" : (tan_bad ? "MISMATCH" : "within 1 LSB"),
                tan_bad, tan_checked, t_ref, t_got);
    return bad + (dens_bad != 0) + (tan_bad != 0);
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && std::strcmp(argv[1], "--quick") == 0;
    int bad = 0;
    struct Case { const char* name; unsigned strands, segs; float seg_len; size_t w, h, d; bool vertices, normalize; };
    const Case cases[] = {
        {"segments 64^3", 2000, 12, 1.5f, 64, 64, 64, false, false},
        {"vertices 64^3", 2000, 12, 1.5f, 64, 64, 64, true, false},
        {"segments 30x20x10 saturating", 4000, 12, 1.5f, 30, 20, 10, false, false},
        {"segments 128^3 + normalize", 20000, 12, 0.8f, 128, 128, 128, false, true},
        {"ponytail-shaped 256^3 + norm", 136320, 12, 0.5f, 256, 256, 256, false, true},   // the call rasterizer/hair_style.cc:75,:77 makes
    };
    try {
        vkhr_b200::Context::instance();
        for (const Case& c : cases) {
            if (quick && c.strands > 50000) continue;
            vkhr::HairStyle hs = make_style(c.strands, c.segs, 7u + c.strands, c.seg_len, false);
            double t0 = now_s();
            vkhr::HairStyle::Volume ref = c.vertices ? hs.voxelize_vertices(c.w, c.h, c.d) : hs.voxelize_segments(c.w, c.h, c.d);
            if (c.normalize) ref.normalize();
            double t1 = now_s();
            vkhr::HairStyle::Volume got = c.vertices ? vkhr_b200::voxelize_vertices(hs, c.w, c.h, c.d) : vkhr_b200::voxelize_segments(hs, c.w, c.h, c.d);
            if (c.normalize) vkhr_b200::normalize(got);
            double t2 = now_s();
            bad += compare(c.name, ref, got, t1 - t0, t2 - t1);
        }
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 2;
    }
    std::printf(bad ? "DROP-IN TEST FAILED\n" : "drop-in test ok\n");
    return bad ? 1 : 0;
}
