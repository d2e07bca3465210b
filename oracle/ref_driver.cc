// oracle/ref_driver.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin first-party `extern "C"` driver around the UNMODIFIED reference
// voxeliser.  It is compiled together with the reference's own translation
// unit  /root/reference/src/vkhr/scene_graph/hair_style.cc  (where it lies,
// nothing is copied) by oracle/Makefile into oracle/_ref/libvkhr_ref.so.
//
// Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.
//
// What each entry point drives (paths relative to /root/reference):
//   vkhr_ref_create       HairStyle public members + generate_* helpers
//                         (src/vkhr/scene_graph/hair_style.cc:171-234) or, when
//                         an explicit AABB is requested, HairStyle::load
//                         (:24-45) of a temporary .hair file carrying the
//                         has_bounding_box header bit (hair_style.hh:147-174).
//   vkhr_ref_voxelize     HairStyle::voxelize_segments (:296-342) /
//                         HairStyle::voxelize_vertices (:257-294)
//   vkhr_ref_normalize    HairStyle::Volume::normalize (:344-357)
//   vkhr_ref_downsample   HairStyle::Volume::downsample (hair_style.hh:228-257)
#include <vkhr/scene_graph/hair_style.hh>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>

namespace {
struct HairFileHeader {            // on-disk layout, hair_style.hh:147-174
    char     signature[4];
    uint32_t strand_count, vertex_count, bitfield, default_segment_count;
    float    default_thickness, default_transparency, default_color[3];
    char     information[64];
    float    bbox_min[3], bbox_max[3];
};
static_assert(sizeof(HairFileHeader) == 128, ".hair header must be 128 bytes");

double now_s() {
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

void* vkhr_ref_load(const char* path) {
    auto* hs = new vkhr::HairStyle{};
    if (!hs->load(path)) { delete hs; return nullptr; }
    return hs;
}

// What SceneGraph::add_style does to a freshly loaded style (src/vkhr/scene_graph.cc:235-242), minus the
// random shuffle: generate whatever the file did not carry.
void vkhr_ref_prepare(void* h) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    if (!hs->has_tangents()) hs->generate_tangents();
    if (!hs->has_indices()) hs->generate_indices();
    if (!hs->has_bounding_box()) hs->generate_bounding_box();
}

// Build a HairStyle the way SceneGraph::add_style prepares one
// (src/vkhr/scene_graph.cc:222-245), minus the random shuffle: tangents,
// indices, and (when aabb_min == NULL) the generated bounding box.
void* vkhr_ref_create(const float* xyz, uint32_t n_vertices, uint32_t n_strands,
                      uint32_t default_segments, const uint16_t* segments,
                      const float* aabb_min, const float* aabb_max) {
    vkhr::HairStyle* hs = nullptr;
    if (aabb_min && aabb_max) {
        // No public setter for the header AABB: go through the reference's own
        // file loader with a header that has bit 7 (has_bounding_box) set.
        char tmpl[] = "/tmp/vkhr_ref_XXXXXX";
        int fd = mkstemp(tmpl);
        if (fd < 0) return nullptr;
        HairFileHeader h{};
        std::memcpy(h.signature, "HAIR", 4);
        h.strand_count = n_strands;
        h.vertex_count = n_vertices;
        h.bitfield = (segments ? 1u : 0u) | 2u | (1u << 7);
        h.default_segment_count = default_segments;
        std::memcpy(h.bbox_min, aabb_min, 12);
        std::memcpy(h.bbox_max, aabb_max, 12);
        FILE* f = fdopen(fd, "wb");
        bool ok = f && fwrite(&h, sizeof h, 1, f) == 1;
        if (ok && segments) ok = fwrite(segments, 2, n_strands, f) == n_strands;
        if (ok) ok = fwrite(xyz, 12, n_vertices, f) == n_vertices;
        if (f) fclose(f);
        if (ok) {
            hs = new vkhr::HairStyle{};
            if (!hs->load(tmpl)) { delete hs; hs = nullptr; }
        }
        unlink(tmpl);
        if (!hs) return nullptr;
    } else {
        hs = new vkhr::HairStyle{};
        hs->set_strand_count(n_strands);
        hs->set_default_segment_count(default_segments);
        if (segments) hs->segments.assign(segments, segments + n_strands);
        hs->vertices.resize(n_vertices);
        std::memcpy(hs->vertices.data(), xyz, size_t(n_vertices) * 12);
        hs->generate_bounding_box();
    }
    hs->generate_tangents();
    hs->generate_indices();
    return hs;
}

void vkhr_ref_destroy(void* h) { delete static_cast<vkhr::HairStyle*>(h); }

int vkhr_ref_save(void* h, const char* path) {
    return static_cast<vkhr::HairStyle*>(h)->save(path) ? 0 : -1;
}

uint32_t vkhr_ref_vertex_count(void* h)  { return static_cast<vkhr::HairStyle*>(h)->get_vertex_count(); }
uint32_t vkhr_ref_strand_count(void* h)  { return static_cast<vkhr::HairStyle*>(h)->get_strand_count(); }
uint32_t vkhr_ref_segment_count(void* h) { return static_cast<vkhr::HairStyle*>(h)->get_segment_count(); }
uint64_t vkhr_ref_index_count(void* h)   { return static_cast<vkhr::HairStyle*>(h)->indices.size(); }
uint32_t vkhr_ref_default_segment_count(void* h) { return static_cast<vkhr::HairStyle*>(h)->get_default_segment_count(); }
int vkhr_ref_has_bounding_box(void* h)   { return static_cast<vkhr::HairStyle*>(h)->has_bounding_box(); }

void vkhr_ref_get_vertices(void* h, float* out) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    std::memcpy(out, hs->vertices.data(), hs->vertices.size() * 12);
}
void vkhr_ref_get_tangents(void* h, float* out) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    std::memcpy(out, hs->tangents.data(), hs->tangents.size() * 12);
}
void vkhr_ref_get_indices(void* h, uint32_t* out) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    std::memcpy(out, hs->indices.data(), hs->indices.size() * 4);
}
uint32_t vkhr_ref_get_segments(void* h, uint16_t* out) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    if (out) std::memcpy(out, hs->segments.data(), hs->segments.size() * 2);
    return uint32_t(hs->segments.size());
}
// out8 = origin.xyz, radius, size.xyz, volume  (struct AABB, hair_style.hh:16-21)
void vkhr_ref_get_aabb(void* h, float* out8) {
    vkhr::AABB b = static_cast<vkhr::HairStyle*>(h)->get_bounding_box();
    out8[0] = b.origin.x; out8[1] = b.origin.y; out8[2] = b.origin.z; out8[3] = b.radius;
    out8[4] = b.size.x;   out8[5] = b.size.y;   out8[6] = b.size.z;   out8[7] = b.volume;
}

// mode 0 = voxelize_segments, 1 = voxelize_vertices.  Returns the wall time of
// the reference call itself (seconds) or a negative value on error.
double vkhr_ref_voxelize(void* h, int mode, uint64_t W, uint64_t H, uint64_t D,
                         int normalize, uint8_t* densities_out, int8_t* tangents_out) {
    auto* hs = static_cast<vkhr::HairStyle*>(h);
    if (mode == 0 && hs->indices.size() < 2) return -1.0;   // size()-1 underflows in the reference
    double t0 = now_s();
    vkhr::HairStyle::Volume v = (mode == 0) ? hs->voxelize_segments(W, H, D)
                                            : hs->voxelize_vertices(W, H, D);
    double t1 = now_s();
    if (normalize) v.normalize();
    if (densities_out) std::memcpy(densities_out, v.densities.data(), v.densities.size());
    if (tangents_out)  std::memcpy(tangents_out,  v.tangents.data(),  v.tangents.size() * 4);
    return t1 - t0;
}

// HairStyle::Volume::save (:359-369): raw dump of the densities.  Returns 1 on success like the reference's bool.
int vkhr_ref_volume_save(const uint8_t* densities, uint64_t n, const char* path) {
    vkhr::HairStyle::Volume v{};
    v.densities.assign(densities, densities + n);
    return v.save(path) ? 1 : 0;
}

void vkhr_ref_normalize(uint8_t* densities, uint64_t n) {
    vkhr::HairStyle::Volume v{};
    v.densities.assign(densities, densities + n);
    v.normalize();
    std::memcpy(densities, v.densities.data(), n);
}

// filter: 0 = max, 1 = truncated mean (sum/8), 2 = sum wrapped to u8, 3 = min
void vkhr_ref_downsample(const uint8_t* densities, uint32_t W, uint32_t H, uint32_t D,
                         int filter, uint8_t* out) {
    vkhr::HairStyle::Volume v{};
    v.resolution = glm::vec3(float(W), float(H), float(D));
    v.densities.assign(densities, densities + size_t(W) * H * D);
    vkhr::HairStyle::Volume d;
    switch (filter) {
    case 0:  d = v.downsample([](const std::array<unsigned char, 8>& n) {
                 return *std::max_element(n.begin(), n.end()); }); break;
    case 1:  d = v.downsample([](const std::array<unsigned char, 8>& n) {
                 unsigned s = 0; for (auto x : n) s += x; return (unsigned char)(s / 8); }); break;
    case 2:  d = v.downsample([](const std::array<unsigned char, 8>& n) {
                 unsigned s = 0; for (auto x : n) s += x; return (unsigned char)(s); }); break;
    default: d = v.downsample([](const std::array<unsigned char, 8>& n) {
                 return *std::min_element(n.begin(), n.end()); }); break;
    }
    std::memcpy(out, d.densities.data(), d.densities.size());
}

}  // extern "C"
