// vkhr_b200_adapter.hh -- the C++ drop-in for the reference's voxelisation surface.
//
// Host side of the boundary, written in the reference's own language (C++17) against the
// reference's own headers (compile with -I<vkhr>/include -I<vkhr>/foreign/glm).  It has the exact
// shape of the two member functions it replaces
//     vkhr::HairStyle::Volume HairStyle::voxelize_segments(size_t w, size_t h, size_t d) const   hair_style.hh:104, hair_style.cc:296-342
//     vkhr::HairStyle::Volume HairStyle::voxelize_vertices(size_t w, size_t h, size_t d) const   hair_style.hh:103, hair_style.cc:257-294
//     void HairStyle::Volume::normalize()                                                         hair_style.cc:344-357
// and fills the same `Volume` (resolution as a glm::vec3 of floats, bounds = get_bounding_box(), x-fastest u8
// densities, i8vec4 tangents) so that the only caller, vulkan::HairStyle::load (src/vkhr/rasterizer/hair_style.cc:75,:77),
// switches with a one-line change:
//     strand_volume = vkhr_b200::voxelize_segments(*hair_style, 256, 256, 256);   // was hair_style->voxelize_segments(256, 256, 256)
// It only marshals pointers into the C ABI of include/vkhr_b200.h; all work happens in libvkhr_b200.so on the GPU.
// There is no CPU fallback: if no sm_100 device is usable the functions throw std::runtime_error.
#pragma once
#include <vkhr/scene_graph/hair_style.hh>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../include/vkhr_b200.h"

namespace vkhr_b200 {

// One lazily created context per process and device (the reference calls the voxeliser from its main thread only).
class Context {
public:
    explicit Context(int device = 0) {
        if (vkhr_b200_create(device, &ctx_) != VKHR_B200_OK)
            throw std::runtime_error(std::string("vkhr_b200_create: ") + vkhr_b200_last_error(nullptr));
    }
    ~Context() { vkhr_b200_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    vkhr_b200_ctx* get() const { return ctx_; }
    static Context& instance() { static Context c{0}; return c; }
private:
    vkhr_b200_ctx* ctx_ = nullptr;
};

inline void check(vkhr_b200_ctx* ctx, int rc, const char* what) {
    if (rc != VKHR_B200_OK) throw std::runtime_error(std::string(what) + ": " + vkhr_b200_last_error(ctx));
}

// Strands of equal length whose index buffer is the one HairStyle::generate_indices (hair_style.cc:196-213) emits
// need no index buffer on the GPU.  Returns segments per strand, or 0 when the explicit indices must be used.
inline std::uint32_t uniform_segments(const vkhr::HairStyle& hs) {
    if (hs.segments.size() != 0) return 0;                                   // per-strand counts: keep the indices
    const std::uint32_t s = hs.get_default_segment_count();
    const std::size_t v = hs.vertices.size();
    if (s == 0 || v % (s + 1) != 0 || hs.indices.size() != 2 * (v / (s + 1)) * s) return 0;
    std::size_t i = 0;
    for (std::size_t strand = 0, first = 0; strand < v / (s + 1); ++strand, first += s + 1)
        for (std::uint32_t k = 0; k < s; ++k, i += 2)
            if (hs.indices[i] != first + k || hs.indices[i + 1] != first + k + 1) return 0;
    return s;
}

namespace detail {
inline vkhr::HairStyle::Volume make_volume(const vkhr::HairStyle& hs, std::size_t w, std::size_t h, std::size_t d, bool tangents) {
    vkhr::HairStyle::Volume volume;
    volume.resolution = glm::vec3 { w, h, d };                               // hair_style.cc:297-304
    volume.bounds = hs.get_bounding_box();                                    // hair_style.cc:306
    volume.densities.resize(w * h * d);
    if (tangents) volume.tangents.resize(w * h * d);
    return volume;
}
}  // namespace detail

// HairStyle::voxelize_segments on the B200.  `flags`: VKHR_B200_* (0 = the reference's behaviour).
inline vkhr::HairStyle::Volume voxelize_segments(const vkhr::HairStyle& hs, std::size_t w, std::size_t h, std::size_t d,
                                                 std::uint32_t flags = 0, bool want_tangents = true, Context& c = Context::instance()) {
    const bool tangents = want_tangents && hs.tangents.size() == hs.vertices.size();
    vkhr::HairStyle::Volume volume = detail::make_volume(hs, w, h, d, tangents);
    const std::uint32_t segs = uniform_segments(hs);
    static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::i8vec4) == 4, "GLM types must be tightly packed");
    check(c.get(), vkhr_b200_voxelize_segments(c.get(),
              reinterpret_cast<const float*>(hs.vertices.data()), static_cast<std::uint32_t>(hs.vertices.size()),
              segs ? nullptr : hs.indices.data(), segs ? 0 : hs.indices.size(), segs,
              tangents ? reinterpret_cast<const float*>(hs.tangents.data()) : nullptr,
              &volume.bounds.origin.x, &volume.bounds.size.x,
              static_cast<std::uint32_t>(w), static_cast<std::uint32_t>(h), static_cast<std::uint32_t>(d), flags,
              volume.densities.data(), tangents ? reinterpret_cast<std::int8_t*>(volume.tangents.data()) : nullptr),
          "vkhr_b200_voxelize_segments");
    return volume;
}

// HairStyle::voxelize_vertices on the B200.
inline vkhr::HairStyle::Volume voxelize_vertices(const vkhr::HairStyle& hs, std::size_t w, std::size_t h, std::size_t d,
                                                 std::uint32_t flags = 0, bool want_tangents = true, Context& c = Context::instance()) {
    const bool tangents = want_tangents && hs.tangents.size() == hs.vertices.size();
    vkhr::HairStyle::Volume volume = detail::make_volume(hs, w, h, d, tangents);
    check(c.get(), vkhr_b200_voxelize_vertices(c.get(),
              reinterpret_cast<const float*>(hs.vertices.data()), static_cast<std::uint32_t>(hs.vertices.size()),
              tangents ? reinterpret_cast<const float*>(hs.tangents.data()) : nullptr,
              &volume.bounds.origin.x, &volume.bounds.size.x,
              static_cast<std::uint32_t>(w), static_cast<std::uint32_t>(h), static_cast<std::uint32_t>(d), flags,
              volume.densities.data(), tangents ? reinterpret_cast<std::int8_t*>(volume.tangents.data()) : nullptr),
          "vkhr_b200_voxelize_vertices");
    return volume;
}

// Volume::normalize on the B200 (in place).
inline void normalize(vkhr::HairStyle::Volume& volume, Context& c = Context::instance()) {
    check(c.get(), vkhr_b200_normalize(c.get(), volume.densities.data(), volume.densities.size()), "vkhr_b200_normalize");
}

}  // namespace vkhr_b200
