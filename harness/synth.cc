// synth.cc -- host-side synthetic strand generator (HARNESS input: tests, bench.py, golden generation;
// never linked into the product library, not part of parity).
//
// The reference's assets (share/styles/*.hair) are Git-LFS pointers in the
// checkout, so every workload is generated: seeded random-walk
// strands of the named shape.  The same buffer feeds the CPU oracle and the
// GPU path, so nothing here influences parity.
//
// RNG: the xorshift64 recurrence the reference uses for its strand shuffle
// (src/vkhr/scene_graph/hair_style.cc:679-685), one stream per strand seeded
// seed ^ 0x9E3779B97F4A7C15 * (strand + 1); u = (x >> 11) * 2^-53.
#include <cmath>
#include <cstdint>

namespace {
struct Rng {
    uint64_t x;
    explicit Rng(uint64_t s) : x(s ? s : 0x1234567887654321ull) {}
    double next() {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        return double(x >> 11) * (1.0 / 9007199254740992.0);
    }
};
}  // namespace

extern "C" {

// n_strands strands of segs_per_strand segments (segs+1 vertices each), written
// strand-major into xyz_out (n_strands * (segs+1) * 3 floats).
//   root      uniform in [root_min, root_max]
//   step k    p += seg_len * normalize(dir + curl * (u3 - 0.5) + gravity * (0,-1,0) + gather * toward_axis)
// `gather` pulls strands toward the vertical line through the root box centre
// (a ponytail-like band); 0 disables it.
// `first_strand`: strand ids first_strand .. first_strand + n_strands - 1 of the set (a rank generates only its shard).
int vkhr_harness_synth_strands(uint32_t n_strands, uint32_t segs_per_strand, uint64_t seed,
                            const float root_min[3], const float root_max[3],
                            float seg_len, float curl, float gravity, float gather, uint64_t first_strand, float* xyz_out) {
    if (!xyz_out || !root_min || !root_max || segs_per_strand == 0) return -1;
    const double cx = 0.5 * (double(root_min[0]) + root_max[0]);
    const double cz = 0.5 * (double(root_min[2]) + root_max[2]);
    const size_t vps = size_t(segs_per_strand) + 1;
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < (long long)n_strands; ++s) {
        Rng rng(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t(s) + first_strand + 1)));
        for (int w = 0; w < 4; ++w) rng.next();       // decorrelate neighbouring seeds
        double p[3], d[3];
        for (int c = 0; c < 3; ++c)
            p[c] = double(root_min[c]) + rng.next() * (double(root_max[c]) - double(root_min[c]));
        d[0] = rng.next() - 0.5; d[1] = -0.5 * rng.next(); d[2] = rng.next() - 0.5;
        float* out = xyz_out + size_t(s) * vps * 3;
        out[0] = float(p[0]); out[1] = float(p[1]); out[2] = float(p[2]);
        for (uint32_t k = 1; k <= segs_per_strand; ++k) {
            double n[3];
            n[0] = d[0] + curl * (rng.next() - 0.5) + gather * (cx - p[0]) * 0.05;
            n[1] = d[1] + curl * (rng.next() - 0.5) - gravity;
            n[2] = d[2] + curl * (rng.next() - 0.5) + gather * (cz - p[2]) * 0.05;
            double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (!(len > 1e-12)) { n[0] = 0; n[1] = -1; n[2] = 0; len = 1; }
            for (int c = 0; c < 3; ++c) {
                d[c] = n[c] / len;
                p[c] += double(seg_len) * d[c];
                out[3 * k + c] = float(p[c]);
            }
        }
    }
    return 0;
}

// Per-frame sway of config 5: vertex k of a strand moves by
// amplitude * (k/S)^2 * (sin(omega*t + phi_s), 0, cos(omega*t + phi_s)), phi_s from the strand id.
int vkhr_harness_synth_sway(const float* xyz_in, uint32_t n_strands, uint32_t segs_per_strand,
                         float t, float amplitude, float omega, float* xyz_out) {
    if (!xyz_in || !xyz_out || segs_per_strand == 0) return -1;
    const size_t vps = size_t(segs_per_strand) + 1;
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < (long long)n_strands; ++s) {
        const double phi = 6.283185307179586 * double((uint64_t(s) * 2654435761ull) & 0xFFFFu) / 65536.0;
        const double sx = std::sin(double(omega) * t + phi), cz = std::cos(double(omega) * t + phi);
        for (size_t k = 0; k < vps; ++k) {
            const double w = double(k) / double(segs_per_strand);
            const size_t i = (size_t(s) * vps + k) * 3;
            xyz_out[i + 0] = float(double(xyz_in[i + 0]) + double(amplitude) * w * w * sx);
            xyz_out[i + 1] = xyz_in[i + 1];
            xyz_out[i + 2] = float(double(xyz_in[i + 2]) + double(amplitude) * w * w * cz);
        }
    }
    return 0;
}

}  // extern "C"
