// Host emulation of the AO arithmetic of vkhr_b200/csrc/prefilter.cuh (test infrastructure, CPU only).
//
// tests/test_host.py extracts the device functions lao_at / lao_column from prefilter.cuh (the text between the tile
// constants and the Gaussian weight function), turns `__device__ __forceinline__` into `static inline`, writes it to
// LAO_EXTRACT and compiles this file with g++ -ffp-contract=off.  The rounding intrinsics become plain fp32
// operations, so the check is about the INDEXING of the register-tiled column form: for random float tiles and row
// flags, every output of lao_column (and of lao_column_rolled, its period-unrolled form) must be bit-identical to
// lao_at at the same voxel, for every instantiated tap
// offset pair, tile halo and number of valid outputs.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }

namespace vkhr_b200 {
#include LAO_EXTRACT
}
using namespace vkhr_b200;

// the host's tap table (vkhr_b200.cu: axis_taps)
static AxisTaps axis_taps(float r, bool positive) {
    AxisTaps a;
    const float fl = std::floor(r), fp = r - fl;
    const int ifl = (int)fl;
    if (positive) { a.o0 = ifl; a.o1 = ifl + 1; a.w0 = 1.0f - fp; a.w1 = fp; }
    else if (fp != 0.0f) { a.o0 = -ifl - 1; a.o1 = -ifl; a.w0 = fp; a.w1 = 1.0f - fp; }
    else { a.o0 = -ifl; a.o1 = -ifl + 1; a.w0 = 1.0f; a.w1 = 0.0f; }
    return a;
}

template <int NO0, int PO0, int TZ>
static int run1(float radius, int h, double fill, int n_out) {
    PrefilterArgs A{};
    A.neg = axis_taps(radius, false); A.pos = axis_taps(radius, true); A.ao_max = 0.16f; A.ao_exponent = 10.0f;
    if (A.neg.o0 != NO0 || A.pos.o0 != PO0) { std::printf("offset mismatch r=%g: %d %d\n", radius, A.neg.o0, A.pos.o0); return 1; }
    const int BY = kPfTY + 2 * h, BZ = TZ + 2 * h, FX = kPfTX + 2 * h;
    std::vector<float> ftile((size_t)FX * BY * BZ);
    std::vector<uint32_t> flag((size_t)BY * BZ, 0u);
    for (int r = 0; r < BY * BZ; ++r) {
        const bool rowfill = (std::rand() / (double)RAND_MAX) < fill;
        for (int x = 0; x < FX; ++x) {
            const int b = rowfill && (std::rand() % 3 == 0) ? std::rand() % 256 : 0;
            ftile[(size_t)r * FX + x] = (float)b / 255.0f;
            if (b) flag[r] = 1u;
        }
    }
    bool tile_any = false;
    for (uint32_t f : flag) tile_any |= f != 0u;
    const float ao_empty = lao_at(A, [](int, int, int) -> float { return 0.0f; });
    int bad = 0, nontrivial = 0;
    for (int jy = 0; jy < kPfTY; ++jy)
        for (int lane = 0; lane < 32; ++lane) {
            float got[TZ];
            for (int k = 0; k < TZ; ++k) got[k] = -777.0f;
            float rolled[TZ];
            for (int k = 0; k < TZ; ++k) rolled[k] = -777.0f;
            lao_column<NO0, PO0, TZ>(A, ftile.data() + (h * BY + (jy + h)) * FX + (lane + h), flag.data() + h * BY + (jy + h), BY * FX, FX, BY,
                                 tile_any, ao_empty, n_out, [&](int kz, float r) { got[kz] = r; });
            lao_column_rolled<NO0, PO0, TZ>(A, ftile.data() + (h * BY + (jy + h)) * FX + (lane + h), flag.data() + h * BY + (jy + h), BY * FX, FX, BY,
                                        tile_any, ao_empty, n_out, [&](int kz, float r) { rolled[kz] = r; });
            if (std::memcmp(got, rolled, sizeof got)) { if (bad < 5) std::printf("r=%g jy=%d lane=%d: rolled form differs\n", radius, jy, lane); ++bad; }
            for (int kz = 0; kz < TZ; ++kz) {
                const float* centre = ftile.data() + ((kz + h) * BY + (jy + h)) * FX + (lane + h);
                const float want = lao_at(A, [&](int ox, int oy, int oz) -> float { return centre[(oz * BY + oy) * FX + ox]; });
                if (kz >= n_out) { if (got[kz] != -777.0f) ++bad; continue; }       // outputs beyond the grid are not emitted
                if (want != ao_empty) ++nontrivial;
                if (std::memcmp(&want, &got[kz], 4)) {
                    if (bad < 5) std::printf("r=%g jy=%d lane=%d kz=%d want %.9g got %.9g\n", radius, jy, lane, kz, want, got[kz]);
                    ++bad;
                }
            }
        }
    std::printf("r=%g (%d,%d) TZ=%d halo=%d fill=%.2f n_out=%d: bad=%d nontrivial=%d\n", radius, NO0, PO0, TZ, h, fill, n_out, bad, nontrivial);
    return bad;
}

// n_out <= 0: all outputs of the tile; otherwise that many (a tile cut by the end of the grid)
template <int NO0, int PO0>
static int run(float radius, int h, double fill, int n_out) {
    return run1<NO0, PO0, kPfTZ>(radius, h, fill, n_out > 0 ? n_out : kPfTZ) +
           run1<NO0, PO0, kPfTZDeep>(radius, h, fill, n_out > 0 ? n_out + 6 : kPfTZDeep);
}

int main() {
    int bad = 0;
    std::srand(12345);
    for (double fill : {1.0, 0.3, 0.05, 0.0}) {
        bad += run<0, 0>(0.0f, 1, fill, 0);
        bad += run<-1, 0>(0.5f, 1, fill, 0);
        bad += run<-1, 0>(0.5f, 4, fill, 0);
        bad += run<-1, 1>(1.0f, 2, fill, 0);
        bad += run<-2, 1>(1.3f, 2, fill, 0);
        bad += run<-2, 2>(2.0f, 3, fill, 5);
        bad += run<-3, 2>(2.5f, 3, fill, 0);
        bad += run<-3, 2>(2.5f, 4, fill, 3);
        bad += run<-3, 3>(3.0f, 4, fill, 0);
        bad += run<-4, 3>(3.75f, 4, fill, 1);
        bad += run<-4, 4>(4.0f, 5, fill, 0);
    }
    std::printf("TOTAL bad=%d\n", bad);
    return bad != 0;
}
