// selftest.cu -- device self tests of the walk's numerics building blocks (HARNESS: tests only, never linked
// into the product library).  Includes the product header vkhr_b200/csrc/walk.cuh and is compiled with the
// product's numerics flags (harness/__init__.py), so the functions under test are the ones the kernels inline.
#include "../vkhr_b200/csrc/walk.cuh"

#include <cstdint>
#include <cuda_runtime.h>

using namespace vkhr_b200;

namespace {

// Bitwise comparison of div_exact against the IEEE division on pseudo-random operands.
__global__ void __launch_bounds__(256)
k_selftest_division(float d, float y, uint64_t seed, uint32_t per_thread, unsigned long long* mismatches) {
    uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1));
    uint32_t bad = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        // random sign and mantissa, exponent spread over [2^-70, 2^70] (both sides of the guard)
        const uint32_t e = 127u - 70u + (uint32_t)((x >> 40) % 141u);
        const float a = __uint_as_float(((uint32_t)x & 0x807FFFFFu) | (e << 23));
        const float q = div_exact(a, d, y);
        const float want = __fdiv_rn(a, d);
        bad += (__float_as_uint(q) != __float_as_uint(want));
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// EVERY float in [lo_bits, hi_bits] (bit patterns of positive floats): rcp_steps(s) -- the reciprocal the walk feeds to
// its FMA division of `direction /= steps` -- against the correctly rounded __frcp_rn(s), and the division of three
// numerators per divisor (|a| <= s, as |direction| <= steps) against __fdiv_rn.
__global__ void __launch_bounds__(256)
k_selftest_rcp(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* mismatches) {
    uint32_t bad = 0;
    const uint64_t n = (uint64_t)hi_bits - lo_bits + 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float s = __uint_as_float(lo_bits + (uint32_t)i);
        const float y = rcp_steps(s);
        bad += (__float_as_uint(y) != __float_as_uint(__frcp_rn(s)));
        uint64_t x = 0x9E3779B97F4A7C15ull * (i + 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            // |a| = s * u with u in (0, 1]: the magnitude range of a direction component
            const float u = __uint_as_float(0x3F800000u - (uint32_t)((x >> 20) % 0x0C000000u));   // 2^-24 .. 1
            float a = __fmul_rn(s, u);
            if (k == 2) a = s;                                                                    // the major axis: exactly +-1
            if (x & 1u) a = -a;
            if (div_fast_ok(a)) bad += (__float_as_uint(div_fast(a, s, y)) != __float_as_uint(__fdiv_rn(a, s)));
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

int run(int device, unsigned long long* out, void (*launch)(unsigned long long*, void*), void* arg) {
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    unsigned long long* d_bad = nullptr;
    if (cudaMalloc(&d_bad, 8) != cudaSuccess) return -3;
    cudaMemset(d_bad, 0, 8);
    launch(d_bad, arg);
    unsigned long long bad = ~0ull;
    const cudaError_t e = cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    *out = bad;
    return 0;
}

}  // namespace

extern "C" {

__attribute__((visibility("default")))
int vkhr_selftest_division(int device, float divisor, uint64_t n_trials, uint64_t seed, uint64_t* mismatches) {
    if (!mismatches || !(divisor > 0.0f) || divisor != divisor || divisor > 3.0e38f) return -1;
    struct Arg { float d, y; uint64_t seed, n; } arg{divisor, 0.0f, seed, n_trials};
    arg.y = (divisor >= 9.094947e-13f && divisor <= 1.0995116e12f) ? 1.0f / divisor : 0.0f;   // as make_grid (vkhr_b200.cu)
    unsigned long long bad = 0;
    const int rc = run(device, &bad, [](unsigned long long* d_bad, void* p) {
        const Arg& a = *static_cast<Arg*>(p);
        const unsigned blocks = 148 * 16, threads = 256;
        const uint64_t per = (a.n + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
        k_selftest_division<<<blocks, threads>>>(a.d, a.y, a.seed, (uint32_t)per, d_bad);
    }, &arg);
    if (rc == 0) *mismatches = bad;
    return rc;
}

// All floats of [lo, hi] (positive, finite, lo <= hi).
__attribute__((visibility("default")))
int vkhr_selftest_rcp(int device, float lo, float hi, uint64_t* mismatches) {
    if (!mismatches || !(lo > 0.0f) || !(hi >= lo) || hi > 3.0e38f) return -1;
    struct Arg { uint32_t lo, hi; } arg;
    static_assert(sizeof(float) == 4, "");
    __builtin_memcpy(&arg.lo, &lo, 4);
    __builtin_memcpy(&arg.hi, &hi, 4);
    unsigned long long bad = 0;
    const int rc = run(device, &bad, [](unsigned long long* d_bad, void* p) {
        const Arg& a = *static_cast<Arg*>(p);
        k_selftest_rcp<<<148 * 16, 256>>>(a.lo, a.hi, d_bad);
    }, &arg);
    if (rc == 0) *mismatches = bad;
    return rc;
}

}  // extern "C"
