// microbench3.cu -- shared-memory atomic throughput on B200 (is a shared-memory slab histogram
// a faster sink than the global `red` walk?).  Each thread adds packed-byte increments (1 << 8 * (i & 3)) to
// 32-bit words of a shared slab, the operation a z-slab-binned walk would issue once per sample.
//   pattern 0: word = hash(thread, i) over the slab (strand samples of a binned slab land like this)
//   pattern 1: lane-consecutive words (conflict-free best case)
//   pattern 2: strand-like: lanes 1.2 voxels apart along y in a 256 x 256 x 2 byte slab (x-fastest)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int PATTERN>
__global__ void __launch_bounds__(256) k_atoms(uint32_t slab_words, uint32_t iters, uint32_t* sink) {
    extern __shared__ uint32_t slab[];
    for (uint32_t i = threadIdx.x; i < slab_words; i += blockDim.x) slab[i] = 0;
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = 0; i < iters; ++i) {
        uint32_t w, sh;
        if (PATTERN == 0) { x = x * 1664525u + 1013904223u; w = (x >> 8) % slab_words; sh = (x >> 3) & 24u; }
        else if (PATTERN == 1) { w = (warp * 32u + lane + i * 256u) % slab_words; sh = (i & 3u) * 8u; }
        else {
            x = x * 1664525u + 1013904223u;
            const uint32_t x0 = (x >> 8) & 255u, y0 = (x >> 16) & 127u, z = (x >> 24) & 1u;
            const uint32_t y = (y0 + (lane * 6u) / 5u) & 255u, xx = (x0 + (lane >> 2)) & 255u;
            const uint32_t idx = (z * 256u + y) * 256u + xx;
            w = (idx >> 2) % slab_words; sh = (idx & 3u) * 8u;
        }
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slab + w)), "r"(1u << sh) : "memory");
    }
    __syncthreads();
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i < slab_words; i += blockDim.x) s += slab[i];
    if (s == 0xFFFFFFFFu) sink[0] = s;
}

template <int PATTERN>
void run(const char* name, uint32_t slab_bytes, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_atoms<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes);
    uint32_t* sink; cudaMalloc(&sink, 4);
    const uint32_t iters = 4096;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) k_atoms<PATTERN><<<grid, 256, slab_bytes>>>(slab_bytes / 4, iters, sink);
    cudaEventRecord(a);
    for (int rep = 0; rep < 5; ++rep) k_atoms<PATTERN><<<grid, 256, slab_bytes>>>(slab_bytes / 4, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double ops = (double)grid * 256 * iters;
    printf("{\"pattern\": \"%s\", \"slab_bytes\": %u, \"ctas_per_sm\": %d, \"ms\": %.4f, \"G_atomics_per_s\": %.1f, \"err\": \"%s\"}\n",
           name, slab_bytes, ctas_per_sm, ms, ops / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    for (int c : {1, 2, 4}) {
        const uint32_t slab = c == 1 ? 131072u : c == 2 ? 98304u : 49152u;
        run<0>("random", slab, c);
        run<1>("lane-consecutive", slab, c);
        run<2>("strand-like 256x256x2 slab", slab, c);
    }
    return 0;
}
