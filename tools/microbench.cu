// tools/microbench.cu -- B200 micro-benchmarks that size the design of the voxeliser:
// global RED/ATOM throughput on L2-resident vs HBM-resident grids, shared-memory atomics,
// fill and copy bandwidth.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints one JSON object.  Measurement tool only; not part of the product path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// MODE 0: red (no return)  1: atom (return consumed)  2: atom, packed byte add + overflow test
// LOCAL: 0 = uniform random address; 1 = a warp's 32 lanes hit addresses inside a 4 KB window (strand-like locality)
template <int MODE, int LOCAL>
__global__ void k_atomics(uint32_t* grid, uint32_t mask, uint32_t per_thread, uint32_t* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        uint32_t h;
        if (LOCAL) {
            uint32_t base = hash32((tid >> 5) * 977u + i * 0x9E3779B9u) & mask & ~1023u;
            h = base + (hash32(tid + i * 31u) & 1023u);
        } else h = hash32(tid * 0x9E3779B9u + i * 0x85EBCA6Bu) & mask;
        if (MODE == 0) atomicAdd(grid + h, 1u);
        else if (MODE == 1) acc ^= atomicAdd(grid + h, 1u);
        else { uint32_t sh = (h & 3u) * 8u; uint32_t old = atomicAdd(grid + (h >> 2), 1u << sh); if (((old >> sh) & 0xFF) == 0xFF) acc |= 1u; }
    }
    if (acc == 0xDEADBEEF) *sink = acc;
}

// Atomic throughput when only `active` of the 32 lanes of each warp take part (divergent sample loops).
__global__ void k_atomics_partial(uint32_t* grid, uint32_t mask, uint32_t per_thread, uint32_t active, uint32_t* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((threadIdx.x & 31u) >= active) return;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        uint32_t h = hash32(tid * 0x9E3779B9u + i * 0x85EBCA6Bu) & mask;
        uint32_t sh = (h & 3u) * 8u; uint32_t old = atomicAdd(grid + (h >> 2), 1u << sh); if (((old >> sh) & 0xFF) == 0xFF) acc |= 1u;
    }
    if (acc == 0xDEADBEEF) *sink = acc;
}

// Same total work, but each atomic is separated by `pad` dependent FMAs (models the per-sample arithmetic).
__global__ void k_atomics_padded(uint32_t* grid, uint32_t mask, uint32_t per_thread, uint32_t pad, uint32_t* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0; float f = (float)tid;
    for (uint32_t i = 0; i < per_thread; ++i) {
        for (uint32_t k = 0; k < pad; ++k) f = f * 1.0000001f + 0.5f;
        uint32_t h = hash32(tid * 0x9E3779B9u + i * 0x85EBCA6Bu + (uint32_t)f) & mask;
        uint32_t sh = (h & 3u) * 8u; uint32_t old = atomicAdd(grid + (h >> 2), 1u << sh); if (((old >> sh) & 0xFF) == 0xFF) acc |= 1u;
    }
    if (acc == 0xDEADBEEF) *sink = acc;
}

// Sector locality: the 32 lanes of a warp instruction hit 32/K random 32-byte sectors, K lanes per sector
// (distinct words for K <= 8; K = 16/32 put 2/4 lanes on each word, different bytes).
__global__ void k_atomics_grouped(uint32_t* grid, uint32_t mask_sectors, uint32_t per_thread, uint32_t K, uint32_t* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lane = threadIdx.x & 31u, warp = tid >> 5;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        uint32_t sector = hash32((warp * 32u + lane / K) * 0x9E3779B9u + i * 0x85EBCA6Bu) & mask_sectors;
        uint32_t slot = lane % K;                       // position inside the sector
        uint32_t word = sector * 8u + (slot & 7u);
        uint32_t sh = ((slot >> 3) & 3u) * 8u;
        uint32_t old = atomicAdd(grid + word, 1u << sh);
        if (((old >> sh) & 0xFF) == 0xFF) acc |= 1u;
    }
    if (acc == 0xDEADBEEF) *sink = acc;
}

__global__ void k_smem_atomics(uint32_t per_thread, uint32_t words, uint32_t* sink) {
    extern __shared__ uint32_t s[];
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = 0; i < per_thread; ++i) {
        uint32_t h = hash32(tid * 0x9E3779B9u + i * 0x85EBCA6Bu) % words;
        atomicAdd(s + h, 1u);
    }
    __syncthreads();
    if (s[threadIdx.x % words] == 0xDEADBEEF) *sink = 1;
}

__global__ void k_fill(uint4* p, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
}
__global__ void k_copy(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        b[i] = a[i];
}
__global__ void k_empty() {}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    uint32_t* grid; uint32_t* sink; uint4* big2;
    const size_t GB = 1ull << 30;
    CK(cudaMalloc(&grid, GB)); CK(cudaMalloc(&big2, GB)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(grid, 0, GB));
    printf("{\n");
    const int threads = 256, blocks = 148 * 64;           // 2.4 M threads
    const uint32_t per = 16;                               // 38.8 M atomics per launch
    const double nops = (double)threads * blocks * per;
    struct { const char* name; size_t bytes; } sizes[] = {{"4MiB", 4u << 20}, {"16MiB", 16u << 20}, {"64MiB", 64u << 20}, {"256MiB", 256u << 20}, {"1GiB", GB}};
    for (auto& sz : sizes) {
        uint32_t mask = (uint32_t)(sz.bytes / 4 - 1);
        float r0 = time_ms([&] { k_atomics<0, 0><<<blocks, threads>>>(grid, mask, per, sink); }, 5);
        float r1 = time_ms([&] { k_atomics<1, 0><<<blocks, threads>>>(grid, mask, per, sink); }, 5);
        float r2 = time_ms([&] { k_atomics<2, 0><<<blocks, threads>>>(grid, mask, per, sink); }, 5);
        float l0 = time_ms([&] { k_atomics<0, 1><<<blocks, threads>>>(grid, mask, per, sink); }, 5);
        float l2 = time_ms([&] { k_atomics<2, 1><<<blocks, threads>>>(grid, mask, per, sink); }, 5);
        printf(" \"atomics_%s\": {\"red_random_Gops\": %.2f, \"atom_random_Gops\": %.2f, \"atom_packed8_random_Gops\": %.2f, \"red_local_Gops\": %.2f, \"atom_packed8_local_Gops\": %.2f},\n",
               sz.name, nops / r0 / 1e6, nops / r1 / 1e6, nops / r2 / 1e6, nops / l0 / 1e6, nops / l2 / 1e6);
    }
    {
        uint32_t mask = (uint32_t)((16u << 20) - 1);   // 16 MiB of packed bytes
        for (uint32_t active : {32u, 24u, 16u, 8u, 4u, 1u}) {
            float r = time_ms([&] { k_atomics_partial<<<blocks, threads>>>(grid, mask, per, active, sink); }, 5);
            printf(" \"atom_packed8_active%u_of_32\": {\"Gops\": %.2f, \"Ginstr\": %.3f},\n", active, nops * active / 32 / r / 1e6, nops / 32 / r / 1e6);
        }
        for (uint32_t K : {1u, 2u, 4u, 8u, 16u, 32u}) {
            float r = time_ms([&] { k_atomics_grouped<<<blocks, threads>>>(grid, (16u << 20) / 32 - 1, per, K, sink); }, 5);
            printf(" \"atom_packed8_%u_lanes_per_sector\": {\"Gops\": %.2f},\n", K, nops / r / 1e6);
        }
        for (uint32_t pad : {0u, 32u}) {
            float r = time_ms([&] { k_atomics_padded<<<blocks, threads>>>(grid, mask, per, pad, sink); }, 5);
            printf(" \"atom_packed8_pad%u_fma\": {\"Gops\": %.2f},\n", pad, nops / r / 1e6);
        }
        for (int bl : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
            float r = time_ms([&] { k_atomics<2, 0><<<bl, threads>>>(grid, mask, per * 4, sink); }, 5);
            printf(" \"atom_packed8_blocks%d\": {\"Gops\": %.2f},\n", bl, (double)threads * bl * per * 4 / r / 1e6);
        }
    }
    {
        const uint32_t words = 8192;   // 32 KB
        float s = time_ms([&] { k_smem_atomics<<<148 * 6, 256, words * 4>>>(256, words, sink); }, 5);
        printf(" \"smem_atomics_32KB\": {\"Gops\": %.2f},\n", (double)148 * 6 * 256 * 256 / s / 1e6);
    }
    for (auto& sz : sizes) {
        float f = time_ms([&] { k_fill<<<148 * 8, 256>>>((uint4*)grid, sz.bytes / 16); }, 10);
        float m = time_ms([&] { cudaMemsetAsync(grid, 0, sz.bytes); }, 10);
        printf(" \"fill_%s\": {\"kernel_us\": %.2f, \"kernel_GBs\": %.1f, \"memset_us\": %.2f},\n", sz.name, f * 1e3, sz.bytes / f / 1e6, m * 1e3);
    }
    {
        float c = time_ms([&] { k_copy<<<148 * 16, 256>>>((const uint4*)grid, big2, GB / 16); }, 5);
        printf(" \"copy_1GiB\": {\"GBs_read_plus_write\": %.1f},\n", 2.0 * GB / c / 1e6);
        float e = time_ms([&] { k_empty<<<1, 32>>>(); }, 200);
        printf(" \"empty_kernel_back_to_back_us\": %.2f\n", e * 1e3);
    }
    printf("}\n");
    return 0;
}
