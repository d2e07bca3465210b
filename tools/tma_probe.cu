// tma_probe.cu -- which 3-D u8 tensor-map / TMA box-load variants work on this part (measurement tool, not product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool FROM_GLOBAL>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, int x, int y, int z, uint32_t bytes, uint8_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = s32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* m = FROM_GLOBAL ? gm : &pm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s32(sm)), "l"(m), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 22); ++spin)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(0) : "memory");
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? sm[i] : 0xEE;
}
#include <cstdlib>
int main(int argc, char** argv) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 64, H = 32, D = 16;
    std::vector<uint8_t> h(W * H * D);
    for (int i = 0; i < W * H * D; ++i) h[i] = (uint8_t)(1 + (i * 7 + (i >> 6) * 3 + (i >> 11) * 5) % 250);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint8_t* out; cudaMalloc(&out, 1 << 16);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap));
    struct V { int bx, by, bz; CUtensorMapL2promotion l2; int x, y, z; } vs[] = {
        {64, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -16, -3, -3}, {64, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 16, 21, 5},
        {64, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 48, -3, 9}, {64, 26, 26, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -16, -9, -9},
        {48, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, -3, -3}, {48, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_NONE, 29, 21, 5},
        {48, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_NONE, -3, 0, 0}, {48, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_NONE, -4, 0, 0}, {48, 14, 14, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4, 0, 0}};
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int vi = -1;
    for (auto& v : vs) if (++vi == only || only < 0) for (int from_global = 0; from_global < 2; ++from_global) {
        CUtensorMap m;
        const cuuint64_t gdim[3] = {W, H, D}, gstr[2] = {W, (cuuint64_t)W * H};
        const cuuint32_t box[3] = {(cuuint32_t)v.bx, (cuuint32_t)v.by, (cuuint32_t)v.bz}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, v.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box %dx%dx%d l2 %d: encode failed %d\n", v.bx, v.by, v.bz, (int)v.l2, (int)r); continue; }
        cudaMemcpy(gm, &m, sizeof m, cudaMemcpyHostToDevice);
        const uint32_t bytes = v.bx * v.by * v.bz;
        cudaMemset(out, 0xCD, 1 << 16);
        if (from_global) k<true><<<1, 128, bytes>>>(m, gm, v.x, v.y, v.z, bytes, out);
        else k<false><<<1, 128, bytes>>>(m, gm, v.x, v.y, v.z, bytes, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("box %dx%dx%d l2 %d at (%d,%d,%d) %s: %s\n", v.bx, v.by, v.bz, (int)v.l2, v.x, v.y, v.z, from_global ? "global" : "param", cudaGetErrorString(e)); return 2; }
        std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int c = 0; c < v.bz; ++c) for (int b = 0; b < v.by; ++b) for (int a = 0; a < v.bx; ++a) {
            const int X = v.x + a, Y = v.y + b, Z = v.z + c;
            const uint8_t want = (X < 0 || Y < 0 || Z < 0 || X >= W || Y >= H || Z >= D) ? 0 : h[X + Y * W + Z * W * H];
            bad += o[(c * v.by + b) * v.bx + a] != want;
        }
        printf("box %dx%dx%d l2 %d at (%d,%d,%d) %s: ok, %zu mismatches (first byte %u)\n", v.bx, v.by, v.bz, (int)v.l2, v.x, v.y, v.z, from_global ? "global" : "param", bad, o[0]);
    }
    return 0;
}
