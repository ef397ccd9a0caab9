// tools/tma_piece_probe.cu -- round-2 planning aid: does HBM stay efficient when every lane of a warp pulls its
// OWN 256 B piece per step (32 bulk copies, lane stride 2 KiB, consecutive pieces per lane over 8 steps)
// instead of one contiguous 8 KiB tile per warp?  (Would let K1 carry a lane's 64-entry ring across steps.)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred P1;\n W:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra D;\n bra W;\n D:\n }" ::"r"(bar), "r"(parity) : "memory");
}
template <int MODE>   // 0: one 8 KiB copy per warp step; 1: 32 x 256 B pieces, lane stride 2 KiB
__global__ void __launch_bounds__(256, 1) k_probe(const uint8_t *base, uint64_t bytes, uint32_t *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *buf0 = sm + warp * 2 * 8192;
    uint64_t *bars = (uint64_t *)(sm + 8 * 2 * 8192) + warp * 2;
    const uint32_t bar[2] = {s32(&bars[0]), s32(&bars[1])};
    if (lane == 0) { for (int i = 0; i < 2; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[i])); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint64_t SUPER = 64 * 1024;                                   // bytes per warp super-tile (8 steps x 8 KiB)
    const uint64_t nsuper = bytes / SUPER, gw = (uint64_t)blockIdx.x * 8 + warp, nw = (uint64_t)gridDim.x * 8;
    uint32_t acc = 0, uses[2] = {0, 0};
    auto issue = [&](uint64_t st, int step, int b) {
        const uint8_t *tile = base + st * SUPER;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[b]), "r"(8192) : "memory");
        __syncwarp();
        if (MODE == 0) {
            if (lane == 0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf0 + b * 8192)), "l"(tile + step * 8192), "r"(8192), "r"(bar[b]) : "memory");
        } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf0 + b * 8192 + lane * 256)), "l"(tile + lane * 2048 + step * 256), "r"(256), "r"(bar[b]) : "memory");
        }
    };
    for (uint64_t st = gw; st < nsuper; st += nw) {
        issue(st, 0, 0);
        for (int step = 0; step < 8; step++) {
            const int b = step & 1;
            if (step + 1 < 8) issue(st, step + 1, b ^ 1);
            bar_wait(bar[b], uses[b] & 1); uses[b]++;
            const uint4 *q = (const uint4 *)(buf0 + b * 8192 + lane * 256);
#pragma unroll
            for (int k = 0; k < 16; k++) { uint4 v = q[k]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
            __syncwarp();
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
int main() {
    const uint64_t bytes = 16ull << 30;
    uint8_t *d; uint32_t *sink; CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(d, 1, bytes));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int smem = 8 * 2 * 8192 + 8 * 2 * 8;
    CK(cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k_probe<0><<<p.multiProcessorCount, 256, smem>>>(d, bytes, sink);
            else k_probe<1><<<p.multiProcessorCount, 256, smem>>>(d, bytes, sink);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("mode %d (%s): %.2f ms, %.0f GB/s\n", mode, mode ? "32 x 256 B pieces per warp step, lane stride 2 KiB" : "one contiguous 8 KiB copy per warp step", ms, bytes / ms / 1e6);
        }
    }
    return 0;
}
