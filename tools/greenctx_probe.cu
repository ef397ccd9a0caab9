// tools/greenctx_probe.cu -- can runtime-API kernels be confined to SM partitions via green contexts here?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <set>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define CU(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { printf("driver error %d at line %d (%s)\n", (int)r, __LINE__, #x); return 1; } } while (0)

__global__ void k_smid(unsigned *out, long long spin) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if (threadIdx.x == 0) out[blockIdx.x] = smid;
}
template <typename T> static T ep(const char *name) {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return (T)p;
}
int main() {
    CK(cudaSetDevice(0)); CK(cudaFree(0));
    auto getRes = ep<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto split = ep<CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
    auto genDesc = ep<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned)>("cuDevResourceGenerateDesc");
    auto gcreate = ep<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
    auto gstream = ep<CUresult (*)(CUstream *, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
    auto devget = ep<CUresult (*)(CUdevice *, int)>("cuDeviceGet");
    if (!getRes || !split || !genDesc || !gcreate || !gstream || !devget) { printf("green ctx entry points missing\n"); return 1; }
    CUdevice dev; CU(devget(&dev, 0));
    CUdevResource all; CU(getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("total SMs %u\n", all.sm.smCount);
    CUdevResource grp[1], rest; unsigned n = 1;
    CU(split(grp, &n, &all, &rest, 0, 24));
    printf("groups %u: A=%u SMs, remaining=%u SMs\n", n, grp[0].sm.smCount, rest.sm.smCount);
    CUdevResourceDesc dA, dB; CU(genDesc(&dA, &grp[0], 1)); CU(genDesc(&dB, &rest, 1));
    CUgreenCtx gA, gB; CU(gcreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM)); CU(gcreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream sA, sB; CU(gstream(&sA, gA, CU_STREAM_NON_BLOCKING, 0)); CU(gstream(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
    unsigned *oA, *oB; const int NB = 2000;
    CK(cudaMalloc(&oA, NB * 4)); CK(cudaMalloc(&oB, NB * 4));
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    k_smid<<<NB, 64, 0, (cudaStream_t)sA>>>(oA, 20000); CK(cudaGetLastError());
    k_smid<<<NB, 64, 0, (cudaStream_t)sB>>>(oB, 20000); CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned hA[NB], hB[NB]; CK(cudaMemcpy(hA, oA, NB * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hB, oB, NB * 4, cudaMemcpyDeviceToHost));
    std::set<unsigned> a(hA, hA + NB), b(hB, hB + NB); int inter = 0; for (auto x : a) inter += b.count(x);
    printf("stream A used %zu distinct SMs, stream B used %zu, overlap %d\n", a.size(), b.size(), inter);
    // event interop between a normal stream, green streams
    cudaStream_t s0; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    CK(cudaEventRecord(e0, s0)); CK(cudaStreamWaitEvent((cudaStream_t)sA, e0, 0));
    k_smid<<<48, 64, 0, (cudaStream_t)sA>>>(oA, 2000000000LL / 10); CK(cudaEventRecord(e1, (cudaStream_t)sA));
    k_smid<<<NB, 64, 0, (cudaStream_t)sB>>>(oB, 2000000); CK(cudaEventRecord(e2, (cudaStream_t)sB));
    CK(cudaStreamWaitEvent(s0, e1, 0)); CK(cudaStreamWaitEvent(s0, e2, 0)); CK(cudaStreamSynchronize(s0));
    float ms1, ms2; cudaEventElapsedTime(&ms1, e0, e1); cudaEventElapsedTime(&ms2, e0, e2);
    printf("concurrent: A long kernel %.1f ms, B many-block kernel %.1f ms (B should not wait for A)\n", ms1, ms2);
    printf("GREENCTX OK\n");
    return 0;
}
