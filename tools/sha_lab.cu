// tools/sha_lab.cu -- K3 instruction-mix laboratory (measurement aid, not product).
//
// SHA-256 on CUDA cores is bound by the two integer pipes of an SM sub-partition (ALU: SHF/LOP3/IADD3/PRMT/SEL,
// FMA: IMAD; one warp instruction per 2 clk each, one issue per clk in total -- profiles/r01_microbench.txt).
// This program times variants of the one-lane-per-range kernel that differ ONLY in which pipe their additions and
// logical shifts are sent to and in the launch shape, on N equal ranges (no length imbalance), and checks every
// digest against variant 0 (which is checked against a host SHA-256).  `tools/sass_mix.py` counts the loop's SASS
// per pipe for the same variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/sha_lab.bin tools/sha_lab.cu
//   ./tools/sha_lab.bin [range_kib=256] [total_gib=16] [only_variant=-1]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e__), __LINE__); exit(1); } } while (0)

#define K256_LIST                                                                                          \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,        \
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,        \
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,        \
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,        \
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,        \
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,        \
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,        \
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2

struct Opq { uint32_t one, p29, p22; };

__device__ __forceinline__ uint32_t rotr(uint32_t x, int r) { return __funnelshift_r(x, x, r); }
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, const Opq &o) {   // a + b on the FMA pipe: IMAD R, R, UR(=1), R
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(o.one), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t fshr(uint32_t x, uint32_t pow2, const Opq &) {   // x >> n as mul.hi by 2^(32-n): IMAD.HI (half rate)
    uint32_t d; asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(pow2)); return d;
}

// variant bits: 1 = schedule additions on the FMA pipe, 2 = K+W on the FMA pipe, 4 = off-critical round additions
// (h+kw, d+h+kw) on the FMA pipe, 8 = ALL round additions on the FMA pipe (T = S1+ch shared), 16 = schedule shifts as IMAD.HI
template <int V>
__device__ __forceinline__ void compress(uint32_t (&st)[8], uint32_t (&w)[16], const Opq &o) {
    constexpr uint32_t K[64] = {K256_LIST};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ ((V & 16) ? fshr(w15, o.p29, o) : (w15 >> 3));
            const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ ((V & 16) ? fshr(w2, o.p22, o) : (w2 >> 10));
            if (V & 1) w[i & 15] = fadd(fadd(w[i & 15], s0, o), fadd(w[(i + 9) & 15], s1, o), o);
            else w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t kw = (V & 2) ? fadd(w[i & 15], K[i], o) : w[i & 15] + K[i];
        const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t en, an;
        if (V & 8) {
            const uint32_t hk = fadd(h, kw, o), dhk = fadd(d, hk, o), t = fadd(S1, ch, o);
            en = fadd(dhk, t, o);
            an = fadd(fadd(fadd(S0, mj, o), hk, o), t, o);
        } else if (V & 4) {
            const uint32_t hk = fadd(h, kw, o), dhk = fadd(d, hk, o);
            en = S1 + ch + dhk;                 // IADD3
            const uint32_t x = S0 + mj + hk;    // IADD3
            an = x + S1 + ch;                   // IADD3
        } else {
            const uint32_t t1 = h + S1 + ch + kw, t2 = S0 + mj;
            en = d + t1; an = t1 + t2;
        }
        h = g; g = f; f = e; e = en; d = c; c = b; b = a; a = an;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

__device__ __forceinline__ void sha_init(uint32_t (&s)[8]) {
    s[0] = 0x6a09e667; s[1] = 0xbb67ae85; s[2] = 0x3c6ef372; s[3] = 0xa54ff53a;
    s[4] = 0x510e527f; s[5] = 0x9b05688c; s[6] = 0x1f83d9ab; s[7] = 0x5be0cd19;
}

__device__ __noinline__ void finish(uint32_t (&s)[8], const uint8_t *tail, uint32_t rem, uint64_t len, uint8_t *out) {
    const uint64_t bits = len * 8;
    const int nblk = rem < 56 ? 1 : 2;
    Opq o{1u, 1u << 29, 1u << 22};
    for (int blk = 0; blk < nblk; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t idx = blk * 64 + i * 4 + k;
                uint32_t byte = idx < rem ? tail[idx] : (idx == rem ? 0x80u : 0u);
                v = (v << 8) | byte;
            }
            w[i] = v;
        }
        if (blk == nblk - 1) { w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits; }
        compress<0>(s, w, o);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s[i] >> 24); out[4 * i + 1] = (uint8_t)(s[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s[i] >> 8); out[4 * i + 3] = (uint8_t)s[i];
    }
}

// one lane per range; same load path as the product's k_sha_tuned (LDG.128 + prefetch + SEL/PRMT realign)
template <int V, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_lab(const uint8_t *base, const uint64_t *off, const uint64_t *len, uint32_t n,
                                                       uint8_t *digests, Opq o) {
    const uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= n) return;
    const uint8_t *p = base + off[t];
    const uint64_t L = len[t];
    uint32_t s[8];
    sha_init(s);
    const uint32_t nblk = (uint32_t)(L >> 6);
    const uint32_t delta = (uint32_t)((uintptr_t)p & 15), dw = delta >> 2, sh = delta & 3;
    const uint4 *q = (const uint4 *)(p - delta);
    const uint32_t sel = (sh + 3) | ((sh + 2) << 4) | ((sh + 1) << 8) | (sh << 12);
    const bool need5 = delta != 0, d1 = dw & 1, d2 = dw & 2;
    uint4 v0, v1, v2, v3, v4 = make_uint4(0, 0, 0, 0);
    if (nblk) { v0 = __ldg(q); v1 = __ldg(q + 1); v2 = __ldg(q + 2); v3 = __ldg(q + 3); if (need5) v4 = __ldg(q + 4); }
    for (uint32_t b = 0; b < nblk; b++) {
        uint32_t x[20] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w,
                          v3.x, v3.y, v3.z, v3.w, v4.x, v4.y, v4.z, v4.w};
        if (b + 1 < nblk) {
            const uint4 *qn = q + (uint64_t)(b + 1) * 4;
            v0 = need5 ? v4 : __ldg(qn);
            v1 = __ldg(qn + 1); v2 = __ldg(qn + 2); v3 = __ldg(qn + 3);
            if (need5) v4 = __ldg(qn + 4);
        }
        uint32_t y[18], w[16];
#pragma unroll
        for (int i = 0; i < 18; i++) y[i] = d2 ? x[i + 2] : x[i];
#pragma unroll
        for (int i = 0; i < 17; i++) y[i] = d1 ? y[i + 1] : y[i];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = __byte_perm(y[i], y[i + 1], sel);
        compress<V>(s, w, o);
    }
    finish(s, p + (uint64_t)nblk * 64, (uint32_t)(L & 63), L, digests + (uint64_t)t * 32);
}

__global__ void k_fill(uint64_t *d, uint64_t nwords, uint64_t seed) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + i * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; d[i] = z ^ (z >> 31);
    }
}

// ---- host SHA-256 (check of variant 0) ----
static void host_sha256(const uint8_t *msg, size_t len, uint8_t out[32]) {
    static const uint32_t K[64] = {K256_LIST};
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<uint8_t> m(msg, msg + len);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(((uint64_t)len * 8) >> (8 * i)));
    auto rr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
    for (size_t b = 0; b < m.size(); b += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (m[b + 4 * i] << 24) | (m[b + 4 * i + 1] << 16) | (m[b + 4 * i + 2] << 8) | m[b + 4 * i + 3];
        for (int i = 16; i < 64; i++) w[i] = w[i - 16] + (rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10));
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rr(e, 6) ^ rr(e, 11) ^ rr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rr(a, 2) ^ rr(a, 13) ^ rr(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
        }
        h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}

struct Variant { const char *name; void (*launch)(const uint8_t *, const uint64_t *, const uint64_t *, uint32_t, uint8_t *, Opq); };
template <int V, int THREADS, int MINB>
static void launch(const uint8_t *base, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *dig, Opq o) {
    k_lab<V, THREADS, MINB><<<(n + THREADS - 1) / THREADS, THREADS>>>(base, off, len, n, dig, o);
}

int main(int argc, char **argv) {
    const uint64_t rk = argc > 1 ? atoll(argv[1]) : 256, tg = argc > 2 ? atoll(argv[2]) : 16;
    const int only = argc > 3 ? atoi(argv[3]) : -1;
    const uint64_t total = tg << 30, R = rk << 10;
    const uint32_t n = (uint32_t)(total / R);
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs; %u ranges x %llu KiB (misaligned by 3)\n", prop.name, prop.multiProcessorCount, n, (unsigned long long)rk);
    uint8_t *buf, *dig, *dig0; uint64_t *d_off, *d_len;
    CK(cudaMalloc(&buf, total + 256)); CK(cudaMalloc(&dig, (size_t)n * 32)); CK(cudaMalloc(&dig0, (size_t)n * 32));
    CK(cudaMalloc(&d_off, n * 8)); CK(cudaMalloc(&d_len, n * 8));
    k_fill<<<4096, 256>>>((uint64_t *)buf, total / 8, 77); CK(cudaDeviceSynchronize());
    std::vector<uint64_t> off(n), len(n);
    for (uint32_t i = 0; i < n; i++) { off[i] = (uint64_t)i * R + 3; len[i] = R - 64 + 5; }
    CK(cudaMemcpy(d_off, off.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_len, len.data(), n * 8, cudaMemcpyHostToDevice));
    const Variant vs[] = {
        {"v0  plain C (ptxas decides)                 32 thr", launch<0, 32, 1>},
        {"v3  sched adds + K+W on FMA                 32 thr", launch<3, 32, 1>},
        {"v7  v3 + off-critical round adds on FMA     32 thr", launch<7, 32, 1>},
        {"v11 ALL adds on FMA (T shared)              32 thr", launch<11, 32, 1>},
        {"v23 v7 + sched shifts as IMAD.HI            32 thr", launch<23, 32, 1>},
        {"v27 v11 + sched shifts as IMAD.HI           32 thr", launch<27, 32, 1>},
        {"v16 only sched shifts as IMAD.HI (mode 2)   32 thr", launch<16, 32, 1>},
        {"v0                                          64 thr", launch<0, 64, 1>},
        {"v11                                         64 thr", launch<11, 64, 1>},
        {"v7                                          64 thr", launch<7, 64, 1>},
        {"v0                                         128 thr", launch<0, 128, 1>},
        {"v11                                        128 thr", launch<11, 128, 1>},
        {"v7                                         128 thr", launch<7, 128, 1>},
        {"v11 regs<=64 (128 thr x 8 blocks)          128 thr", launch<11, 128, 8>},
        {"v7  regs<=64 (128 thr x 8 blocks)          128 thr", launch<7, 128, 8>},
        {"v0  regs<=64 (128 thr x 8 blocks)          128 thr", launch<0, 128, 8>},
        {"v11 regs<=80 (128 thr x 6 blocks)          128 thr", launch<11, 128, 6>},
        {"v27 regs<=80 (128 thr x 6 blocks)          128 thr", launch<27, 128, 6>},
    };
    const int nv = (int)(sizeof vs / sizeof vs[0]);
    Opq o{1u, 1u << 29, 1u << 22};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint8_t> h0((size_t)n * 32), h1((size_t)n * 32);
    for (int v = 0; v < nv; v++) {
        if (only >= 0 && v != only && v != 0) continue;
        uint8_t *out = v == 0 ? dig0 : dig;
        CK(cudaMemset(out, 0, (size_t)n * 32));
        vs[v].launch(buf, d_off, d_len, n, out, o); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            vs[v].launch(buf, d_off, d_len, n, out, o);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        bool ok = true;
        if (v == 0) {
            CK(cudaMemcpy(h0.data(), dig0, (size_t)n * 32, cudaMemcpyDeviceToHost));
            std::vector<uint8_t> m(len[5]); uint8_t ref[32];
            CK(cudaMemcpy(m.data(), buf + off[5], len[5], cudaMemcpyDeviceToHost));
            host_sha256(m.data(), m.size(), ref);
            ok = memcmp(ref, &h0[5 * 32], 32) == 0;
        } else {
            CK(cudaMemcpy(h1.data(), dig, (size_t)n * 32, cudaMemcpyDeviceToHost));
            ok = memcmp(h0.data(), h1.data(), (size_t)n * 32) == 0;
        }
        const double bytes = (double)n * (double)len[0];
        printf("[%2d] %-52s %8.2f ms  %7.1f GB/s  %s\n", v, vs[v].name, best, bytes / best / 1e6, ok ? "ok" : "MISMATCH");
        fflush(stdout);
    }
    return 0;
}
