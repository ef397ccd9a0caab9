// pbs_plus_b200/csrc/corpus.cu -- K5: synthetic corpus generator (measurement aid).
//
// Generates the SURVEY.md section 8d corpora directly in HBM so that bench inputs are
// resident before the timed region.  Counter-based (stateless): every 8-byte word is
// a function of (seed, canonical block, word index), so any thread can produce any
// word.  The integer recipe is the product's own definition of the synthetic corpus;
// tests/ compare it bit-for-bit with the oracle's restatement (oracle/oracle.c).
// Not part of the reference (which has no benchmark corpus).
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.cuh"

namespace pbsgpu {

#define GOLD 0x9E3779B97F4A7C15ULL
__device__ __forceinline__ uint64_t fmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t canonical_block(const pbsgpu_corpus &c, uint64_t gblock) {
    uint64_t run = gblock / c.run_blocks, within = gblock % c.run_blocks;
    while (run > 0 && c.dup_permille > 0 &&
           fmix64((c.seed ^ 0xD1B54A32D192ED03ULL) + run * GOLD) % 1000u < c.dup_permille)
        run = fmix64((c.seed ^ 0x8CB92BA72F3D8DD7ULL) + run * GOLD) % run;
    return run * c.run_blocks + within;
}

__device__ __forceinline__ uint64_t edit_word(const pbsgpu_corpus &c, uint64_t gblock, uint64_t w, uint64_t v) {
    if (c.edit_mode == 1) {
        uint64_t gw = gblock * (c.block_len / 8) + w;
        uint64_t s0 = fmix64(c.edit_seed + gw * 3 * GOLD + 1);
        uint64_t s1 = fmix64(c.edit_seed + (gw * 3 + 1) * GOLD + 1);
        uint64_t nv = fmix64(c.edit_seed + (gw * 3 + 2) * GOLD + 1);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t lane = ((j < 4 ? s0 : s1) >> (16 * (j & 3))) & 0xFFFF;
            if (lane < c.edit_thresh16) {
                uint64_t m = 0xFFULL << (8 * j);
                v = (v & ~m) | (nv & m);
            }
        }
    } else if (c.edit_mode == 2) {
        uint64_t e = fmix64(c.edit_seed + gblock * GOLD + 7);
        if (e % 100u == 0) {
            uint64_t pos = (e >> 20) % c.block_len;
            if (pos / 8 == w) v ^= 0x5AULL << (8 * (pos & 7));
        }
    }
    return v;
}

// grid.y = file, grid.x = corpus block (duplicate granule) of that file; the block's seed is resolved
// once per thread (two 64-bit divisions and a hash chain), then the thread strides over the words.
__global__ void __launch_bounds__(256) k_corpus_fill(pbsgpu_corpus c, uint64_t first_file, uint8_t *dst, uint64_t stride) {
    const uint64_t words_per_block = c.block_len / 8;
    const uint64_t bpf = (c.file_len + c.block_len - 1) / c.block_len;
    const uint64_t file = first_file + blockIdx.y;
    uint8_t *out = dst + (uint64_t)blockIdx.y * stride;
    for (uint64_t bi = blockIdx.x; bi < bpf; bi += gridDim.x) {
        const uint64_t gblock = file * bpf + bi;
        const uint64_t cb = canonical_block(c, gblock);
        const uint64_t bseed = fmix64(c.seed * 0xA0761D6478BD642FULL + cb * 0xE7037ED1A0B428DBULL + 0x1234567ULL);
        const uint64_t block_off = bi * c.block_len;
        const uint64_t remain = c.file_len - block_off;
        const uint64_t nwords = remain >= c.block_len ? words_per_block : (remain + 7) / 8;
        for (uint64_t w = threadIdx.x; w < nwords; w += blockDim.x) {
            uint64_t v = fmix64(bseed + (w + 1) * GOLD);
            if (c.edit_mode) v = edit_word(c, gblock, w, v);
            const uint64_t byte_off = block_off + w * 8;
            if (byte_off + 8 <= c.file_len) {
                *(uint64_t *)(out + byte_off) = v;
            } else {
                for (uint64_t j = 0; byte_off + j < c.file_len; j++) out[byte_off + j] = (uint8_t)(v >> (8 * j));
            }
        }
    }
}

cudaError_t launch_corpus_fill(const pbsgpu_corpus &c, uint64_t first_file, uint32_t n_files, uint8_t *dst,
                               uint64_t stride, cudaStream_t st) {
    if (n_files == 0 || c.file_len == 0) return cudaSuccess;
    uint64_t bpf = (c.file_len + c.block_len - 1) / c.block_len;
    uint64_t bx = bpf > 65535 ? 65535 : bpf;
    for (uint32_t f0 = 0; f0 < n_files; f0 += 65535) {
        uint32_t nf = n_files - f0 < 65535 ? n_files - f0 : 65535;
        dim3 grid((unsigned)bx, nf);
        k_corpus_fill<<<grid, 256, 0, st>>>(c, first_file + f0, dst + (uint64_t)f0 * stride, stride);
    }
    return cudaGetLastError();
}

}  // namespace pbsgpu
