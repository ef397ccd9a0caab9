/*
 * oracle/oracle.c -- CPU restatement of the reference's chunk + digest + probe path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product (pbs_plus_b200/, libpbsgpu.so) never links, imports or calls it.
 *
 * PARITY UNPINNED.  The reference (pbs-plus @ 26d6969) holds none of this
 * arithmetic: it calls the un-vendored Go module github.com/pbs-plus/pxar v0.19.2
 * (go.mod:28) at
 *     internal/pxarmount/commit.go:302-305   buzhash.NewConfig(4096)
 *     internal/pxarmount/commit.go:329       transfer.NewRemoteDedupSplitArchiveWriter
 *     internal/pxarmount/commit.go:720,:858  writer.WriteEntryReader(entry, reader, size)
 *     internal/pxarmount/commit.go:286-294,:324-329  previous-index digest seed
 * and no reference test holds a golden vector for chunk boundaries or chunk
 * digests.  What is restated here is therefore
 * the published upstream algorithm that module implements:
 *   - Proxmox Backup Server `pbs-datastore/src/chunker.rs` (ChunkerImpl::new,
 *     ::scan, ::shall_break) for the buzhash content-defined chunker,
 *   - FIPS 180-4 for SHA-256 (pinned by the NIST known-answer vectors and by
 *     Python hashlib / OpenSSL in tests/test_oracle.py -- the digest half of the
 *     oracle IS pinned),
 *   - plain set semantics for the known-chunk probe ("Only new chunks are
 *     uploaded", docs/pxar-mount.md:105).
 *
 * Everything is plain C11 (+ optional x86 SHA-NI, which is what Go's
 * crypto/sha256 uses on amd64 and therefore what a fair CPU baseline must use).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

#include "buzhash_table.h"

#define ORC_WINDOW 64u

/* ------------------------------------------------------------------------- */
/* a1: chunker parameters  (commit.go:302-305 -> buzhash.NewConfig)           */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t avg;       /* average chunk size in BYTES, power of two           */
    uint32_t min;       /* avg >> 2                                            */
    uint32_t max;       /* avg << 2                                            */
    uint32_t mask;      /* break_test_mask    = 2*avg - 1                      */
    uint32_t break_min; /* break_test_minimum = mask - 2                       */
    uint32_t window;    /* 64                                                  */
    uint32_t table[256];
} orc_cfg;

/* Upstream ChunkerImpl::new(chunk_size_avg): avg must be a power of two;
 * min = avg>>2, max = avg<<2, break_test_mask = avg*2-1, minimum = mask-2.
 * `table` NULL selects the default table.  Returns 0, or -22 (EINVAL). */
int orc_config(uint32_t avg_bytes, const uint32_t *table, orc_cfg *out) {
    if (!out || avg_bytes < 256u || avg_bytes > (1u << 29) || (avg_bytes & (avg_bytes - 1)))
        return -22;
    out->avg = avg_bytes;
    out->min = avg_bytes >> 2;
    out->max = avg_bytes << 2;
    out->mask = avg_bytes * 2u - 1u;
    out->break_min = out->mask - 2u;
    out->window = ORC_WINDOW;
    memcpy(out->table, table ? table : ORC_BUZHASH_TABLE, sizeof out->table);
    return 0;
}

/* commit.go:303 passes 4096.  Upstream proxmox-backup-client's --chunk-size is
 * in KiB with default 4096 (= 4 MiB), which is also what BASELINE.json's
 * "4 MiB avg chunk" says; this helper is that interpretation. */
int orc_config_kib(uint32_t avg_kib, const uint32_t *table, orc_cfg *out) {
    if (avg_kib == 0 || avg_kib > (1u << 19)) return -22;
    return orc_config(avg_kib << 10, table, out);
}

const uint32_t *orc_default_table(void) { return ORC_BUZHASH_TABLE; }

static inline uint32_t rotl32(uint32_t x, unsigned r) { return (x << (r & 31)) | (x >> ((32 - r) & 31)); }

/* ------------------------------------------------------------------------- */
/* a2: streaming chunker -- statement-for-statement restatement of upstream  */
/*     ChunkerImpl::scan / shall_break                                        */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t h;
    uint32_t window_size;
    uint64_t chunk_size;
    uint8_t window[ORC_WINDOW];
} orc_chunker;

void orc_chunker_reset(orc_chunker *c) { memset(c, 0, sizeof *c); }

static inline int orc_shall_break(const orc_cfg *cfg, const orc_chunker *c) {
    if (c->chunk_size >= cfg->max) return 1;
    if (c->chunk_size < cfg->min) return 0;
    return (c->h & cfg->mask) >= cfg->break_min;
}

/* Returns the offset just after the cut inside data[0..len), or 0 if no cut was
 * found (state is carried to the next call).  On a cut the state is reset. */
uint64_t orc_chunker_scan(const orc_cfg *cfg, orc_chunker *c, const uint8_t *data, uint64_t len) {
    uint64_t pos = 0;
    if (c->window_size < ORC_WINDOW) {
        uint64_t need = ORC_WINDOW - c->window_size;
        uint64_t copy_len = need < len ? need : len;
        for (uint64_t i = 0; i < copy_len; i++) {
            uint8_t byte = data[pos];
            c->window[c->window_size] = byte;
            c->h = rotl32(c->h, 1) ^ cfg->table[byte];
            pos++;
            c->window_size++;
        }
        c->chunk_size += copy_len;
        if (c->window_size < ORC_WINDOW) return 0;
    }
    uint32_t idx = (uint32_t)(c->chunk_size & 0x3f);
    while (pos < len) {
        uint8_t enter = data[pos];
        uint8_t leave = c->window[idx];
        /* window is 64 = 2*32 bytes, so the leaving term's rotation is the identity */
        c->h = rotl32(c->h, 1) ^ cfg->table[leave] ^ cfg->table[enter];
        c->chunk_size++;
        pos++;
        c->window[idx] = enter;
        if (orc_shall_break(cfg, c)) {
            c->h = 0;
            c->chunk_size = 0;
            c->window_size = 0;
            return pos;
        }
        idx = (uint32_t)(c->chunk_size & 0x3f);
    }
    return 0;
}

/* Whole-buffer form: chunk END offsets (exclusive) of data[0..len), feeding the
 * streaming chunker `feed` bytes at a time (feed==0: all at once).  The final
 * partial chunk is emitted at EOF regardless of min.  Returns the number of
 * chunks (which may exceed cap; only cap are stored). */
uint64_t orc_chunk_buffer(const orc_cfg *cfg, const uint8_t *data, uint64_t len, uint64_t feed,
                          uint64_t *ends, uint64_t cap) {
    orc_chunker c;
    orc_chunker_reset(&c);
    uint64_t n = 0, pos = 0, last = 0;
    if (feed == 0) feed = len ? len : 1;
    while (pos < len) {
        uint64_t piece_end = pos + feed < len ? pos + feed : len;
        while (pos < piece_end) {
            uint64_t r = orc_chunker_scan(cfg, &c, data + pos, piece_end - pos);
            if (r == 0) { pos = piece_end; break; }
            pos += r;
            if (n < cap) ends[n] = pos;
            n++;
            last = pos;
        }
    }
    if (last < len) {
        if (n < cap) ends[n] = len;
        n++;
    }
    return n;
}

/* Position-independent closed form used to cross-check the rolling recurrence:
 * H(i) = XOR_{j=0..63} rotl32(T[b[i-j]], j mod 32), i >= 63. */
uint32_t orc_window_hash(const uint32_t *table, const uint8_t *win64_ending_at_i) {
    uint32_t h = 0;
    for (unsigned j = 0; j < ORC_WINDOW; j++) h ^= rotl32(table[win64_ending_at_i[63 - j]], j & 31);
    return h;
}

/* Independent second implementation of the cut rule from the closed form:
 * next cut length L = smallest L in [min',max] with L==max or test(H(start+L-1)),
 * where min' = max(min, window+1): scan() only tests after the window has been
 * filled (64 bytes, untested) AND rolled once, so a chunk of exactly 64 bytes is
 * never cut by the hash test.  Only matters for avg == 256 (min == 64). */
uint64_t orc_chunk_buffer_closed_form(const orc_cfg *cfg, const uint8_t *data, uint64_t len,
                                      uint64_t *ends, uint64_t cap) {
    uint64_t n = 0, start = 0;
    while (start < len) {
        uint64_t end = len;
        uint64_t min_eff = cfg->min > ORC_WINDOW ? cfg->min : ORC_WINDOW + 1;
        for (uint64_t L = min_eff; start + L <= len; L++) {
            if (L >= cfg->max) { end = start + L; break; }
            uint32_t h = orc_window_hash(cfg->table, data + start + L - 64);
            if ((h & cfg->mask) >= cfg->break_min) { end = start + L; break; }
        }
        if (n < cap) ends[n] = end;
        n++;
        start = end;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* a3: SHA-256 (FIPS 180-4)                                                   */
/* ------------------------------------------------------------------------- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static const uint32_t H256_INIT[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                      0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

static inline uint32_t rotr32(uint32_t x, unsigned r) { return (x >> r) | (x << (32 - r)); }
static inline uint32_t be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

static void sha256_blocks_portable(uint32_t st[8], const uint8_t *p, uint64_t nblk) {
    uint32_t w[64];
    while (nblk--) {
        for (int i = 0; i < 16; i++) w[i] = be32(p + 4 * i);
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + K256[i] + w[i];
            uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
        p += 64;
    }
}

#if defined(__x86_64__)
__attribute__((target("sha,sse4.1,ssse3")))
static void sha256_blocks_shani(uint32_t st[8], const uint8_t *p, uint64_t nblk) {
    const __m128i BSWAP = _mm_set_epi64x(0x0c0d0e0f08090a0bLL, 0x0405060700010203LL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&st[0]);   /* DCBA */
    __m128i s1 = _mm_loadu_si128((const __m128i *)&st[4]);    /* HGFE */
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                       /* CDAB */
    s1 = _mm_shuffle_epi32(s1, 0x1B);                         /* EFGH */
    __m128i s0 = _mm_alignr_epi8(tmp, s1, 8);                 /* ABEF */
    s1 = _mm_blend_epi16(s1, tmp, 0xF0);                      /* CDGH */
    while (nblk--) {
        __m128i save0 = s0, save1 = s1, m[4], msg;
        for (int i = 0; i < 16; i++) {
            if (i < 4) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16 * i)), BSWAP);
            msg = _mm_add_epi32(m[i & 3], _mm_loadu_si128((const __m128i *)&K256[4 * i]));
            s1 = _mm_sha256rnds2_epu32(s1, s0, msg);
            if (i >= 3 && i < 15) {
                __m128i t = _mm_alignr_epi8(m[i & 3], m[(i - 1) & 3], 4);
                m[(i + 1) & 3] = _mm_add_epi32(m[(i + 1) & 3], t);
                m[(i + 1) & 3] = _mm_sha256msg2_epu32(m[(i + 1) & 3], m[i & 3]);
            }
            msg = _mm_shuffle_epi32(msg, 0x0E);
            s0 = _mm_sha256rnds2_epu32(s0, s1, msg);
            if (i >= 1 && i < 13) m[(i - 1) & 3] = _mm_sha256msg1_epu32(m[(i - 1) & 3], m[i & 3]);
        }
        s0 = _mm_add_epi32(s0, save0);
        s1 = _mm_add_epi32(s1, save1);
        p += 64;
    }
    tmp = _mm_shuffle_epi32(s0, 0x1B);       /* FEBA */
    s1 = _mm_shuffle_epi32(s1, 0xB1);        /* DCHG */
    s0 = _mm_blend_epi16(tmp, s1, 0xF0);     /* DCBA */
    s1 = _mm_alignr_epi8(s1, tmp, 8);        /* HGFE */
    _mm_storeu_si128((__m128i *)&st[0], s0);
    _mm_storeu_si128((__m128i *)&st[4], s1);
}
#endif

static int g_shani = -1; /* -1 unknown, 0 no, 1 yes */
static int g_force_portable = 0;

int orc_have_shani(void) {
    if (g_shani < 0) {
#if defined(__x86_64__)
        unsigned a, b, c, d;
        g_shani = (__get_cpuid_count(7, 0, &a, &b, &c, &d) && (b & (1u << 29))) ? 1 : 0;
#else
        g_shani = 0;
#endif
    }
    return g_shani;
}
void orc_force_portable_sha(int on) { g_force_portable = on; }

static void sha256_blocks(uint32_t st[8], const uint8_t *p, uint64_t nblk) {
#if defined(__x86_64__)
    if (!g_force_portable && orc_have_shani()) { sha256_blocks_shani(st, p, nblk); return; }
#endif
    sha256_blocks_portable(st, p, nblk);
}

void orc_sha256(const uint8_t *data, uint64_t len, uint8_t out[32]) {
    uint32_t st[8];
    memcpy(st, H256_INIT, sizeof st);
    uint64_t nblk = len / 64;
    sha256_blocks(st, data, nblk);
    uint8_t tail[128];
    uint64_t rem = len - nblk * 64;
    memset(tail, 0, sizeof tail);
    memcpy(tail, data + nblk * 64, rem);
    tail[rem] = 0x80;
    unsigned tl = rem < 56 ? 64 : 128;
    uint64_t bits = len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_blocks(st, tail, tl / 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}

/* ------------------------------------------------------------------------- */
/* a2+a3: stream -> chunks -> digests  (what WriteEntryReader does per stream) */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t stream;
    uint32_t flags;      /* bit0: digest was already known (dedup hit)          */
    uint64_t end_off;    /* exclusive end offset of the chunk within its stream */
    uint8_t digest[32];
} orc_chunk;

uint64_t orc_chunk_digest(const orc_cfg *cfg, uint32_t stream, const uint8_t *data, uint64_t len,
                          orc_chunk *out, uint64_t cap) {
    orc_chunker c;
    orc_chunker_reset(&c);
    uint64_t n = 0, pos = 0, start = 0;
    while (pos < len) {
        uint64_t r = orc_chunker_scan(cfg, &c, data + pos, len - pos);
        uint64_t end = r ? pos + r : len;
        if (n < cap) {
            out[n].stream = stream; out[n].flags = 0; out[n].end_off = end;
            orc_sha256(data + start, end - start, out[n].digest);
        }
        n++;
        start = end; pos = end;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* a2 caveat: SUGGESTED BOUNDARIES.  Newer upstream PBS wraps the chunker of
 * the archive PAYLOAD stream in a `PayloadChunker` (pbs-datastore/src/chunker.rs) that may also cut
 * at "suggested boundaries" = the offsets at which a file's PAYLOAD header starts in the ppxar
 * stream (what the reference's writer produces per file at commit.go:720 / pxarfs.go:408-411).
 * Upstream's rule, per chunk starting at `base`: a pending boundary B is dropped when
 * B - base < min; when min <= B - base <= max the chunk is cut AT B unless the hash test cuts
 * earlier; when B - base > max the plain chunker decides and B stays pending.  (Upstream's outcome
 * additionally depends on how many unscanned bytes its async reader happens to hold when B becomes
 * known; restated here in the limit of byte-wise arrival, the only timing-independent reading.)
 * Whether pbs-plus/pxar v0.19.2 does this is UNVERIFIED -- it is an optional input everywhere.
 * forced[] must be strictly increasing stream offsets in (0, len). */
uint64_t orc_chunk_digest_forced(const orc_cfg *cfg, uint32_t stream, const uint8_t *data, uint64_t len,
                                 const uint64_t *forced, uint64_t n_forced, orc_chunk *out, uint64_t cap) {
    orc_chunker c;
    orc_chunker_reset(&c);
    uint64_t n = 0, base = 0, fi = 0;
    while (base < len) {
        while (fi < n_forced && (forced[fi] <= base || forced[fi] - base < cfg->min)) fi++;   /* past or too small: ignored */
        uint64_t limit = len;            /* bytes the plain chunker may look at for this chunk */
        int at_boundary = 0;
        if (fi < n_forced && forced[fi] < len && forced[fi] - base <= cfg->max) { limit = forced[fi]; at_boundary = 1; }
        uint64_t r = orc_chunker_scan(cfg, &c, data + base, limit - base);
        uint64_t end;
        if (r) { end = base + r; if (at_boundary && end == limit) fi++; }
        else if (at_boundary) { end = limit; orc_chunker_reset(&c); fi++; }
        else end = len;
        if (n < cap) {
            out[n].stream = stream; out[n].flags = 0; out[n].end_off = end;
            orc_sha256(data + base, end - base, out[n].digest);
        }
        n++;
        base = end;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* a4: known-digest set (open addressing, exact 32-byte compare)              */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t cap;   /* power of two */
    uint64_t count;
    uint8_t *keys;  /* cap * 32 */
    uint8_t *used;  /* cap */
} orc_set;

orc_set *orc_set_create(uint64_t min_capacity) {
    orc_set *s = (orc_set *)calloc(1, sizeof *s);
    if (!s) return NULL;
    uint64_t cap = 64;
    while (cap < min_capacity * 2) cap <<= 1;
    s->cap = cap;
    s->keys = (uint8_t *)malloc(cap * 32);
    s->used = (uint8_t *)calloc(cap, 1);
    if (!s->keys || !s->used) { free(s->keys); free(s->used); free(s); return NULL; }
    return s;
}
void orc_set_destroy(orc_set *s) { if (s) { free(s->keys); free(s->used); free(s); } }
uint64_t orc_set_count(const orc_set *s) { return s->count; }

static int orc_set_find(const orc_set *s, const uint8_t *d, uint64_t *slot) {
    uint64_t h; memcpy(&h, d, 8);
    uint64_t i = h & (s->cap - 1);
    while (s->used[i]) {
        if (memcmp(s->keys + i * 32, d, 32) == 0) { *slot = i; return 1; }
        i = (i + 1) & (s->cap - 1);
    }
    *slot = i;
    return 0;
}
static void orc_set_grow(orc_set *s) {
    orc_set *n = orc_set_create(s->cap); /* doubles */
    for (uint64_t i = 0; i < s->cap; i++)
        if (s->used[i]) { uint64_t sl; orc_set_find(n, s->keys + i * 32, &sl); n->used[sl] = 1; memcpy(n->keys + sl * 32, s->keys + i * 32, 32); n->count++; }
    free(s->keys); free(s->used);
    *s = *n; free(n);
}
/* hit[i] = 1 if digest i was already present (including earlier in this call). */
void orc_set_probe_insert(orc_set *s, const uint8_t *d32, uint64_t n, uint8_t *hit, int insert) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t sl;
        int f = orc_set_find(s, d32 + i * 32, &sl);
        if (hit) hit[i] = (uint8_t)f;
        if (!f && insert) {
            s->used[sl] = 1; memcpy(s->keys + sl * 32, d32 + i * 32, 32); s->count++;
            if (s->count * 2 > s->cap) orc_set_grow(s);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Synthetic corpus -- identical integer recipe on the  */
/* device (pbs_plus_b200/csrc/corpus.cu); the two are compared bit-for-bit.    */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t seed;
    uint64_t file_len;      /* every file has this length                              */
    uint64_t block_len;     /* duplicate granularity, multiple of 8 (4 MiB at full size) */
    uint32_t run_blocks;    /* blocks per duplicate run (>=1)                          */
    uint32_t dup_permille;  /* probability a run copies an earlier run, 0..1000        */
    uint32_t edit_mode;     /* 0 none, 1 uniform byte edits, 2 one edited byte in ~1% of blocks */
    uint32_t edit_thresh16; /* mode 1: byte edited iff its 16-bit lane < thresh (655 ~ 1%) */
    uint64_t edit_seed;
} orc_corpus;

static inline uint64_t fmix64(uint64_t z) {   /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define GOLD 0x9E3779B97F4A7C15ULL

static inline uint64_t corpus_blocks_per_file(const orc_corpus *c) {
    return (c->file_len + c->block_len - 1) / c->block_len;
}
static uint64_t corpus_canonical_block(const orc_corpus *c, uint64_t gblock) {
    uint64_t run = gblock / c->run_blocks, within = gblock % c->run_blocks;
    while (run > 0 && c->dup_permille > 0 &&
           fmix64((c->seed ^ 0xD1B54A32D192ED03ULL) + run * GOLD) % 1000u < c->dup_permille)
        run = fmix64((c->seed ^ 0x8CB92BA72F3D8DD7ULL) + run * GOLD) % run;
    return run * c->run_blocks + within;
}
/* 8-byte little-endian word `w` (index within block) of canonical block cb */
static inline uint64_t corpus_word(const orc_corpus *c, uint64_t cb, uint64_t w) {
    uint64_t bseed = fmix64(c->seed * 0xA0761D6478BD642FULL + cb * 0xE7037ED1A0B428DBULL + 0x1234567ULL);
    return fmix64(bseed + (w + 1) * GOLD);
}
static inline uint64_t corpus_edit_word(const orc_corpus *c, uint64_t file_id, uint64_t gblock,
                                        uint64_t word_in_block, uint64_t v) {
    if (c->edit_mode == 1) {
        uint64_t gw = gblock * (c->block_len / 8) + word_in_block;
        uint64_t s0 = fmix64(c->edit_seed + gw * 3 * GOLD + 1);
        uint64_t s1 = fmix64(c->edit_seed + (gw * 3 + 1) * GOLD + 1);
        uint64_t nv = fmix64(c->edit_seed + (gw * 3 + 2) * GOLD + 1);
        for (int j = 0; j < 8; j++) {
            uint64_t lane = ((j < 4 ? s0 : s1) >> (16 * (j & 3))) & 0xFFFF;
            if (lane < c->edit_thresh16) {
                uint64_t m = 0xFFULL << (8 * j);
                v = (v & ~m) | (nv & m);
            }
        }
    } else if (c->edit_mode == 2) {
        uint64_t e = fmix64(c->edit_seed + gblock * GOLD + 7);
        if (e % 100u == 0) {
            uint64_t pos = (e >> 20) % c->block_len;
            if (pos / 8 == word_in_block) v ^= 0x5AULL << (8 * (pos & 7));
        }
    }
    (void)file_id;
    return v;
}

/* Fill out[0..n) with bytes [off, off+n) of file `file_id`. */
void orc_corpus_fill(const orc_corpus *c, uint64_t file_id, uint64_t off, uint8_t *out, uint64_t n) {
    uint64_t bpf = corpus_blocks_per_file(c);
    uint64_t pos = off, endp = off + n;
    uint64_t cur_block = UINT64_MAX, cb = 0, gblock = 0;
    while (pos < endp) {
        uint64_t bi = pos / c->block_len;
        if (bi != cur_block) { cur_block = bi; gblock = file_id * bpf + bi; cb = corpus_canonical_block(c, gblock); }
        uint64_t inb = pos - bi * c->block_len;
        uint64_t w = inb / 8;
        uint64_t v = corpus_word(c, cb, w);
        if (c->edit_mode) v = corpus_edit_word(c, file_id, gblock, w, v);
        unsigned b0 = (unsigned)(inb & 7);
        for (unsigned j = b0; j < 8 && pos < endp && pos < (bi + 1) * c->block_len; j++, pos++)
            out[pos - off] = (uint8_t)(v >> (8 * j));
    }
}

/* ------------------------------------------------------------------------- */
/* Multi-threaded driver: one stream per task, host threads pull tasks.       */
/* This is the "reference CPU implementation on all host cores" leg.          */
/* ------------------------------------------------------------------------- */
typedef struct {
    const orc_cfg *cfg;
    const uint8_t *const *ptrs;
    const uint64_t *lens;
    uint32_t n;
    orc_chunk *out;           /* n * cap_per_stream                     */
    uint64_t cap_per_stream;
    uint64_t *n_out;          /* per stream                             */
    volatile uint32_t next;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        j->n_out[i] = orc_chunk_digest(j->cfg, i, j->ptrs[i], j->lens[i], j->out + (uint64_t)i * j->cap_per_stream,
                                       j->cap_per_stream);
    }
    return NULL;
}

int orc_chunk_digest_mt(const orc_cfg *cfg, const uint8_t *const *ptrs, const uint64_t *lens, uint32_t n,
                        uint32_t threads, orc_chunk *out, uint64_t cap_per_stream, uint64_t *n_out) {
    mt_job j = {cfg, ptrs, lens, n, out, cap_per_stream, n_out, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (uint32_t t = 0; t < threads; t++)
        if (pthread_create(&th[t], NULL, mt_worker, &j)) return -11;
    for (uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}

typedef struct { const orc_corpus *c; uint8_t *const *ptrs; uint64_t first_file; uint32_t n; volatile uint32_t next; } fill_job;
static void *fill_worker(void *arg) {
    fill_job *j = (fill_job *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        orc_corpus_fill(j->c, j->first_file + i, 0, j->ptrs[i], j->c->file_len);
    }
    return NULL;
}
int orc_corpus_fill_mt(const orc_corpus *c, uint64_t first_file, uint8_t *const *ptrs, uint32_t n, uint32_t threads) {
    fill_job j = {c, ptrs, first_file, n, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (uint32_t t = 0; t < threads; t++)
        if (pthread_create(&th[t], NULL, fill_worker, &j)) return -11;
    for (uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* Whole corpus files without materialising the corpus: every worker generates one file at a time into its
 * own buffer and chunks + digests it (the full-size cfg2 / cfg5 parity checks: 1024 x 64 MiB would
 * otherwise need 64 GiB of host memory).  out: n_files * cap_per_file records, n_out per file. */
typedef struct {
    const orc_cfg *cfg; const orc_corpus *c; uint64_t first_file; uint32_t n;
    orc_chunk *out; uint64_t cap; uint64_t *n_out; volatile uint32_t next; volatile int oom;
} cgen_job;
static void *cgen_worker(void *arg) {
    cgen_job *j = (cgen_job *)arg;
    uint8_t *buf = (uint8_t *)malloc(j->c->file_len ? j->c->file_len : 1);
    if (!buf) { j->oom = 1; return NULL; }
    for (;;) {
        uint32_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        orc_corpus_fill(j->c, j->first_file + i, 0, buf, j->c->file_len);
        j->n_out[i] = orc_chunk_digest(j->cfg, i, buf, j->c->file_len, j->out + (uint64_t)i * j->cap, j->cap);
    }
    free(buf);
    return NULL;
}
int orc_corpus_chunk_digest_mt(const orc_cfg *cfg, const orc_corpus *c, uint64_t first_file, uint32_t n_files,
                               uint32_t threads, orc_chunk *out, uint64_t cap_per_file, uint64_t *n_out) {
    cgen_job j = {cfg, c, first_file, n_files, out, cap_per_file, n_out, 0, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (uint32_t t = 0; t < threads; t++)
        if (pthread_create(&th[t], NULL, cgen_worker, &j)) return -11;
    for (uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return j.oom ? -12 : 0;
}

/* ====================================================================================================
 * XXH3-64 (seed 0, default secret) -- SURVEY.md section 8 row f2: the per-file content hash the commit
 * walk computes through xxh3.New()/Sum64() (reference internal/pxarmount/commit.go:717-725) and
 * re-computes in verifyBackedFileHashes (commit.go:957-976).  The Go module github.com/zeebo/xxh3
 * (go.mod) implements the published XXH3 algorithm (xxHash v0.8 specification); this is a scalar
 * restatement of that specification.  PINNED: tests/test_oracle.py checks it against the independent
 * python-xxhash binding (libxxhash) for every length 0..2100 and random long inputs.
 * ==================================================================================================== */
static const uint8_t XXH3_SECRET[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
const uint8_t *orc_xxh3_secret(void) { return XXH3_SECRET; }

#define XP32_1 0x9E3779B1u
#define XP32_2 0x85EBCA77u
#define XP32_3 0xC2B2AE3Du
#define XP64_1 0x9E3779B185EBCA87ull
#define XP64_2 0xC2B2AE3D27D4EB4Full
#define XP64_3 0x165667B19E3779F9ull
#define XP64_4 0x85EBCA77C2B2AE63ull
#define XP64_5 0x27D4EB2F165667C5ull
#define XPMX1 0x165667919E3779F9ull
#define XPMX2 0x9FB21C651E98DF25ull

static inline uint64_t x_le64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }   /* little-endian host */
static inline uint32_t x_le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t x_rotl64(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t x_fold(uint64_t a, uint64_t b) { __uint128_t m = (__uint128_t)a * b; return (uint64_t)m ^ (uint64_t)(m >> 64); }
static inline uint64_t x_aval3(uint64_t h) { h ^= h >> 37; h *= XPMX1; h ^= h >> 32; return h; }
static inline uint64_t x_aval64(uint64_t h) { h ^= h >> 33; h *= XP64_2; h ^= h >> 29; h *= XP64_3; h ^= h >> 32; return h; }
static inline uint64_t x_mix16(const uint8_t *in, const uint8_t *sec) { return x_fold(x_le64(in) ^ x_le64(sec), x_le64(in + 8) ^ x_le64(sec + 8)); }
static inline void x_acc512(uint64_t acc[8], const uint8_t *in, const uint8_t *sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t dv = x_le64(in + 8 * i), dk = dv ^ x_le64(sec + 8 * i);
        acc[i ^ 1] += dv;
        acc[i] += (uint64_t)(uint32_t)dk * (dk >> 32);
    }
}

uint64_t orc_xxh3_64(const uint8_t *in, uint64_t len) {
    const uint8_t *S = XXH3_SECRET;
    if (len == 0) return x_aval64(x_le64(S + 56) ^ x_le64(S + 64));
    if (len <= 3) {
        uint32_t comb = ((uint32_t)in[0] << 16) | ((uint32_t)in[len >> 1] << 24) | in[len - 1] | ((uint32_t)len << 8);
        return x_aval64((uint64_t)comb ^ (uint64_t)(x_le32(S) ^ x_le32(S + 4)));
    }
    if (len <= 8) {
        uint64_t flip = x_le64(S + 8) ^ x_le64(S + 16);
        uint64_t h = ((uint64_t)x_le32(in + len - 4) + ((uint64_t)x_le32(in) << 32)) ^ flip;
        h ^= x_rotl64(h, 49) ^ x_rotl64(h, 24); h *= XPMX2; h ^= (h >> 35) + len; h *= XPMX2;
        return h ^ (h >> 28);
    }
    if (len <= 16) {
        uint64_t lo = x_le64(in) ^ (x_le64(S + 24) ^ x_le64(S + 32)), hi = x_le64(in + len - 8) ^ (x_le64(S + 40) ^ x_le64(S + 48));
        return x_aval3(len + __builtin_bswap64(lo) + hi + x_fold(lo, hi));
    }
    if (len <= 128) {
        uint64_t acc = len * XP64_1;
        for (int i = (int)((len - 1) / 32); i >= 0; i--) {
            acc += x_mix16(in + 16 * i, S + 32 * i);
            acc += x_mix16(in + len - 16 * (i + 1), S + 32 * i + 16);
        }
        return x_aval3(acc);
    }
    if (len <= 240) {
        uint64_t acc = len * XP64_1;
        for (int i = 0; i < 8; i++) acc += x_mix16(in + 16 * i, S + 16 * i);
        uint64_t end = x_mix16(in + len - 16, S + 136 - 17);
        acc = x_aval3(acc);
        for (unsigned i = 8; i < len / 16; i++) end += x_mix16(in + 16 * i, S + 16 * (i - 8) + 3);
        return x_aval3(acc + end);
    }
    uint64_t acc[8] = {XP32_3, XP64_1, XP64_2, XP64_3, XP64_4, XP32_2, XP64_5, XP32_1};
    const uint64_t nb = (len - 1) / 1024;
    for (uint64_t n = 0; n < nb; n++) {
        for (int s = 0; s < 16; s++) x_acc512(acc, in + n * 1024 + 64 * s, S + 8 * s);
        for (int i = 0; i < 8; i++) acc[i] = (acc[i] ^ (acc[i] >> 47) ^ x_le64(S + 128 + 8 * i)) * XP32_1;
    }
    const uint64_t ns = ((len - 1) - 1024 * nb) / 64;
    for (uint64_t s = 0; s < ns; s++) x_acc512(acc, in + nb * 1024 + 64 * s, S + 8 * s);
    x_acc512(acc, in + len - 64, S + 192 - 64 - 7);
    uint64_t r = len * XP64_1;
    for (int i = 0; i < 4; i++) r += x_fold(acc[2 * i] ^ x_le64(S + 11 + 16 * i), acc[2 * i + 1] ^ x_le64(S + 11 + 16 * i + 8));
    return x_aval3(r);
}
