// pbs_plus_b200/csrc/capi.cu -- C ABI (include/pbsgpu.h) and host orchestration.
//
// Host side of the drop-in boundary: what a Go caller reaches through cgo in place of
// buzhash.NewConfig / backupproxy.NewPBSStore / transfer...WriteEntryReader of the
// reference (internal/pxarmount/commit.go:296-329, :720).  Pure C++ over the CUDA
// runtime; no torch types.  There is no CPU fallback anywhere in this file: every data
// path launches the kernels in scan.cu / resolve.cu / sha256.cu / digestset.cu.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <string>
#include <vector>

#include "internal.cuh"

using namespace pbsgpu;

static const uint32_t DEFAULT_TABLE[256] = {
#include "default_table.inc"
};

// ---------------------------------------------------------------------------
// small caching allocators (device + pinned host): steady-state batches do not
// hit cudaMalloc / cudaHostAlloc.
// ---------------------------------------------------------------------------
struct Block { void *p; size_t size; bool used; };
struct Pool {
    std::vector<Block> blocks;
    bool pinned = false;
    void *get(size_t need) {
        need = (need + 255) & ~(size_t)255;
        if (need == 0) need = 256;
        int best = -1;
        for (size_t i = 0; i < blocks.size(); i++)
            if (!blocks[i].used && blocks[i].size >= need && blocks[i].size <= need * 2 + 4096 &&
                (best < 0 || blocks[i].size < blocks[best].size)) best = (int)i;
        if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
        void *p = nullptr;
        cudaError_t e = pinned ? cudaHostAlloc(&p, need, cudaHostAllocDefault) : cudaMalloc(&p, need);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            trim();   // drop cached free blocks and retry once
            e = pinned ? cudaHostAlloc(&p, need, cudaHostAllocDefault) : cudaMalloc(&p, need);
            if (e != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        }
        blocks.push_back({p, need, true});
        return p;
    }
    void put(void *p) {
        if (!p) return;
        for (auto &b : blocks) if (b.p == p) { b.used = false; return; }
    }
    void trim() {
        std::vector<Block> keep;
        for (auto &b : blocks) {
            if (b.used) keep.push_back(b);
            else if (pinned) cudaFreeHost(b.p); else cudaFree(b.p);
        }
        blocks.swap(keep);
    }
    void destroy() {
        for (auto &b : blocks) { if (pinned) cudaFreeHost(b.p); else cudaFree(b.p); }
        blocks.clear();
    }
};

constexpr int N_STREAMS = 13;   // main + side stream per slot: 28 streams (+ copy stream) <= 32 HW connections

struct pbsgpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaDeviceProp prop;
    std::string err;
    std::recursive_mutex mu;
    cudaStream_t streams[N_STREAMS];
    cudaStream_t streams2[N_STREAMS];   // forked side stream per job stream (latency kernel of the hybrid SHA launch)
    cudaStream_t copy_stream;
    int next_stream = 0;
    Pool dev, pin;
    bool profiling = false;
    int variant = 0;
    // device copies of the chunker table (re-uploaded when the cfg table changes)
    uint32_t *d_table = nullptr, *d_rot = nullptr;
    void *d_crc_tables = nullptr;   // K6 tables, uploaded on first use
    void *d_xxh_tab = nullptr;      // K7 secret words, uploaded on first use
    uint32_t table_cache[256];
    bool table_valid = false;
    pbsgpu_timing last_timing;
    struct pbsgpu_job *pending_back = nullptr;   // async job whose SHA half is not enqueued yet (see flush_pending)
    cudaEvent_t epoch = nullptr;   // recorded at open; kernel intervals are reported relative to it
    // optional spatial partition (CUDA green contexts): `part_sms` SMs are reserved for the latency
    // kernels of long chunks (streams2), everything else runs on the remaining SMs (streams)
    int part_sms = 0, bulk_sms = 0;
    CUgreenCtx g_long = nullptr, g_bulk = nullptr;
    uint64_t stage_bytes = 0;   // host-input staging size (0 = auto)
    bool scan_lanes = false;
    uint64_t xxh3_cap_blocks = 8ull << 20;   // 8 GiB of input, 512 MiB of block sums per pass
};

static int fail(pbsgpu_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        c->err = buf;
    }
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            (void)cudaGetLastError();                                                                  \
            return fail(ctx, PBSGPU_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
        }                                                                                              \
    } while (0)

struct Guard {   // one call at a time per ctx + device binding for this OS thread (goroutines migrate)
    std::lock_guard<std::recursive_mutex> lk;
    explicit Guard(pbsgpu_ctx *c) : lk(c->mu) { cudaSetDevice(c->device); }
};

// ---------------------------------------------------------------------------
extern "C" int pbsgpu_version(void) { return PBSGPU_VERSION; }
extern "C" const uint32_t *pbsgpu_default_table(void) { return DEFAULT_TABLE; }

extern "C" int pbsgpu_config(uint32_t avg, const uint32_t *table, pbsgpu_cfg *out) {
    if (!out || avg < 256u || avg > (1u << 29) || (avg & (avg - 1))) return PBSGPU_EINVAL;
    out->avg = avg; out->min = avg >> 2; out->max = avg << 2;
    out->mask = avg * 2u - 1u; out->break_min = out->mask - 2u; out->window = 64;
    memcpy(out->table, table ? table : DEFAULT_TABLE, sizeof out->table);
    return PBSGPU_OK;
}
extern "C" int pbsgpu_config_kib(uint32_t avg_kib, const uint32_t *table, pbsgpu_cfg *out) {
    if (avg_kib == 0 || avg_kib > (1u << 19)) return PBSGPU_EINVAL;
    return pbsgpu_config(avg_kib << 10, table, out);
}
static bool cfg_ok(const pbsgpu_cfg *c) {
    return c && c->avg >= 256u && c->avg <= (1u << 29) && !(c->avg & (c->avg - 1)) && c->min == c->avg >> 2 &&
           c->max == c->avg << 2 && c->mask == c->avg * 2u - 1u && c->break_min == c->mask - 2u && c->window == 64;
}

// Spatial partition with CUDA green contexts (driver API, resolved at run time so the library does
// not link libcuda): long-chunk latency kernels get `want` SMs of their own, so they are neither
// slowed by co-resident bulk warps nor packed onto a few SMs.  Returns false (and leaves the ctx
// untouched) when the driver does not offer it.
template <typename T> static T driver_ep(const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return (T)p;
}
static bool make_partition(pbsgpu_ctx *ctx, int want) {
    auto devget = driver_ep<CUresult (*)(CUdevice *, int)>("cuDeviceGet");
    auto getRes = driver_ep<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto split = driver_ep<CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
    auto genDesc = driver_ep<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned)>("cuDevResourceGenerateDesc");
    auto gcreate = driver_ep<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
    auto gstream = driver_ep<CUresult (*)(CUstream *, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
    if (!devget || !getRes || !split || !genDesc || !gcreate || !gstream) return false;
    cudaFree(0);   // make sure the primary context exists
    CUdevice dev;
    CUdevResource all, grp[1], rest;
    unsigned n = 1;
    CUdevResourceDesc dA, dB;
    if (devget(&dev, ctx->device) != CUDA_SUCCESS || getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    if (split(grp, &n, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || n < 1 || rest.sm.smCount == 0) return false;
    if (genDesc(&dA, &grp[0], 1) != CUDA_SUCCESS || genDesc(&dB, &rest, 1) != CUDA_SUCCESS) return false;
    if (gcreate(&ctx->g_long, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (gcreate(&ctx->g_bulk, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    for (int i = 0; i < N_STREAMS; i++) {
        CUstream a, b;
        if (gstream(&b, ctx->g_bulk, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
        if (gstream(&a, ctx->g_long, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
        ctx->streams[i] = (cudaStream_t)b;
        ctx->streams2[i] = (cudaStream_t)a;
    }
    ctx->part_sms = (int)grp[0].sm.smCount;
    ctx->bulk_sms = (int)rest.sm.smCount;
    ctx->sm_count = ctx->bulk_sms;   // persistent kernels (scan) size their grid to the bulk partition
    return true;
}

extern "C" int pbsgpu_open(int device, pbsgpu_ctx **out) {
    if (!out) return PBSGPU_EINVAL;
    *out = nullptr;
    // Streams that share a hardware work queue serialise (false dependencies); the default is 8 queues.
    // Only effective if the CUDA context has not been created yet -- hosts that initialise CUDA first
    // (e.g. torch) should export CUDA_DEVICE_MAX_CONNECTIONS=32 themselves (bench.py does).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); return PBSGPU_ENODEV; }
    if (device < 0 || device >= n) return PBSGPU_ENODEV;
    pbsgpu_ctx *ctx = new pbsgpu_ctx();
    ctx->device = device;
    ctx->dev.pinned = false; ctx->pin.pinned = true;
    memset(&ctx->last_timing, 0, sizeof ctx->last_timing);
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        (void)cudaGetLastError(); delete ctx; return PBSGPU_ENODEV;
    }
    ctx->sm_count = ctx->prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char *pe = getenv("PBSGPU_HYBRID_PRIO");
    int side_prio = (pe && atoi(pe)) ? prio_hi : prio_lo;   // 1: long-chunk kernels get the high-priority stream
    // default: 24 SMs reserved for the long-chunk latency kernels (green contexts); 0 disables
    const char *ps = getenv("PBSGPU_PARTITION_SMS");
    int want_part = ps ? atoi(ps) : 24;
    bool partitioned = want_part > 0 && want_part + 8 <= ctx->sm_count && make_partition(ctx, want_part);
    for (int i = 0; i < N_STREAMS && !partitioned; i++)
        if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithPriority(&ctx->streams2[i], cudaStreamNonBlocking, side_prio) != cudaSuccess) { delete ctx; return PBSGPU_ECUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PBSGPU_ECUDA; }
    if (cudaMalloc(&ctx->d_table, 1024) != cudaSuccess || cudaMalloc(&ctx->d_rot, 65536) != cudaSuccess) {
        delete ctx; return PBSGPU_ENOMEM;
    }
    if (cudaEventCreate(&ctx->epoch) != cudaSuccess || cudaEventRecord(ctx->epoch, ctx->streams[0]) != cudaSuccess ||
        cudaEventSynchronize(ctx->epoch) != cudaSuccess) { delete ctx; return PBSGPU_ECUDA; }
    const char *sb = getenv("PBSGPU_STAGE_BYTES");
    if (sb) ctx->stage_bytes = strtoull(sb, nullptr, 0);
    const char *v = getenv("PBSGPU_VARIANT");
    if (v) ctx->variant = atoi(v);
    const char *sl = getenv("PBSGPU_SCAN_LANES");      // 1 = lane-contiguous scan kernel (k_scan_lanes), default off
    if (sl) ctx->scan_lanes = atoi(sl) != 0;
    const char *xc = getenv("PBSGPU_XXH3_CAP_BLOCKS");  // per-pass block budget of K7 (tests force several passes)
    if (xc && atoll(xc) > 0) ctx->xxh3_cap_blocks = (uint64_t)atoll(xc);
    *out = ctx;
    return PBSGPU_OK;
}

extern "C" void pbsgpu_close(pbsgpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < N_STREAMS; i++) { cudaStreamDestroy(ctx->streams[i]); cudaStreamDestroy(ctx->streams2[i]); }
    cudaStreamDestroy(ctx->copy_stream);
    cudaFree(ctx->d_table); cudaFree(ctx->d_rot);
    if (ctx->d_crc_tables) cudaFree(ctx->d_crc_tables);
    if (ctx->d_xxh_tab) cudaFree(ctx->d_xxh_tab);
    if (ctx->epoch) cudaEventDestroy(ctx->epoch);
    if (ctx->g_long || ctx->g_bulk) {
        auto gdestroy = driver_ep<CUresult (*)(CUgreenCtx)>("cuGreenCtxDestroy");
        if (gdestroy) { if (ctx->g_long) gdestroy(ctx->g_long); if (ctx->g_bulk) gdestroy(ctx->g_bulk); }
    }
    ctx->dev.destroy(); ctx->pin.destroy();
    delete ctx;
}

extern "C" const char *pbsgpu_strerror(const pbsgpu_ctx *ctx) {
    if (!ctx) return "pbsgpu: no context (pbsgpu_open failed: no usable CUDA device?)";
    return ctx->err.c_str();
}

extern "C" int pbsgpu_device_info(pbsgpu_ctx *ctx, pbsgpu_devinfo *out) {
    if (!ctx || !out) return PBSGPU_EINVAL;
    Guard g(ctx);
    memset(out, 0, sizeof *out);
    out->device = ctx->device; out->sm_count = ctx->sm_count;
    out->cc_major = ctx->prop.major; out->cc_minor = ctx->prop.minor;
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    out->free_mem = fr; out->total_mem = tot;
    strncpy(out->name, ctx->prop.name, sizeof out->name - 1);
    return PBSGPU_OK;
}
extern "C" int pbsgpu_partition_info(pbsgpu_ctx *ctx, int *long_sms, int *bulk_sms) {
    if (!ctx) return PBSGPU_EINVAL;
    if (long_sms) *long_sms = ctx->part_sms;
    if (bulk_sms) *bulk_sms = ctx->bulk_sms;
    return PBSGPU_OK;
}
extern "C" int pbsgpu_set_profiling(pbsgpu_ctx *ctx, int on) { if (!ctx) return PBSGPU_EINVAL; ctx->profiling = on != 0; return 0; }
extern "C" int pbsgpu_set_kernel_variant(pbsgpu_ctx *ctx, int v) { if (!ctx || v < 0 || v > 1) return PBSGPU_EINVAL; ctx->variant = v; return 0; }

extern "C" void *pbsgpu_host_alloc(pbsgpu_ctx *ctx, uint64_t bytes) {
    if (!ctx) return nullptr;
    Guard g(ctx);
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void pbsgpu_host_free(pbsgpu_ctx *ctx, void *p) { if (ctx && p) { Guard g(ctx); cudaFreeHost(p); } }

static int upload_table(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, cudaStream_t st) {
    if (ctx->table_valid && memcmp(ctx->table_cache, cfg->table, 1024) == 0) return PBSGPU_OK;
    // all streams must be done with the old table before it is replaced
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyAsync(ctx->d_table, cfg->table, 1024, cudaMemcpyHostToDevice, st));
    CK(launch_build_rot_table(ctx->d_table, ctx->d_rot, st));
    CK(cudaStreamSynchronize(st));
    memcpy(ctx->table_cache, cfg->table, 1024);
    ctx->table_valid = true;
    return PBSGPU_OK;
}

// ---------------------------------------------------------------------------
// Job: one batch of device-resident streams through K1..K3 on one CUDA stream.
// ---------------------------------------------------------------------------
enum { EV_START, EV_SCAN, EV_SORT, EV_RESOLVE, EV_SHA, EV_END, EV_FORK, EV_JOIN, EV_BULK, EV_BACK, EV_COUNT };

struct pbsgpu_job {
    pbsgpu_ctx *ctx = nullptr;
    cudaStream_t st = nullptr, st2 = nullptr;
    pbsgpu_cfg cfg;
    const uint8_t *base = nullptr;
    std::vector<uint64_t> off, len, tile_first;
    uint32_t n = 0;
    uint64_t total_bytes = 0, total_tiles = 0, chunk_cap = 0, cand_cap = 0;
    int eof = 1, want_digests = 1, variant = 0;
    bool scan_lanes = false;
    // device
    uint64_t *d_off = nullptr, *d_len = nullptr, *d_tile_first = nullptr, *d_cand = nullptr, *d_cand_sorted = nullptr;
    unsigned long long *d_counters = nullptr;   // [0] cand_count [1] n_chunks
    uint32_t *d_counts = nullptr;
    uint64_t *d_chunk_first = nullptr, *d_consumed = nullptr;
    ChunkRef *d_chunks = nullptr;
    uint32_t *d_keys = nullptr, *d_keys2 = nullptr, *d_vals = nullptr, *d_vals2 = nullptr;
    uint8_t *d_digests = nullptr;
    pbsgpu_chunk *d_out = nullptr;
    void *d_temp = nullptr; size_t temp_bytes = 0;
    // pinned host
    unsigned long long *h_counters = nullptr;
    pbsgpu_chunk *h_out = nullptr;
    uint64_t *h_consumed = nullptr;
    cudaEvent_t ev[EV_COUNT];
    bool have_events = false, profiling = false, enqueued = false, front_done = false, back_done = false;
    uint32_t reruns = 0;
};

static void job_release(pbsgpu_job *j) {
    if (!j) return;
    pbsgpu_ctx *c = j->ctx;
    if (c->pending_back == j) c->pending_back = nullptr;
    void *devp[] = {j->d_off, j->d_len, j->d_tile_first, j->d_cand, j->d_cand_sorted, j->d_counters, j->d_counts,
                    j->d_chunk_first, j->d_consumed, j->d_chunks, j->d_keys, j->d_keys2, j->d_vals, j->d_vals2,
                    j->d_digests, j->d_out, j->d_temp};
    for (void *p : devp) c->dev.put(p);
    c->pin.put(j->h_counters); c->pin.put(j->h_out); c->pin.put(j->h_consumed);
    if (j->have_events) for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(j->ev[i]);
    delete j;
}

static uint64_t expected_cand_cap(const pbsgpu_cfg &cfg, uint64_t total) {
    // candidates occur with probability 3/(mask+1) per byte on random data
    long double e = (long double)total * 3.0L / ((long double)cfg.mask + 1.0L);
    uint64_t cap = (uint64_t)(e * 4.0L) + 4096;
    return cap;
}

static int job_alloc(pbsgpu_job *j) {
    pbsgpu_ctx *ctx = j->ctx;
    Pool &d = ctx->dev;
    const uint32_t n = j->n;
#define DALLOC(ptr, type, count)                                                           \
    do {                                                                                   \
        ptr = (type *)d.get(sizeof(type) * (size_t)(count));                               \
        if (!ptr) return fail(ctx, PBSGPU_ENOMEM, "device allocation of %zu bytes failed", \
                              sizeof(type) * (size_t)(count));                             \
    } while (0)
    DALLOC(j->d_off, uint64_t, n + 1);
    DALLOC(j->d_len, uint64_t, n + 1);
    DALLOC(j->d_tile_first, uint64_t, n + 2);
    DALLOC(j->d_cand, uint64_t, j->cand_cap);
    DALLOC(j->d_cand_sorted, uint64_t, j->cand_cap);
    DALLOC(j->d_counters, unsigned long long, 4);
    DALLOC(j->d_counts, uint32_t, n + 1);
    DALLOC(j->d_chunk_first, uint64_t, n + 2);
    DALLOC(j->d_consumed, uint64_t, n + 1);
    DALLOC(j->d_chunks, ChunkRef, j->chunk_cap + 1);
    DALLOC(j->d_out, pbsgpu_chunk, j->chunk_cap + 1);
    if (j->want_digests) {
        DALLOC(j->d_keys, uint32_t, j->chunk_cap + 1);
        DALLOC(j->d_keys2, uint32_t, j->chunk_cap + 1);
        DALLOC(j->d_vals, uint32_t, j->chunk_cap + 1);
        DALLOC(j->d_vals2, uint32_t, j->chunk_cap + 1);
        DALLOC(j->d_digests, uint8_t, (j->chunk_cap + 1) * 32);
    }
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)j->cand_cap);
    if (j->want_digests)
        cub::DeviceRadixSort::SortPairsDescending(nullptr, t2, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                                  (uint32_t *)nullptr, (uint32_t *)nullptr, (int)j->chunk_cap);
    j->temp_bytes = std::max(t1, t2) + 256;
    DALLOC(j->d_temp, uint8_t, j->temp_bytes);
#undef DALLOC
    j->h_counters = (unsigned long long *)ctx->pin.get(4 * sizeof(unsigned long long));
    j->h_out = (pbsgpu_chunk *)ctx->pin.get(sizeof(pbsgpu_chunk) * (j->chunk_cap + 1));
    j->h_consumed = (uint64_t *)ctx->pin.get(sizeof(uint64_t) * (n + 1));
    if (!j->h_counters || !j->h_out || !j->h_consumed) return fail(ctx, PBSGPU_ENOMEM, "pinned host allocation failed");
    return PBSGPU_OK;
}

static int job_create(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const void *base_dev, const uint64_t *off,
                      const uint64_t *len, uint32_t n, int eof, int want_digests, pbsgpu_job **out) {
    if (!cfg_ok(cfg)) return fail(ctx, PBSGPU_EINVAL, "invalid pbsgpu_cfg (use pbsgpu_config)");
    if (n >= (1u << 24)) return fail(ctx, PBSGPU_EINVAL, "too many streams in one batch (%u >= 2^24)", n);
    pbsgpu_job *j = new pbsgpu_job();
    j->ctx = ctx; j->cfg = *cfg; j->base = (const uint8_t *)base_dev; j->n = n; j->eof = eof;
    j->want_digests = want_digests; j->variant = ctx->variant; j->profiling = ctx->profiling;
    j->off.assign(off, off + n); j->len.assign(len, len + n);
    const uint64_t tile = j->variant == 1 ? (uint64_t)SIMPLE_SPAN : (uint64_t)WARP_TILE;
    j->scan_lanes = j->variant == 0 && ctx->scan_lanes;
    const uint64_t super = scan_lanes_super_bytes(), super_steps = scan_lanes_steps();
    j->tile_first.resize(n + 1);
    uint64_t tiles = 0, total = 0, chunks = 0;
    const uint64_t min_eff = min_effective(cfg->min);
    for (uint32_t i = 0; i < n; i++) {
        if (len[i] >= (1ull << KEY_POS_BITS)) { delete j; return fail(ctx, PBSGPU_EINVAL, "stream %u longer than 2^40 bytes", i); }
        j->tile_first[i] = tiles;
        if (j->scan_lanes) {   // k_scan_lanes: 8 steps per 64 KiB super-tile of a 16 B aligned stream, then plain tiles
            const uint64_t ns = (((uintptr_t)(j->base + off[i])) & 15) == 0 ? len[i] / super : 0;
            tiles += ns * super_steps + (len[i] - ns * super + tile - 1) / tile;
        } else {
            tiles += (len[i] + tile - 1) / tile;
        }
        total += len[i];
        chunks += len[i] / min_eff + 1;
    }
    j->tile_first[n] = tiles;
    j->total_tiles = tiles; j->total_bytes = total; j->chunk_cap = chunks;
    if (chunks >= (1ull << 31)) { delete j; return fail(ctx, PBSGPU_EINVAL, "batch too large (%llu chunk slots)", (unsigned long long)chunks); }
    j->cand_cap = expected_cand_cap(*cfg, total);
    if (j->cand_cap >= (1ull << 31)) { delete j; return fail(ctx, PBSGPU_EINVAL, "batch too large (candidate buffer)"); }
    j->st = ctx->streams[ctx->next_stream];
    j->st2 = ctx->streams2[ctx->next_stream];
    ctx->next_stream = (ctx->next_stream + 1) % N_STREAMS;
    int rc = job_alloc(j);
    if (rc) { job_release(j); return rc; }
    for (int i = 0; i < EV_COUNT; i++)
        if (cudaEventCreateWithFlags(&j->ev[i], j->profiling ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) {
            for (int k = 0; k < i; k++) cudaEventDestroy(j->ev[k]);
            job_release(j);
            return fail(ctx, PBSGPU_ECUDA, "cudaEventCreate failed");
        }
    j->have_events = true;
    *out = j;
    return PBSGPU_OK;
}

// The hybrid SHA launch pays off when the long-chunk kernels have SMs of their own (partition); without
// a partition their CTAs pin 137 KB of shared memory per SM for ~0.3 s and starve the whole-SM scan CTAs of
// later batches (measured: 163 vs 143 ms/step).  PBSGPU_SHA_HYBRID=2 forces it on regardless.
static bool hybrid_for(const pbsgpu_ctx *ctx) {
    static int force = -1;
    if (force < 0) { const char *e = getenv("PBSGPU_SHA_HYBRID"); force = (e && atoi(e) == 2) ? 1 : 0; }
    return ctx->part_sms > 0 || force;
}

// front half: inputs -> K1 scan -> sort -> K2 resolve (chunk list known on the device afterwards)
static int job_enqueue_front(pbsgpu_job *j) {
    pbsgpu_ctx *ctx = j->ctx;
    cudaStream_t st = j->st;
    const uint32_t n = j->n;
    int rc = upload_table(ctx, &j->cfg, st);
    if (rc) return rc;
    if (n) {
        CK(cudaMemcpyAsync(j->d_off, j->off.data(), n * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(j->d_len, j->len.data(), n * 8, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(j->d_tile_first, j->tile_first.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(j->d_counters, 0, 4 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(j->d_cand, 0xFF, j->cand_cap * 8, st));   // KEY_SENTINEL padding for the sort
    if (j->profiling) CK(cudaEventRecord(j->ev[EV_START], st));
    // K1
    ScanArgs sa;
    sa.base = j->base; sa.off = j->d_off; sa.len = j->d_len; sa.tile_first = j->d_tile_first; sa.n_streams = n;
    sa.total_tiles = j->total_tiles; sa.mask = j->cfg.mask; sa.break_min = j->cfg.break_min; sa.table = ctx->d_table;
    sa.cand = j->d_cand; sa.cand_cap = j->cand_cap; sa.cand_count = &j->d_counters[0];
    if (j->variant == 1) CK(launch_scan_simple(sa, st));
    else if (j->scan_lanes) CK(launch_scan_lanes(sa, ctx->d_rot, ctx->sm_count, st));
    else CK(launch_scan_tuned(sa, ctx->d_rot, ctx->sm_count, st));
    CK(cudaEventRecord(j->ev[EV_SCAN], st));   // also the "scan done" signal a predecessor's back half waits for
    // candidates -> sorted by (stream, position)
    size_t tb = j->temp_bytes;
    CK(cub::DeviceRadixSort::SortKeys(j->d_temp, tb, j->d_cand, j->d_cand_sorted, (int)j->cand_cap, 0, 64, st));
    if (j->profiling) CK(cudaEventRecord(j->ev[EV_SORT], st));
    // K2
    ResolveArgs ra;
    ra.keys_sorted = j->d_cand_sorted; ra.cand_count = &j->d_counters[0]; ra.cand_cap = j->cand_cap; ra.len = j->d_len;
    ra.n_streams = n; ra.cmin = j->cfg.min; ra.cmax = j->cfg.max; ra.eof = j->eof; ra.counts = j->d_counts;
    ra.chunk_first = j->d_chunk_first; ra.chunks = j->d_chunks; ra.chunk_cap = j->chunk_cap; ra.consumed = j->d_consumed;
    ra.n_chunks = &j->d_counters[1];
    CK(launch_resolve(ra, st));
    if (j->profiling) CK(cudaEventRecord(j->ev[EV_RESOLVE], st));
    j->front_done = true;
    j->back_done = false;
    return PBSGPU_OK;
}

// back half: K3 SHA-256 (hybrid launch) -> pack -> D2H of the results
static int job_enqueue_back(pbsgpu_job *j) {
    pbsgpu_ctx *ctx = j->ctx;
    cudaStream_t st = j->st;
    const uint32_t n = j->n;
    size_t tb = j->temp_bytes;
    if (j->profiling) CK(cudaEventRecord(j->ev[EV_BACK], st));
    if (j->want_digests && j->chunk_cap) {
        CK(launch_len_keys(j->d_chunks, &j->d_counters[1], j->chunk_cap, j->d_keys, j->d_vals, st));
        tb = j->temp_bytes;
        CK(cub::DeviceRadixSort::SortPairsDescending(j->d_temp, tb, j->d_keys, j->d_keys2, j->d_vals, j->d_vals2,
                                                     (int)j->chunk_cap, 0, 32, st));
        ShaArgs ha;
        ha.base = j->base; ha.off = j->d_off; ha.chunks = j->d_chunks; ha.order = j->d_vals2;
        ha.n_chunks = &j->d_counters[1]; ha.chunk_cap = j->chunk_cap; ha.digests = j->d_digests;
        ha.n_head = nullptr; ha.part = 0;
        if (j->variant == 1) CK(launch_sha_simple(ha, st));
        else if (!sha_hybrid_enabled() || !hybrid_for(ctx)) CK(launch_sha_tuned(ha, ctx->sm_count, st));
        else {
            // hybrid: chunks longer than 2.5 x avg (their serial chains bound the batch's makespan) run on
            // the latency-optimised split kernel on a forked stream, concurrently with the rest
            static int thr_x10 = -1, serial = -1;
            if (thr_x10 < 0) { const char *e = getenv("PBSGPU_HYBRID_THR_X10"); thr_x10 = e ? atoi(e) : 25;
                               const char *s2 = getenv("PBSGPU_HYBRID_SERIAL"); serial = s2 ? atoi(s2) : 0; }
            cudaStream_t side = serial ? st : j->st2;
            uint64_t thr64 = (uint64_t)j->cfg.avg * (uint64_t)thr_x10 / 10;
            uint32_t thr = thr64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)thr64;
            // at most one latency CTA (32 chunks) per SM of the partition and job; the longest chunks first
            const unsigned long long max_head = 32ull * (unsigned long long)(ctx->part_sms > 0 ? ctx->part_sms : 24);
            CK(launch_split_point(j->d_keys2, &j->d_counters[1], j->chunk_cap, thr, max_head, &j->d_counters[2], st));
            CK(cudaEventRecord(j->ev[EV_FORK], st));
            if (!serial) CK(cudaStreamWaitEvent(side, j->ev[EV_FORK], 0));
            ha.n_head = &j->d_counters[2];
            ha.part = 1; CK(launch_sha_split(ha, side));
            CK(cudaEventRecord(j->ev[EV_JOIN], side));
            ha.part = 2; CK(launch_sha_tuned(ha, ctx->sm_count, st));
            CK(cudaEventRecord(j->ev[EV_BULK], st));
            if (!serial) CK(cudaStreamWaitEvent(st, j->ev[EV_JOIN], 0));
        }
    }
    if (j->profiling) CK(cudaEventRecord(j->ev[EV_SHA], st));
    if (j->want_digests)
        CK(launch_pack_chunks(j->d_chunks, j->d_digests, nullptr, &j->d_counters[1], j->chunk_cap, j->d_out, st));
    CK(cudaMemcpyAsync(j->h_counters, j->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (j->want_digests)
        CK(cudaMemcpyAsync(j->h_out, j->d_out, sizeof(pbsgpu_chunk) * j->chunk_cap, cudaMemcpyDeviceToHost, st));
    if (!j->eof && n) CK(cudaMemcpyAsync(j->h_consumed, j->d_consumed, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(j->ev[EV_END], st));
    j->enqueued = true;
    j->back_done = true;
    return PBSGPU_OK;
}

// Scans are whole-SM CTAs (207 KB shared memory, 49k registers) and cannot be placed on an SM that
// SHA CTAs of earlier batches already fill, so a scan submitted behind running SHA work stalls the
// pipeline.  The asynchronous API therefore keeps the back half (SHA) of the most recent job
// pending until the NEXT job's scan has been enqueued (and makes it wait for that scan), so scans
// always run ahead of the SHA work that would block them.
static int flush_pending(pbsgpu_ctx *ctx, pbsgpu_job *successor) {
    pbsgpu_job *p = ctx->pending_back;
    if (!p) return PBSGPU_OK;
    ctx->pending_back = nullptr;
    if (successor) CK(cudaStreamWaitEvent(p->st, successor->ev[EV_SCAN], 0));
    return job_enqueue_back(p);
}

static int job_enqueue(pbsgpu_job *j) {
    int rc = job_enqueue_front(j);
    if (rc) return rc;
    return job_enqueue_back(j);
}

// dense candidates (adversarial / highly structured data): replace the candidate buffers by exactly sized ones
static int job_grow_cands(pbsgpu_job *j, unsigned long long nc) {
    pbsgpu_ctx *ctx = j->ctx;
    ctx->dev.put(j->d_cand); ctx->dev.put(j->d_cand_sorted); ctx->dev.put(j->d_temp);
    j->d_cand = j->d_cand_sorted = nullptr; j->d_temp = nullptr;
    j->cand_cap = nc + nc / 8 + 4096;
    if (j->cand_cap >= (1ull << 31)) return fail(ctx, PBSGPU_ENOMEM, "candidate density too high (%llu candidates)", nc);
    j->d_cand = (uint64_t *)ctx->dev.get(j->cand_cap * 8);
    j->d_cand_sorted = (uint64_t *)ctx->dev.get(j->cand_cap * 8);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)j->cand_cap);
    if (j->want_digests)
        cub::DeviceRadixSort::SortPairsDescending(nullptr, t2, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                                  (uint32_t *)nullptr, (uint32_t *)nullptr, (int)j->chunk_cap);
    j->temp_bytes = std::max(t1, t2) + 256;
    j->d_temp = ctx->dev.get(j->temp_bytes);
    if (!j->d_cand || !j->d_cand_sorted || !j->d_temp) return fail(ctx, PBSGPU_ENOMEM, "device allocation failed (rerun)");
    return PBSGPU_OK;
}

// Blocks until the job is done; reruns it with a larger candidate buffer if the
// (statistically sized) one overflowed -- results are exact either way.
static int job_finish(pbsgpu_job *j) {
    pbsgpu_ctx *ctx = j->ctx;
    for (;;) {
        CK(cudaEventSynchronize(j->ev[EV_END]));
        unsigned long long nc = j->h_counters[0];
        if (nc <= j->cand_cap) break;
        int grc = job_grow_cands(j, nc);
        if (grc) return grc;
        j->reruns++;
        int rc = job_enqueue(j);
        if (rc) return rc;
    }
    pbsgpu_timing &t = ctx->last_timing;
    memset(&t, 0, sizeof t);
    t.bytes = j->total_bytes; t.chunks = j->h_counters[1]; t.candidates = j->h_counters[0]; t.reruns = j->reruns;
    t.scan_launches = 1;
    const bool hyb = j->variant == 0 && sha_hybrid_enabled() && hybrid_for(ctx);
    t.sha_launches = j->want_digests ? (hyb ? 2 : 1) : 0;
    t.other_launches = 3 + (j->want_digests ? 2 + (hyb ? 1 : 0) : 0);
    if (j->profiling) {
        cudaEventElapsedTime(&t.scan_ms, j->ev[EV_START], j->ev[EV_SCAN]);
        cudaEventElapsedTime(&t.sort_ms, j->ev[EV_SCAN], j->ev[EV_SORT]);
        cudaEventElapsedTime(&t.resolve_ms, j->ev[EV_SORT], j->ev[EV_RESOLVE]);
        cudaEventElapsedTime(&t.sha_ms, j->ev[EV_BACK], j->ev[EV_SHA]);
        cudaEventElapsedTime(&t.total_ms, j->ev[EV_START], j->ev[EV_END]);
        cudaEventElapsedTime(&t.scan_t0, ctx->epoch, j->ev[EV_START]);
        cudaEventElapsedTime(&t.scan_t1, ctx->epoch, j->ev[EV_SCAN]);
        cudaEventElapsedTime(&t.sha_t0, ctx->epoch, j->ev[EV_BACK]);
        cudaEventElapsedTime(&t.sha_t1, ctx->epoch, j->ev[EV_SHA]);
        if (j->want_digests && hyb && j->chunk_cap) {
            cudaEventElapsedTime(&t.sha_long_ms, j->ev[EV_FORK], j->ev[EV_JOIN]);
            cudaEventElapsedTime(&t.sha_bulk_ms, j->ev[EV_FORK], j->ev[EV_BULK]);
            (void)cudaGetLastError();
        }
    }
    return PBSGPU_OK;
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

extern "C" int pbsgpu_batch_submit(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const void *base_dev, const uint64_t *off,
                                   const uint64_t *len, uint32_t n, pbsgpu_job **job) {
    if (!ctx || !job || (n && (!off || !len))) return PBSGPU_EINVAL;
    Guard g(ctx);
    *job = nullptr;
    if (n && !is_device_ptr(base_dev)) return fail(ctx, PBSGPU_EINVAL, "pbsgpu_batch_submit needs a device pointer");
    pbsgpu_job *j = nullptr;
    int rc = job_create(ctx, cfg, base_dev, off, len, n, 1, 1, &j);
    if (rc) return rc;
    rc = job_enqueue_front(j);
    if (rc == PBSGPU_OK) rc = flush_pending(ctx, j);     // predecessor's SHA goes in behind this job's scan
    if (rc) { cudaStreamSynchronize(j->st); job_release(j); return rc; }
    static int defer = -1;
    if (defer < 0) { const char *e = getenv("PBSGPU_DEFER_SHA"); defer = e ? atoi(e) : 0; }   // measured: no gain
    if (defer) ctx->pending_back = j;
    else { rc = job_enqueue_back(j); if (rc) { cudaStreamSynchronize(j->st); job_release(j); return rc; } }
    *job = j;
    return PBSGPU_OK;
}

extern "C" int pbsgpu_batch_wait(pbsgpu_job *j, pbsgpu_chunk *out, uint64_t cap, uint64_t *n_out, pbsgpu_timing *timing) {
    if (!j) return PBSGPU_EINVAL;
    pbsgpu_ctx *ctx = j->ctx;
    Guard g(ctx);
    int rc = PBSGPU_OK;
    if (ctx->pending_back == j) rc = flush_pending(ctx, nullptr);
    if (rc == PBSGPU_OK) rc = job_finish(j);
    if (rc == PBSGPU_OK) {
        uint64_t nch = j->h_counters[1];
        if (n_out) *n_out = nch;
        if (timing) *timing = ctx->last_timing;
        if (nch > cap) rc = fail(ctx, PBSGPU_ERANGE, "output capacity %llu < %llu chunks", (unsigned long long)cap, (unsigned long long)nch);
        else if (nch) memcpy(out, j->h_out, nch * sizeof(pbsgpu_chunk));
    } else {
        cudaStreamSynchronize(j->st);
    }
    job_release(j);
    return rc;
}

// ---------------------------------------------------------------------------
// digest set
// ---------------------------------------------------------------------------
struct pbsgpu_set {
    pbsgpu_ctx *ctx;
    SetTable t;
    uint64_t count;
};

static int set_alloc_table(pbsgpu_ctx *ctx, uint64_t cap, SetTable *t) {
    t->cap = cap;
    t->tags = (uint64_t *)ctx->dev.get(cap * 8);
    t->keys = (uint64_t *)ctx->dev.get(cap * 32);
    if (!t->tags || !t->keys) return fail(ctx, PBSGPU_ENOMEM, "digest set allocation failed (%llu slots)", (unsigned long long)cap);
    CK(cudaMemsetAsync(t->tags, 0, cap * 8, ctx->streams[0]));
    return PBSGPU_OK;
}

extern "C" int pbsgpu_set_create(pbsgpu_ctx *ctx, uint64_t capacity_hint, pbsgpu_set **out) {
    if (!ctx || !out) return PBSGPU_EINVAL;
    Guard g(ctx);
    uint64_t cap = 1024;
    while (cap < capacity_hint * 2) cap <<= 1;
    pbsgpu_set *s = new pbsgpu_set();
    s->ctx = ctx; s->count = 0;
    int rc = set_alloc_table(ctx, cap, &s->t);
    if (rc) { delete s; return rc; }
    CK(cudaStreamSynchronize(ctx->streams[0]));
    *out = s;
    return PBSGPU_OK;
}
extern "C" void pbsgpu_set_destroy(pbsgpu_set *s) {
    if (!s) return;
    Guard g(s->ctx);
    cudaStreamSynchronize(s->ctx->streams[0]);
    s->ctx->dev.put(s->t.tags); s->ctx->dev.put(s->t.keys);
    delete s;
}
extern "C" int pbsgpu_set_count(pbsgpu_set *s, uint64_t *count) { if (!s || !count) return PBSGPU_EINVAL; *count = s->count; return 0; }

// d_dig: device pointer to n*32 bytes.  d_hit: device n bytes or NULL.
static int set_process_dev(pbsgpu_set *s, const uint8_t *d_dig, uint64_t n, int do_insert, uint8_t *d_hit) {
    pbsgpu_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->streams[0];
    if (n == 0) return PBSGPU_OK;
    if (n >= (1ull << 31)) return fail(ctx, PBSGPU_EINVAL, "too many digests in one call");
    if (do_insert && (s->count + n) * 2 > s->t.cap) {   // keep load <= 50 %
        uint64_t cap = s->t.cap;
        while ((s->count + n) * 2 > cap) cap <<= 1;
        SetTable nt;
        int rc = set_alloc_table(ctx, cap, &nt);
        if (rc) return rc;
        CK(launch_set_rehash(s->t, nt, st));
        CK(cudaStreamSynchronize(st));
        ctx->dev.put(s->t.tags); ctx->dev.put(s->t.keys);
        s->t = nt;
    }
    uint64_t *tag = (uint64_t *)ctx->dev.get(n * 8), *tag2 = (uint64_t *)ctx->dev.get(n * 8);
    uint32_t *idx = (uint32_t *)ctx->dev.get(n * 4), *idx2 = (uint32_t *)ctx->dev.get(n * 4);
    uint8_t *miss = (uint8_t *)ctx->dev.get(n);
    unsigned long long *d_new = (unsigned long long *)ctx->dev.get(8);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)n);
    void *temp = ctx->dev.get(tb + 256);
    int rc = PBSGPU_OK;
    unsigned long long h_new = 0;
    auto cleanup = [&]() {
        ctx->dev.put(tag); ctx->dev.put(tag2); ctx->dev.put(idx); ctx->dev.put(idx2); ctx->dev.put(miss);
        ctx->dev.put(d_new); ctx->dev.put(temp);
    };
    if (!tag || !tag2 || !idx || !idx2 || !miss || !d_new || !temp) { cleanup(); return fail(ctx, PBSGPU_ENOMEM, "digest set scratch allocation failed"); }
#define CKS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { (void)cudaGetLastError(); cleanup(); \
        return fail(ctx, PBSGPU_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); } } while (0)
    CKS(cudaMemsetAsync(d_new, 0, 8, st));
    CKS(launch_set_make_keys(d_dig, n, tag, idx, st));
    CKS(cub::DeviceRadixSort::SortPairs(temp, tb, tag, tag2, idx, idx2, (int)n, 0, 64, st));   // stable
    CKS(launch_set_mark_probe_insert(s->t, d_dig, tag2, idx2, n, do_insert, d_hit, miss, d_new, st));
    CKS(cudaMemcpyAsync(&h_new, d_new, 8, cudaMemcpyDeviceToHost, st));
    CKS(cudaStreamSynchronize(st));
#undef CKS
    s->count += h_new;
    cleanup();
    return rc;
}

static int set_process(pbsgpu_set *s, const uint8_t *d32, uint64_t n, int do_insert, uint8_t *hit_host) {
    if (!s || (n && !d32)) return PBSGPU_EINVAL;
    pbsgpu_ctx *ctx = s->ctx;
    Guard g(ctx);
    if (n == 0) return PBSGPU_OK;
    cudaStream_t st = ctx->streams[0];
    const uint8_t *d_dig = d32;
    uint8_t *staged = nullptr;
    if (!is_device_ptr(d32)) {
        staged = (uint8_t *)ctx->dev.get(n * 32);
        if (!staged) return fail(ctx, PBSGPU_ENOMEM, "digest staging allocation failed");
        CK(cudaMemcpyAsync(staged, d32, n * 32, cudaMemcpyHostToDevice, st));
        d_dig = staged;
    }
    uint8_t *d_hit = hit_host ? (uint8_t *)ctx->dev.get(n) : nullptr;
    int rc = set_process_dev(s, d_dig, n, do_insert, d_hit);
    if (rc == PBSGPU_OK && hit_host) {
        cudaError_t e = cudaMemcpy(hit_host, d_hit, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(ctx, PBSGPU_ECUDA, "hit copy: %s", cudaGetErrorString(e));
    }
    ctx->dev.put(staged); ctx->dev.put(d_hit);
    return rc;
}
extern "C" int pbsgpu_set_insert(pbsgpu_set *s, const uint8_t *d32, uint64_t n, uint8_t *hit) { return set_process(s, d32, n, 1, hit); }
extern "C" int pbsgpu_set_probe(pbsgpu_set *s, const uint8_t *d32, uint64_t n, uint8_t *hit) { return set_process(s, d32, n, 0, hit); }

extern "C" int pbsgpu_set_seed_didx(pbsgpu_set *s, const uint8_t *didx, uint64_t size, uint64_t *n_entries) {
    if (!s || !didx) return PBSGPU_EINVAL;
    pbsgpu_ctx *ctx = s->ctx;
    if (size < 4096 || (size - 4096) % 40) return fail(ctx, PBSGPU_EINVAL, "not a dynamic index image (size %llu)", (unsigned long long)size);
    uint64_t n = (size - 4096) / 40;
    std::vector<uint8_t> dig(n * 32);
    for (uint64_t i = 0; i < n; i++) memcpy(&dig[i * 32], didx + 4096 + i * 40 + 8, 32);
    if (n_entries) *n_entries = n;
    return set_process(s, dig.data(), n, 1, nullptr);
}

// ---------------------------------------------------------------------------
// f1: dynamic index images
// ---------------------------------------------------------------------------
static const uint8_t DIDX_MAGIC[8] = {28, 145, 78, 165, 25, 186, 179, 205};
extern "C" int pbsgpu_sha256_batch(pbsgpu_ctx *, const void *, const uint64_t *, const uint64_t *, uint32_t, uint8_t *);

extern "C" uint64_t pbsgpu_didx_size(uint64_t n) { return 4096 + n * 40; }

extern "C" int pbsgpu_didx_build(pbsgpu_ctx *ctx, const pbsgpu_chunk *chunks, uint64_t n, const uint8_t uuid[16],
                                 int64_t ctime, uint8_t *out, uint64_t cap) {
    if (!ctx || (n && !chunks) || !out || !uuid) return PBSGPU_EINVAL;
    if (cap < pbsgpu_didx_size(n)) return fail(ctx, PBSGPU_ERANGE, "didx buffer too small: %llu < %llu", (unsigned long long)cap, (unsigned long long)pbsgpu_didx_size(n));
    memset(out, 0, 4096);
    memcpy(out, DIDX_MAGIC, 8);
    memcpy(out + 8, uuid, 16);
    for (int i = 0; i < 8; i++) out[24 + i] = (uint8_t)((uint64_t)ctime >> (8 * i));
    uint64_t total = 0, prev_end = 0;
    uint32_t prev_stream = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (i == 0 || chunks[i].stream != prev_stream) prev_end = 0;
        if (chunks[i].end_off < prev_end) return fail(ctx, PBSGPU_EINVAL, "chunk records not ordered by (stream, end_off) at %llu", (unsigned long long)i);
        total += chunks[i].end_off - prev_end;
        prev_end = chunks[i].end_off; prev_stream = chunks[i].stream;
        uint8_t *e = out + 4096 + i * 40;
        for (int k = 0; k < 8; k++) e[k] = (uint8_t)(total >> (8 * k));
        memcpy(e + 8, chunks[i].digest, 32);
    }
    uint64_t off0 = 0, len0 = n * 40;
    return pbsgpu_sha256_batch(ctx, out + 4096, &off0, &len0, 1, out + 32);   // index_csum (GPU)
}

extern "C" int pbsgpu_didx_parse(pbsgpu_ctx *ctx, const uint8_t *didx, uint64_t size, uint64_t *ends, uint8_t *digests,
                                 uint64_t cap, uint64_t *n_entries, int verify) {
    if (!didx || !n_entries) return PBSGPU_EINVAL;
    if (size < 4096 || (size - 4096) % 40 || memcmp(didx, DIDX_MAGIC, 8) != 0)
        return fail(ctx, PBSGPU_EINVAL, "not a dynamic index image");
    uint64_t n = (size - 4096) / 40;
    *n_entries = n;
    if (ends || digests) {
        if (cap < n) return fail(ctx, PBSGPU_ERANGE, "output capacity %llu < %llu entries", (unsigned long long)cap, (unsigned long long)n);
        uint64_t prev = 0;
        for (uint64_t i = 0; i < n; i++) {
            const uint8_t *e = didx + 4096 + i * 40;
            uint64_t end = 0;
            for (int k = 0; k < 8; k++) end |= (uint64_t)e[k] << (8 * k);
            if (end < prev) return fail(ctx, PBSGPU_EINVAL, "index offsets not monotonic at entry %llu", (unsigned long long)i);
            prev = end;
            if (ends) ends[i] = end;
            if (digests) memcpy(digests + i * 32, e + 8, 32);
        }
    }
    if (verify) {
        if (!ctx) return PBSGPU_EINVAL;
        uint8_t csum[32];
        uint64_t off0 = 0, len0 = n * 40;
        int rc = pbsgpu_sha256_batch(ctx, didx + 4096, &off0, &len0, 1, csum);
        if (rc) return rc;
        if (memcmp(csum, didx + 32, 32) != 0) return fail(ctx, PBSGPU_EINVAL, "index checksum mismatch");
    }
    return PBSGPU_OK;
}

// ---------------------------------------------------------------------------
// f3: DataBlob checksums
// ---------------------------------------------------------------------------
static const uint8_t BLOB_MAGIC_UNCOMPRESSED[8] = {66, 171, 56, 7, 190, 131, 112, 161};

extern "C" void pbsgpu_blob_header(uint32_t crc, uint8_t out[12]) {
    memcpy(out, BLOB_MAGIC_UNCOMPRESSED, 8);
    for (int i = 0; i < 4; i++) out[8 + i] = (uint8_t)(crc >> (8 * i));
}

extern "C" int pbsgpu_crc32_batch(pbsgpu_ctx *ctx, const void *base, const uint64_t *off, const uint64_t *len,
                                  uint32_t n, uint32_t *crc_out) {
    if (!ctx || (n && (!off || !len || !crc_out))) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (n == 0) return PBSGPU_OK;
    cudaStream_t st = ctx->streams[0];
    if (!ctx->d_crc_tables) {
        std::vector<uint8_t> h(crc_tables_bytes());
        crc_fill_tables_host(h.data());
        CK(cudaMalloc(&ctx->d_crc_tables, h.size()));
        CK(cudaMemcpy(ctx->d_crc_tables, h.data(), h.size(), cudaMemcpyHostToDevice));
    }
    static int crc_variant = -1;   // 0 = TMA-tiled kernel (default), 1 = simple lane-strided kernel
    if (crc_variant < 0) { const char *e = getenv("PBSGPU_CRC_VARIANT"); crc_variant = e ? atoi(e) : 0; }
    uint64_t hi = 0;
    for (uint32_t i = 0; i < n; i++) hi = std::max(hi, off[i] + len[i]);
    const uint8_t *dbase = (const uint8_t *)base;
    uint8_t *staged = nullptr;
    if (hi && !is_device_ptr(base)) {
        staged = (uint8_t *)ctx->dev.get(hi + 16);
        if (!staged) return fail(ctx, PBSGPU_ENOMEM, "staging of %llu bytes failed", (unsigned long long)hi);
        CK(cudaMemcpyAsync(staged, base, hi, cudaMemcpyHostToDevice, st));
        dbase = staged;
    }
    // work units: 128 KiB warp blocks (simple) or <= 576 KiB regions of the 16 B aligned body (tiled)
    const uint64_t UNIT = crc_variant ? crc_wb_bytes() : crc_region_bytes();
    std::vector<uint64_t> wb_first(n + 1);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        wb_first[i] = total;
        if (crc_variant) total += (len[i] + UNIT - 1) / UNIT;
        else if (len[i]) {
            uint64_t head = (16 - ((uintptr_t)(dbase + off[i]) & 15)) & 15;
            uint64_t body = len[i] > head ? len[i] - head : 0;
            total += std::max<uint64_t>(1, (body + UNIT - 1) / UNIT);
        }
    }
    wb_first[n] = total;
    uint64_t *d_off = (uint64_t *)ctx->dev.get(n * 8), *d_len = (uint64_t *)ctx->dev.get(n * 8);
    uint64_t *d_first = (uint64_t *)ctx->dev.get((n + 1) * 8);
    uint32_t *d_part = (uint32_t *)ctx->dev.get((total + 1) * 4), *d_out = (uint32_t *)ctx->dev.get((uint64_t)n * 4);
    int rc = PBSGPU_OK;
    if (!d_off || !d_len || !d_first || !d_part || !d_out) rc = fail(ctx, PBSGPU_ENOMEM, "device allocation failed");
    else {
        cudaError_t e = cudaMemcpyAsync(d_off, off, n * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_len, len, n * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_first, wb_first.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = crc_variant ? launch_crc32(dbase, d_off, d_len, d_first, n, total, ctx->d_crc_tables, d_part, d_out, ctx->sm_count, st)
                            : launch_crc32_tiled(dbase, d_off, d_len, d_first, n, total, ctx->d_crc_tables, d_part, d_out, ctx->sm_count, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(crc_out, d_out, (uint64_t)n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { (void)cudaGetLastError(); rc = fail(ctx, PBSGPU_ECUDA, "crc32 batch: %s", cudaGetErrorString(e)); }
    }
    ctx->dev.put(d_off); ctx->dev.put(d_len); ctx->dev.put(d_first); ctx->dev.put(d_part); ctx->dev.put(d_out);
    ctx->dev.put(staged);
    return rc;
}

// ---------------------------------------------------------------------------
// f2: XXH3-64 of n byte ranges (the commit walk's per-file content hash, commit.go:717-725, :957-976).
// Blocks are hashed in passes of at most PBSGPU_XXH3_CAP_BLOCKS (default 8 Mi = 8 GiB of input, 512 MiB
// of per-block sums); a pass takes the same block window of EVERY stream so the chains stay parallel.
// xxh3_enqueue only enqueues (all passes, no host wait) so the fused batch call can put it on a job's
// stream; xxh3_collect waits for the stream and brings the n hashes back.
// ---------------------------------------------------------------------------
struct XxhRun {
    uint64_t *d_off = nullptr, *d_len = nullptr, *d_first = nullptr, *d_out = nullptr, *d_state = nullptr, *d_S = nullptr;
    uint32_t n = 0;
};

static void xxh3_release(pbsgpu_ctx *ctx, XxhRun *r) {
    ctx->dev.put(r->d_off); ctx->dev.put(r->d_len); ctx->dev.put(r->d_first); ctx->dev.put(r->d_out);
    ctx->dev.put(r->d_state); ctx->dev.put(r->d_S);
    *r = XxhRun();
}

static int xxh3_enqueue(pbsgpu_ctx *ctx, const uint8_t *dbase, const uint64_t *off, const uint64_t *len, uint32_t n,
                        cudaStream_t st, XxhRun *r) {
    if (!ctx->d_xxh_tab) {
        std::vector<uint8_t> h(xxh3_tables_bytes());
        xxh3_fill_tables_host(h.data());
        CK(cudaMalloc(&ctx->d_xxh_tab, h.size()));
        CK(cudaMemcpy(ctx->d_xxh_tab, h.data(), h.size(), cudaMemcpyHostToDevice));
    }
    const uint64_t cap_blocks = ctx->xxh3_cap_blocks;
    uint64_t total_all = 0, max_nb = 0, n_long = 0;
    std::vector<uint64_t> nb(n);
    for (uint32_t i = 0; i < n; i++) {
        nb[i] = len[i] > 240 ? (len[i] - 1) >> 10 : 0;
        total_all += nb[i]; max_nb = std::max(max_nb, nb[i]); n_long += nb[i] != 0;
    }
    const bool one = total_all <= cap_blocks;
    const uint64_t win = one ? std::max<uint64_t>(max_nb, 1) : std::max<uint64_t>(1, cap_blocks / n_long);
    const uint64_t s_blocks = one ? total_all : std::min(total_all, n_long * win);
    const uint64_t passes = (max_nb + win - 1) / win;
    r->n = n;
    r->d_off = (uint64_t *)ctx->dev.get(n * 8); r->d_len = (uint64_t *)ctx->dev.get(n * 8);
    r->d_first = (uint64_t *)ctx->dev.get(std::max<uint64_t>(passes, 1) * (n + 1) * 8);
    r->d_out = (uint64_t *)ctx->dev.get((uint64_t)n * 8); r->d_state = (uint64_t *)ctx->dev.get((uint64_t)n * 64);
    r->d_S = (uint64_t *)ctx->dev.get(std::max<uint64_t>(s_blocks, 1) * 64);
    if (!r->d_off || !r->d_len || !r->d_first || !r->d_out || !r->d_state || !r->d_S) {
        xxh3_release(ctx, r);
        return fail(ctx, PBSGPU_ENOMEM, "xxh3: device allocation failed");
    }
    cudaError_t e = cudaMemcpyAsync(r->d_off, off, n * 8, cudaMemcpyHostToDevice, st);   // pageable sources are staged before return
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_len, len, n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_xxh3_small(dbase, r->d_off, r->d_len, n, ctx->d_xxh_tab, r->d_out, st);
    std::vector<uint64_t> first(n + 1);
    uint64_t pass = 0;
    for (uint64_t lo = 0; e == cudaSuccess && lo < max_nb; lo += win, pass++) {
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; i++) {
            first[i] = total;
            if (nb[i] > lo) total += std::min(nb[i] - lo, win);
        }
        first[n] = total;
        uint64_t *df = r->d_first + pass * (n + 1);
        e = cudaMemcpyAsync(df, first.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = launch_xxh3_pass(dbase, r->d_off, r->d_len, df, n, total, lo, win, ctx->d_xxh_tab, r->d_S, r->d_state, r->d_out, st);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        cudaStreamSynchronize(st);
        xxh3_release(ctx, r);
        return fail(ctx, PBSGPU_ECUDA, "xxh3: %s", cudaGetErrorString(e));
    }
    return PBSGPU_OK;
}

static int xxh3_collect(pbsgpu_ctx *ctx, XxhRun *r, uint64_t *hash_out, cudaStream_t st) {
    if (!r->d_out) return PBSGPU_OK;
    cudaError_t e = cudaMemcpyAsync(hash_out, r->d_out, (uint64_t)r->n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    xxh3_release(ctx, r);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, PBSGPU_ECUDA, "xxh3: %s", cudaGetErrorString(e)); }
    return PBSGPU_OK;
}

extern "C" int pbsgpu_xxh3_batch(pbsgpu_ctx *ctx, const void *base, const uint64_t *off, const uint64_t *len,
                                 uint32_t n, uint64_t *hash_out) {
    if (!ctx || (n && (!off || !len || !hash_out))) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (n == 0) return PBSGPU_OK;
    cudaStream_t st = ctx->streams[0];
    uint64_t hi = 0;
    for (uint32_t i = 0; i < n; i++) hi = std::max(hi, off[i] + len[i]);
    const uint8_t *dbase = (const uint8_t *)base;
    uint8_t *staged = nullptr;
    if (hi && !is_device_ptr(base)) {
        staged = (uint8_t *)ctx->dev.get(hi + 16);
        if (!staged) return fail(ctx, PBSGPU_ENOMEM, "staging of %llu bytes failed", (unsigned long long)hi);
        CK(cudaMemcpyAsync(staged, base, hi, cudaMemcpyHostToDevice, st));
        dbase = staged;
    }
    XxhRun run;
    int rc = xxh3_enqueue(ctx, dbase, off, len, n, st, &run);
    if (rc == PBSGPU_OK) rc = xxh3_collect(ctx, &run, hash_out, st);
    else cudaStreamSynchronize(st);
    ctx->dev.put(staged);
    return rc;
}

// flags for chunk records that are already on the host, in order
static int apply_set(pbsgpu_set *set, pbsgpu_chunk *out, uint64_t n) {
    if (!set || n == 0) return PBSGPU_OK;
    std::vector<uint8_t> dig(n * 32), hit(n);
    for (uint64_t i = 0; i < n; i++) memcpy(&dig[i * 32], out[i].digest, 32);
    int rc = set_process(set, dig.data(), n, 1, hit.data());
    if (rc) return rc;
    for (uint64_t i = 0; i < n; i++) if (hit[i]) out[i].flags |= PBSGPU_CHUNK_KNOWN;
    return PBSGPU_OK;
}

// ---------------------------------------------------------------------------
// Synchronous batch; host input is staged through device buffers, group by group,
// H2D of group k+1 overlapping the kernels of group k.
// ---------------------------------------------------------------------------
struct Group { uint32_t first, count; uint64_t bytes; };

static int batch_host(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const uint8_t *base, const uint64_t *off,
                      const uint64_t *len, uint32_t n, pbsgpu_chunk *out, uint64_t cap, uint64_t *n_out,
                      uint64_t *xxh3_out) {
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    uint64_t stage = ctx->stage_bytes ? ctx->stage_bytes : std::min<uint64_t>(4ull << 30, fr / 16);
    auto al = [](uint64_t x) { return (x + 255) & ~255ull; };
    std::vector<Group> groups;
    {
        Group g{0, 0, 0};
        for (uint32_t i = 0; i < n; i++) {
            uint64_t b = al(len[i]);
            if (g.count && g.bytes + b > stage) { groups.push_back(g); g = Group{i, 0, 0}; }
            g.count++; g.bytes += b;
        }
        if (g.count) groups.push_back(g);
    }
    constexpr int NBUF = 6;   // staged groups in flight: their SHA tails overlap the next groups' copies
    uint64_t buf_bytes = 256;
    for (auto &g : groups) buf_bytes = std::max(buf_bytes, g.bytes);
    uint8_t *bufs[NBUF] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int nbuf = (int)std::min<size_t>(NBUF, groups.size());
    for (int b = 0; b < nbuf; b++) {
        bufs[b] = (uint8_t *)ctx->dev.get(buf_bytes);
        if (!bufs[b]) { for (int k = 0; k < b; k++) ctx->dev.put(bufs[k]); return fail(ctx, PBSGPU_ENOMEM, "staging buffer of %llu bytes failed", (unsigned long long)buf_bytes); }
    }
    std::vector<pbsgpu_job *> jobs(groups.size(), nullptr);
    std::vector<XxhRun> xruns(xxh3_out ? groups.size() : 0);   // f2: per-file XXH3-64 from the same staged bytes
    std::vector<cudaEvent_t> copied(groups.size());
    for (auto &e : copied) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    int rc = PBSGPU_OK;
    uint64_t produced = 0;
    bool overflow = false;
    auto collect = [&](size_t gi) -> int {
        pbsgpu_job *j = jobs[gi];
        int r = job_finish(j);
        if (xxh3_out) { int rx = xxh3_collect(ctx, &xruns[gi], xxh3_out + groups[gi].first, j->st); if (r == PBSGPU_OK) r = rx; }
        if (r == PBSGPU_OK) {
            uint64_t nch = j->h_counters[1];
            if (produced + nch > cap) overflow = true;
            else {
                for (uint64_t k = 0; k < nch; k++) {
                    pbsgpu_chunk c = j->h_out[k];
                    c.stream += groups[gi].first;
                    out[produced + k] = c;
                }
            }
            produced += nch;
        } else cudaStreamSynchronize(j->st);
        job_release(j);
        jobs[gi] = nullptr;
        return r;
    };
    for (size_t gi = 0; gi < groups.size() && rc == PBSGPU_OK; gi++) {
        if (gi >= (size_t)nbuf) { rc = collect(gi - nbuf); if (rc) break; }   // frees the buffer we are about to reuse
        const Group &g = groups[gi];
        uint8_t *buf = bufs[gi % nbuf];
        std::vector<uint64_t> goff(g.count), glen(g.count);
        uint64_t pos = 0;
        for (uint32_t k = 0; k < g.count; k++) {
            uint32_t i = g.first + k;
            goff[k] = pos; glen[k] = len[i];
            if (len[i]) {
                cudaError_t e = cudaMemcpyAsync(buf + pos, base + off[i], len[i], cudaMemcpyHostToDevice, ctx->copy_stream);
                if (e != cudaSuccess) { rc = fail(ctx, PBSGPU_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e)); break; }
            }
            pos += al(len[i]);
        }
        if (rc) break;
        cudaEventRecord(copied[gi], ctx->copy_stream);
        pbsgpu_job *j = nullptr;
        rc = job_create(ctx, cfg, buf, goff.data(), glen.data(), g.count, 1, 1, &j);
        if (rc) break;
        cudaStreamWaitEvent(j->st, copied[gi], 0);
        rc = job_enqueue(j);
        if (rc == PBSGPU_OK && xxh3_out) rc = xxh3_enqueue(ctx, buf, goff.data(), glen.data(), g.count, j->st, &xruns[gi]);
        if (rc) { cudaStreamSynchronize(j->st); job_release(j); break; }
        jobs[gi] = j;
    }
    for (size_t gi = 0; gi < groups.size(); gi++)
        if (jobs[gi]) { int r = collect(gi); if (rc == PBSGPU_OK) rc = r; }
    cudaStreamSynchronize(ctx->copy_stream);
    for (auto &e : copied) cudaEventDestroy(e);
    for (int b = 0; b < nbuf; b++) ctx->dev.put(bufs[b]);
    if (n_out) *n_out = produced;
    if (rc == PBSGPU_OK && overflow)
        rc = fail(ctx, PBSGPU_ERANGE, "output capacity %llu < %llu chunks", (unsigned long long)cap, (unsigned long long)produced);
    return rc;
}

extern "C" int pbsgpu_chunk_digest_batch_xxh3(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const void *base, const uint64_t *off,
                                              const uint64_t *len, uint32_t n, pbsgpu_set *set, pbsgpu_chunk *out,
                                              uint64_t cap, uint64_t *n_out, uint64_t *stream_xxh3) {
    if (!ctx || (n && (!off || !len || !base)) || (cap && !out)) return PBSGPU_EINVAL;
    if (set && set->ctx != ctx) return fail(ctx, PBSGPU_EINVAL, "set belongs to another context");
    Guard g(ctx);
    if (!cfg_ok(cfg)) return fail(ctx, PBSGPU_EINVAL, "invalid pbsgpu_cfg (use pbsgpu_config)");
    uint64_t produced = 0;
    int rc;
    if (n == 0) { if (n_out) *n_out = 0; return PBSGPU_OK; }
    if (is_device_ptr(base)) {
        pbsgpu_job *j = nullptr;
        rc = job_create(ctx, cfg, base, off, len, n, 1, 1, &j);
        if (rc) return rc;
        XxhRun xr;
        rc = job_enqueue(j);
        if (rc == PBSGPU_OK && stream_xxh3) rc = xxh3_enqueue(ctx, (const uint8_t *)base, off, len, n, j->st, &xr);
        if (rc == PBSGPU_OK) rc = job_finish(j);
        if (rc == PBSGPU_OK) {
            produced = j->h_counters[1];
            if (produced > cap) rc = fail(ctx, PBSGPU_ERANGE, "output capacity %llu < %llu chunks", (unsigned long long)cap, (unsigned long long)produced);
            else if (produced) memcpy(out, j->h_out, produced * sizeof(pbsgpu_chunk));
        } else cudaStreamSynchronize(j->st);
        if (stream_xxh3) { int rx = xxh3_collect(ctx, &xr, stream_xxh3, j->st); if (rc == PBSGPU_OK) rc = rx; }
        job_release(j);
    } else {
        rc = batch_host(ctx, cfg, (const uint8_t *)base, off, len, n, out, cap, &produced, stream_xxh3);
    }
    if (n_out) *n_out = produced;
    if (rc == PBSGPU_OK && set) rc = apply_set(set, out, produced);
    return rc;
}

extern "C" int pbsgpu_chunk_digest_batch(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const void *base, const uint64_t *off,
                                         const uint64_t *len, uint32_t n, pbsgpu_set *set, pbsgpu_chunk *out,
                                         uint64_t cap, uint64_t *n_out) {
    return pbsgpu_chunk_digest_batch_xxh3(ctx, cfg, base, off, len, n, set, out, cap, n_out, nullptr);
}

extern "C" int pbsgpu_scan_batch(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, const void *base, const uint64_t *off,
                                 const uint64_t *len, uint32_t n, uint64_t *ends, uint64_t cap, uint64_t *stream_first,
                                 uint64_t *n_out) {
    if (!ctx || (n && (!off || !len || !base)) || !stream_first) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (!is_device_ptr(base) && n) return fail(ctx, PBSGPU_EINVAL, "pbsgpu_scan_batch needs a device pointer");
    pbsgpu_job *j = nullptr;
    int rc = job_create(ctx, cfg, base, off, len, n, 1, 0, &j);
    if (rc) return rc;
    rc = job_enqueue(j);
    if (rc == PBSGPU_OK) rc = job_finish(j);
    if (rc == PBSGPU_OK) {
        uint64_t nch = j->h_counters[1];
        if (n_out) *n_out = nch;
        if (nch > cap) rc = fail(ctx, PBSGPU_ERANGE, "output capacity %llu < %llu chunks", (unsigned long long)cap, (unsigned long long)nch);
        else {
            std::vector<ChunkRef> refs(nch);
            std::vector<uint64_t> first(n + 1);
            cudaError_t e = cudaMemcpy(refs.data(), j->d_chunks, nch * sizeof(ChunkRef), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(first.data(), j->d_chunk_first, (n + 1) * 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = fail(ctx, PBSGPU_ECUDA, "D2H: %s", cudaGetErrorString(e));
            else {
                for (uint64_t k = 0; k < nch; k++) ends[k] = refs[k].start + refs[k].len;
                memcpy(stream_first, first.data(), (n + 1) * 8);
            }
        }
    } else cudaStreamSynchronize(j->st);
    job_release(j);
    return rc;
}

extern "C" int pbsgpu_sha256_batch(pbsgpu_ctx *ctx, const void *base, const uint64_t *off, const uint64_t *len,
                                   uint32_t n, uint8_t *digests) {
    if (!ctx || (n && (!off || !len || !digests))) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (n == 0) return PBSGPU_OK;
    cudaStream_t st = ctx->streams[0];
    std::vector<ChunkRef> refs(n);
    uint64_t hi = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (len[i] >= (1ull << 32)) return fail(ctx, PBSGPU_EINVAL, "range %u longer than 4 GiB", i);
        refs[i].stream = i; refs[i].len = (uint32_t)len[i]; refs[i].start = off[i];
        hi = std::max(hi, off[i] + len[i]);
    }
    const uint8_t *dbase = (const uint8_t *)base;
    uint8_t *staged = nullptr;
    if (!is_device_ptr(base)) {
        staged = (uint8_t *)ctx->dev.get(hi + 16);
        if (!staged) return fail(ctx, PBSGPU_ENOMEM, "staging of %llu bytes failed", (unsigned long long)hi);
        CK(cudaMemcpyAsync(staged, base, hi, cudaMemcpyHostToDevice, st));
        dbase = staged;
    }
    ChunkRef *d_refs = (ChunkRef *)ctx->dev.get(sizeof(ChunkRef) * n);
    uint8_t *d_dig = (uint8_t *)ctx->dev.get((uint64_t)n * 32);
    unsigned long long *d_n = (unsigned long long *)ctx->dev.get(8);
    int rc = PBSGPU_OK;
    if (!d_refs || !d_dig || !d_n) rc = fail(ctx, PBSGPU_ENOMEM, "device allocation failed");
    else {
        unsigned long long hn = n;
        cudaError_t e = cudaMemcpyAsync(d_refs, refs.data(), sizeof(ChunkRef) * n, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, &hn, 8, cudaMemcpyHostToDevice, st);
        ShaArgs ha;
        ha.base = dbase; ha.off = nullptr; ha.chunks = d_refs; ha.order = nullptr; ha.n_chunks = d_n; ha.chunk_cap = n;
        ha.digests = d_dig; ha.n_head = nullptr; ha.part = 0;
        if (e == cudaSuccess) e = ctx->variant == 1 ? launch_sha_simple(ha, st) : launch_sha_tuned(ha, ctx->sm_count, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(digests, d_dig, (uint64_t)n * 32, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { (void)cudaGetLastError(); rc = fail(ctx, PBSGPU_ECUDA, "sha256 batch: %s", cudaGetErrorString(e)); }
    }
    ctx->dev.put(d_refs); ctx->dev.put(d_dig); ctx->dev.put(d_n); ctx->dev.put(staged);
    return rc;
}

// ---------------------------------------------------------------------------
// streaming form
// ---------------------------------------------------------------------------
// Windows are pipelined: when a window is full its scan + resolve run at once (the host needs the
// cut points to know which tail is still undecided and must be carried into the next window), the carry
// is copied to the next buffer, and the window's SHA-256 half is enqueued asynchronously -- so the serial
// tail of a window's longest chunk overlaps the copies and scans of the following windows.  poll()
// hands out the chunks of finished windows in stream order.
constexpr int STREAM_NBUF = 6;   // windows in flight hide the ~0.3 s serial SHA tail of a window's longest chunk
struct StreamJob { pbsgpu_job *j; int buf; uint64_t base_off; };

struct pbsgpu_stream {
    pbsgpu_ctx *ctx;
    pbsgpu_cfg cfg;
    pbsgpu_set *set;
    uint64_t window;        // process when this many bytes are buffered
    uint64_t cap;           // device buffer capacity = window + max
    uint8_t *buf[STREAM_NBUF];
    bool busy[STREAM_NBUF]; // referenced by an in-flight window
    int cur;
    uint64_t fill;          // bytes buffered in buf[cur]
    uint64_t base_off;      // stream offset of buf[cur][0]
    bool finished, started;
    std::vector<StreamJob> inflight;   // FIFO
    std::vector<pbsgpu_chunk> ready;
    size_t ready_pos;
};

extern "C" int pbsgpu_stream_open(pbsgpu_ctx *ctx, const pbsgpu_cfg *cfg, pbsgpu_set *set, pbsgpu_stream **out) {
    if (!ctx || !out) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (!cfg_ok(cfg)) return fail(ctx, PBSGPU_EINVAL, "invalid pbsgpu_cfg (use pbsgpu_config)");
    pbsgpu_stream *s = new pbsgpu_stream();
    s->ctx = ctx; s->cfg = *cfg; s->set = set;
    uint64_t w = 1ull << 30;   // bytes in flight (5 windows) / tail latency bounds the rate: ~15 GiB/s
    const char *e = getenv("PBSGPU_STREAM_WINDOW");
    if (e) w = strtoull(e, nullptr, 0);
    s->window = std::max<uint64_t>(w, (uint64_t)cfg->max);
    s->cap = 0; s->cur = 0; s->fill = 0; s->base_off = 0;
    for (int i = 0; i < STREAM_NBUF; i++) { s->buf[i] = nullptr; s->busy[i] = false; }
    s->finished = false; s->started = false; s->ready_pos = 0;
    *out = s;
    return PBSGPU_OK;
}

// collect finished windows (all of them if block) in order
static int stream_collect(pbsgpu_stream *s, bool block) {
    while (!s->inflight.empty()) {
        StreamJob sj = s->inflight.front();
        if (!block) {
            cudaError_t q = cudaEventQuery(sj.j->ev[EV_END]);
            if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }
        }
        int rc = job_finish(sj.j);
        if (rc != PBSGPU_OK) { cudaStreamSynchronize(sj.j->st); }
        s->inflight.erase(s->inflight.begin());
        s->busy[sj.buf] = false;
        if (rc != PBSGPU_OK) { job_release(sj.j); return rc; }
        const uint64_t nch = sj.j->h_counters[1];
        const size_t before = s->ready.size();
        for (uint64_t k = 0; k < nch; k++) {
            pbsgpu_chunk c = sj.j->h_out[k];
            c.stream = 0; c.end_off += sj.base_off;
            s->ready.push_back(c);
        }
        job_release(sj.j);
        if (s->set && nch) { rc = apply_set(s->set, s->ready.data() + before, nch); if (rc) return rc; }
    }
    return PBSGPU_OK;
}

static int stream_process(pbsgpu_stream *s, int eof) {
    pbsgpu_ctx *ctx = s->ctx;
    if (s->fill == 0) return PBSGPU_OK;
    uint64_t off0 = 0, len0 = s->fill;
    pbsgpu_job *j = nullptr;
    int rc = job_create(ctx, &s->cfg, s->buf[s->cur], &off0, &len0, 1, eof, 1, &j);
    if (rc) return rc;
    // front half now: the cut points decide what has to be carried over
    unsigned long long counters[4] = {0, 0, 0, 0};
    uint64_t consumed = s->fill;
    for (;;) {
        rc = job_enqueue_front(j);
        if (rc == PBSGPU_OK) {
            cudaError_t e = cudaMemcpyAsync(counters, j->d_counters, sizeof counters, cudaMemcpyDeviceToHost, j->st);
            if (e == cudaSuccess && !eof) e = cudaMemcpyAsync(&consumed, j->d_consumed, 8, cudaMemcpyDeviceToHost, j->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(j->st);
            if (e != cudaSuccess) { (void)cudaGetLastError(); rc = fail(ctx, PBSGPU_ECUDA, "stream window: %s", cudaGetErrorString(e)); }
        }
        if (rc != PBSGPU_OK) { cudaStreamSynchronize(j->st); job_release(j); return rc; }
        if (counters[0] <= j->cand_cap) break;
        rc = job_grow_cands(j, counters[0]);          // dense candidates: redo the front half with room for all
        if (rc) { job_release(j); return rc; }
        j->reruns++;
    }
    if (eof) consumed = s->fill;
    const uint64_t rest = s->fill - consumed;
    // next buffer (wait for the oldest window if all are referenced)
    int next = -1;
    for (;;) {
        for (int i = 0; i < STREAM_NBUF; i++) if (i != s->cur && !s->busy[i]) { next = i; break; }
        if (next >= 0 || s->inflight.empty()) break;
        StreamJob oldest = s->inflight.front();
        cudaEventSynchronize(oldest.j->ev[EV_END]);
        rc = stream_collect(s, false);
        if (rc) { cudaStreamSynchronize(j->st); job_release(j); return rc; }
    }
    if (next < 0) { cudaStreamSynchronize(j->st); job_release(j); return fail(ctx, PBSGPU_ESTATE, "internal: no free stream buffer"); }
    if (!s->buf[next]) {
        s->buf[next] = (uint8_t *)ctx->dev.get(s->cap);
        if (!s->buf[next]) { cudaStreamSynchronize(j->st); job_release(j); return fail(ctx, PBSGPU_ENOMEM, "stream buffer of %llu bytes failed", (unsigned long long)s->cap); }
    }
    if (rest) {
        CK(cudaMemcpyAsync(s->buf[next], s->buf[s->cur] + consumed, rest, cudaMemcpyDeviceToDevice, ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
    }
    rc = job_enqueue_back(j);                          // SHA-256 etc. run while the next window fills
    if (rc != PBSGPU_OK) { cudaStreamSynchronize(j->st); job_release(j); return rc; }
    s->busy[s->cur] = true;
    s->inflight.push_back(StreamJob{j, s->cur, s->base_off});
    s->cur = next; s->fill = rest; s->base_off += consumed;
    return PBSGPU_OK;
}

extern "C" int pbsgpu_stream_write(pbsgpu_stream *s, const void *data, uint64_t len) {
    if (!s || (len && !data)) return PBSGPU_EINVAL;
    pbsgpu_ctx *ctx = s->ctx;
    Guard g(ctx);
    if (s->finished) return fail(ctx, PBSGPU_ESTATE, "stream already finished");
    if (!s->started) {
        s->cap = s->window + s->cfg.max + 256;
        s->buf[0] = (uint8_t *)ctx->dev.get(s->cap);      // further buffers are allocated when first needed
        if (!s->buf[0]) return fail(ctx, PBSGPU_ENOMEM, "stream buffer of %llu bytes failed", (unsigned long long)s->cap);
        s->started = true;
    }
    const uint8_t *p = (const uint8_t *)data;
    while (len) {
        uint64_t room = s->cap - s->fill;
        uint64_t take = std::min(len, std::min(room, s->window > s->fill ? s->window - s->fill : 0));
        if (take == 0) {   // window full: cut what can be cut, keep the undecided tail
            int rc = stream_process(s, 0);
            if (rc) return rc;
            if (s->fill >= s->window) return fail(ctx, PBSGPU_ESTATE, "internal: stream window did not drain");
            continue;
        }
        // ordered on the copy stream and completed before any kernel may read it
        CK(cudaMemcpyAsync(s->buf[s->cur] + s->fill, p, take, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        s->fill += take; p += take; len -= take;
    }
    if (s->fill >= s->window) return stream_process(s, 0);
    return PBSGPU_OK;
}

extern "C" int pbsgpu_stream_finish(pbsgpu_stream *s) {
    if (!s) return PBSGPU_EINVAL;
    Guard g(s->ctx);
    if (s->finished) return PBSGPU_OK;
    int rc = stream_process(s, 1);
    if (rc == PBSGPU_OK) rc = stream_collect(s, true);
    if (rc == PBSGPU_OK) s->finished = true;
    return rc;
}

extern "C" int pbsgpu_stream_poll(pbsgpu_stream *s, pbsgpu_chunk *out, uint64_t cap, uint64_t *n_out) {
    if (!s || !n_out || (cap && !out)) return PBSGPU_EINVAL;
    Guard g(s->ctx);
    int rc = stream_collect(s, false);
    if (rc) return rc;
    uint64_t avail = s->ready.size() - s->ready_pos;
    uint64_t k = std::min(avail, cap);
    if (k) memcpy(out, s->ready.data() + s->ready_pos, k * sizeof(pbsgpu_chunk));
    s->ready_pos += k;
    if (s->ready_pos == s->ready.size()) { s->ready.clear(); s->ready_pos = 0; }
    *n_out = k;
    return PBSGPU_OK;
}

extern "C" void pbsgpu_stream_close(pbsgpu_stream *s) {
    if (!s) return;
    Guard g(s->ctx);
    for (auto &sj : s->inflight) { cudaEventSynchronize(sj.j->ev[EV_END]); cudaStreamSynchronize(sj.j->st); job_release(sj.j); }
    for (int i = 0; i < STREAM_NBUF; i++) s->ctx->dev.put(s->buf[i]);
    delete s;
}

// ---------------------------------------------------------------------------
extern "C" int pbsgpu_corpus_fill(pbsgpu_ctx *ctx, const pbsgpu_corpus *c, uint64_t first_file, uint32_t n_files,
                                  void *dst_dev, uint64_t stride) {
    if (!ctx || !c || !dst_dev) return PBSGPU_EINVAL;
    Guard g(ctx);
    if (c->block_len == 0 || c->block_len % 8 || c->run_blocks == 0 || ((uintptr_t)dst_dev & 7) || (stride & 7) ||
        stride < c->file_len)
        return fail(ctx, PBSGPU_EINVAL, "corpus: block_len %% 8, run_blocks >= 1, 8-byte aligned dst/stride >= file_len required");
    CK(launch_corpus_fill(*c, first_file, n_files, (uint8_t *)dst_dev, stride, ctx->streams[0]));
    CK(cudaStreamSynchronize(ctx->streams[0]));
    return PBSGPU_OK;
}
