// pbsgpu.hpp -- C++ host mirror of the reference's Go surface for the hot path, over the C ABI.
//
// The reference is compiled Go and the build image has no Go toolchain, so the host side above
// the C ABI is written in C++ (header only).  Names, argument meaning and error behaviour follow
// the Go call sites in the reference (internal/pxarmount/commit.go):
//     buzhash.NewConfig(4096)                                   :302-305
//     transfer.NewRemoteDedupSplitArchiveWriter(..., origPayloadIdx)   :329
//     writer.WriteEntryReader(entry, reader, size)               :720, :858
//     writer.Finish()                                            :383
// Go `error` returns become pbsgpu::Error exceptions carrying the same text a Go caller would wrap.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "pbsgpu.h"

namespace pbsgpu {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

namespace buzhash {
using Config = pbsgpu_cfg;
// buzhash.NewConfig(avgKiB) (Config, error)
inline Config NewConfig(int avg_kib) {
    Config c;
    int rc = pbsgpu_config_kib((uint32_t)avg_kib, nullptr, &c);
    if (rc) throw Error(rc, "buzhash: invalid average chunk size " + std::to_string(avg_kib) + " KiB");
    return c;
}
}  // namespace buzhash

class Engine {
  public:
    explicit Engine(int device = 0) {
        int rc = pbsgpu_open(device, &ctx_);
        if (rc) throw Error(rc, "pbsgpu_open: no usable CUDA device (there is no CPU fallback)");
    }
    ~Engine() { pbsgpu_close(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    pbsgpu_ctx *ctx() const { return ctx_; }
    void check(int rc) const { if (rc) throw Error(rc, std::string("pbsgpu: ") + pbsgpu_strerror(ctx_)); }

  private:
    pbsgpu_ctx *ctx_ = nullptr;
};

class KnownSet {   // the session's known-chunk set (PreviousBackupRef, commit.go:286-294)
  public:
    explicit KnownSet(Engine &e, uint64_t hint = 1 << 16) : e_(e) { e.check(pbsgpu_set_create(e.ctx(), hint, &s_)); }
    ~KnownSet() { pbsgpu_set_destroy(s_); }
    pbsgpu_set *handle() const { return s_; }
    uint64_t SeedFromDidx(const std::vector<uint8_t> &didx) {   // origPayloadIdx, commit.go:324-328
        uint64_t n = 0;
        if (!didx.empty()) e_.check(pbsgpu_set_seed_didx(s_, didx.data(), didx.size(), &n));
        return n;
    }
    uint64_t size() const { uint64_t c = 0; pbsgpu_set_count(s_, &c); return c; }

  private:
    Engine &e_;
    pbsgpu_set *s_ = nullptr;
};

namespace transfer {

struct Entry { std::string Path; uint64_t FileSize; };   // the fields of pxar.Entry the hot path reads
struct IndexRecord { std::string path; uint64_t end_off; uint8_t digest[32]; bool known; };
// io.Reader: fill buf[0..n) and return the number of bytes produced (0 = EOF)
using Reader = std::function<size_t(uint8_t *buf, size_t n)>;

class DedupWriter {
  public:
    DedupWriter(Engine &e, const buzhash::Config &cfg, KnownSet *known, uint64_t staging_bytes = 1ull << 30)
        : e_(e), cfg_(cfg), known_(known), cap_(staging_bytes) {
        buf_ = (uint8_t *)pbsgpu_host_alloc(e.ctx(), cap_);
        if (!buf_) throw Error(PBSGPU_ENOMEM, "pbsgpu: pinned staging allocation failed");
    }
    ~DedupWriter() { pbsgpu_host_free(e_.ctx(), buf_); }

    // writer.WriteEntryReader(entry, reader, size): pulls exactly `size` bytes (io.ReadFull semantics)
    void WriteEntryReader(const Entry &entry, const Reader &reader, uint64_t size) {
        if (finished_) throw Error(PBSGPU_ESTATE, "transfer: writer already finished");
        uint64_t start = (fill_ + 255) & ~255ull;
        if (start + size > cap_) { Flush(); start = 0; }
        if (size > cap_) throw Error(PBSGPU_ENOMEM, "transfer: entry larger than the staging buffer: " + entry.Path);
        uint64_t got = 0;
        while (got < size) {
            size_t k = reader(buf_ + start + got, size - got);
            if (k == 0) throw Error(PBSGPU_EINVAL, "transfer: short read for " + entry.Path + ": unexpected EOF");
            got += k;
        }
        entries_.push_back(entry); off_.push_back(start); len_.push_back(size);
        fill_ = start + size;
    }
    void Flush() {
        if (entries_.empty()) return;
        uint64_t cap = entries_.size();
        for (uint64_t l : len_) cap += l / (cfg_.min > 64 ? cfg_.min : 65);
        std::vector<pbsgpu_chunk> out(cap + 1);
        std::vector<uint64_t> xxh(entries_.size());
        uint64_t n = 0;
        e_.check(pbsgpu_chunk_digest_batch_xxh3(e_.ctx(), &cfg_, buf_, off_.data(), len_.data(), (uint32_t)entries_.size(),
                                                known_ ? known_->handle() : nullptr, out.data(), out.size(), &n, xxh.data()));
        for (size_t i = 0; i < entries_.size(); i++) backed_hashes_[entries_[i].Path] = xxh[i];
        for (uint64_t i = 0; i < n; i++) {
            IndexRecord r;
            r.path = entries_[out[i].stream].Path; r.end_off = out[i].end_off;
            std::memcpy(r.digest, out[i].digest, 32); r.known = (out[i].flags & PBSGPU_CHUNK_KNOWN) != 0;
            index_.push_back(r);
        }
        entries_.clear(); off_.clear(); len_.clear(); fill_ = 0;
    }
    const std::vector<IndexRecord> &Finish() { Flush(); finished_ = true; return index_; }
    // relPath -> XXH3-64 of the uploaded bytes: commitWalkState.backedHashes (commit.go:187, :725)
    const std::map<std::string, uint64_t> &BackedHashes() const { return backed_hashes_; }

  private:
    Engine &e_;
    buzhash::Config cfg_;
    KnownSet *known_;
    uint8_t *buf_ = nullptr;
    uint64_t cap_, fill_ = 0;
    bool finished_ = false;
    std::vector<Entry> entries_;
    std::vector<uint64_t> off_, len_;
    std::vector<IndexRecord> index_;
    std::map<std::string, uint64_t> backed_hashes_;
};

// ---- the layout-faithful writer: ONE pxar v2 payload stream through pbsgpu_stream_* ---------------------------------
// Start marker, then per file a 16-byte PAYLOAD header + content (reference internal/pxarmount/pxarfs.go:408-411),
// chunker state carried across files, a suggested boundary at every file start, bytes read straight into the
// library's pinned ring.  WriteEntryReader (commit.go:720) returns the entry's payload offset.
struct PayloadChunk { uint64_t end_off; uint8_t digest[32]; bool known; };

class PayloadStreamWriter {
  public:
    PayloadStreamWriter(Engine &e, const buzhash::Config &cfg, KnownSet *known, bool suggest = true)
        : e_(e), suggest_(suggest) {
        e.check(pbsgpu_stream_open(e.ctx(), &cfg, known ? known->handle() : nullptr, &s_));
        raw(header(0x834c68c2194a4ed2ull, 16));                     // PXAR_PAYLOAD_START_MARKER
    }
    ~PayloadStreamWriter() { pbsgpu_stream_close(s_); }
    PayloadStreamWriter(const PayloadStreamWriter &) = delete;
    PayloadStreamWriter &operator=(const PayloadStreamWriter &) = delete;

    uint64_t WriteEntryReader(const Entry &entry, const Reader &reader, uint64_t size) {
        if (finished_) throw Error(PBSGPU_ESTATE, "transfer: writer already finished");
        const uint64_t off = pbsgpu_stream_position(s_);
        if (suggest_) e_.check(pbsgpu_stream_suggest(s_, off));
        raw(header(0x28147a1b0b7c1a25ull, 16 + size));              // PXAR_PAYLOAD
        const uint64_t slot = pbsgpu_stream_slot_bytes(s_);
        uint64_t left = size;
        while (left) {
            void *p = nullptr;
            e_.check(pbsgpu_stream_reserve(s_, &p));
            const uint64_t want = left < slot ? left : slot;
            uint64_t got = 0;
            while (got < want) {
                size_t k = reader((uint8_t *)p + got, want - got);
                if (k == 0) { pbsgpu_stream_commit(s_, 0); throw Error(PBSGPU_EINVAL, "transfer: short read for " + entry.Path + ": unexpected EOF"); }
                got += k;
            }
            e_.check(pbsgpu_stream_commit(s_, want));
            left -= want;
        }
        drain();
        return off;
    }
    const std::vector<PayloadChunk> &Finish() {
        if (!finished_) {
            raw(header(0x6c72b78b984c81b5ull, 16));                 // PXAR_PAYLOAD_TAIL_MARKER
            e_.check(pbsgpu_stream_finish(s_));
            drain();
            finished_ = true;
        }
        return index_;
    }

  private:
    static std::vector<uint8_t> header(uint64_t htype, uint64_t full) {
        std::vector<uint8_t> h(16);
        for (int i = 0; i < 8; i++) { h[i] = (uint8_t)(htype >> (8 * i)); h[8 + i] = (uint8_t)(full >> (8 * i)); }
        return h;
    }
    void raw(const std::vector<uint8_t> &b) { e_.check(pbsgpu_stream_write(s_, b.data(), b.size())); }
    void drain() {
        pbsgpu_chunk buf[256];
        for (;;) {
            uint64_t n = 0;
            e_.check(pbsgpu_stream_poll(s_, buf, 256, &n));
            if (!n) break;
            for (uint64_t i = 0; i < n; i++) {
                PayloadChunk c; c.end_off = buf[i].end_off; std::memcpy(c.digest, buf[i].digest, 32);
                c.known = (buf[i].flags & PBSGPU_CHUNK_KNOWN) != 0;
                index_.push_back(c);
            }
        }
    }
    Engine &e_;
    pbsgpu_stream *s_ = nullptr;
    bool suggest_, finished_ = false;
    std::vector<PayloadChunk> index_;
};

}  // namespace transfer
}  // namespace pbsgpu
