// tests/cxx/driver.cpp -- C++ caller of the C ABI through the host mirror (include/pbsgpu.hpp),
// standing in for the Go caller that cannot be built here.  Reads files named on the command
// line, pushes them through transfer::DedupWriter twice (second pass: everything is known) and
// prints one line per chunk:  <pass> <path> <end_off> <digest hex> <known>, and `xxh3 <path> <hex>` per file.  tests/test_gpu_parity.py
// compares the output with the oracle.  Then the same files as ONE pxar payload stream (PayloadStreamWriter):
// `payload-offset <path> <off>` per file and `payload <end_off> <digest hex> <known>` per chunk.
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pbsgpu.hpp"

int main(int argc, char **argv) {
    try {
        int avg_kib = std::stoi(argv[1]);
        pbsgpu::Engine eng(0);
        auto cfg = pbsgpu::buzhash::NewConfig(avg_kib);
        pbsgpu::KnownSet known(eng);
        for (int pass = 0; pass < 2; pass++) {
            pbsgpu::transfer::DedupWriter w(eng, cfg, &known, 64ull << 20);
            for (int i = 2; i < argc; i++) {
                FILE *f = std::fopen(argv[i], "rb");
                if (!f) { std::fprintf(stderr, "open %s failed\n", argv[i]); return 2; }
                std::fseek(f, 0, SEEK_END); uint64_t size = (uint64_t)std::ftell(f); std::fseek(f, 0, SEEK_SET);
                w.WriteEntryReader({argv[i], size}, [f](uint8_t *b, size_t n) { return std::fread(b, 1, n, f); }, size);
                std::fclose(f);
            }
            for (auto &r : w.Finish()) {
                std::printf("%d %s %llu ", pass, r.path.c_str(), (unsigned long long)r.end_off);
                for (int k = 0; k < 32; k++) std::printf("%02x", r.digest[k]);
                std::printf(" %d\n", r.known ? 1 : 0);
            }
            if (pass == 0)   // backedHashes (commit.go:725): XXH3-64 per file from the same staged bytes
                for (auto &kv : w.BackedHashes()) std::printf("xxh3 %s %016llx\n", kv.first.c_str(), (unsigned long long)kv.second);
        }
        {   // the layout-faithful form: all files as ONE pxar payload stream through pbsgpu_stream_* (own known set)
            pbsgpu::KnownSet k2(eng);
            pbsgpu::transfer::PayloadStreamWriter pw(eng, cfg, &k2);
            for (int i = 2; i < argc; i++) {
                FILE *f = std::fopen(argv[i], "rb");
                if (!f) return 2;
                std::fseek(f, 0, SEEK_END); uint64_t size = (uint64_t)std::ftell(f); std::fseek(f, 0, SEEK_SET);
                uint64_t off = pw.WriteEntryReader({argv[i], size}, [f](uint8_t *b, size_t n) { return std::fread(b, 1, n, f); }, size);
                std::fclose(f);
                std::printf("payload-offset %s %llu\n", argv[i], (unsigned long long)off);
            }
            for (auto &c : pw.Finish()) {
                std::printf("payload %llu ", (unsigned long long)c.end_off);
                for (int k = 0; k < 32; k++) std::printf("%02x", c.digest[k]);
                std::printf(" %d\n", c.known ? 1 : 0);
            }
        }
        try { pbsgpu::buzhash::NewConfig(3000); std::printf("ERR expected\n"); return 3; }
        catch (const pbsgpu::Error &e) { std::printf("config-error %d\n", e.code); }
    } catch (const pbsgpu::Error &e) {
        std::fprintf(stderr, "pbsgpu error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
