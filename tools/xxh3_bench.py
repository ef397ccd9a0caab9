#!/usr/bin/env python
"""K7 alone: XXH3-64 of N x 64 MiB files resident in HBM (pbsgpu_xxh3_batch), and of many small files.
usage: python tools/xxh3_bench.py [total_gib]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import xxhash
import pbs_plus_b200 as pg

tg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = pg.Engine(0)
buf = torch.empty(tg << 30, dtype=torch.uint8, device="cuda")
flen = 64 << 20
nf = (tg << 30) // flen
eng.corpus_fill(pg.corpus(seed=13, file_len=flen), 0, nf, buf, flen)


def run(name, off, ln, check):
    eng.xxh3_batch(buf, off, ln)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); h = eng.xxh3_batch(buf, off, ln); best = min(best, time.perf_counter() - t0)
    ok = all(xxhash.xxh3_64_intdigest(buf[int(off[i]): int(off[i] + ln[i])].cpu().numpy().tobytes()) == int(h[i]) for i in check)
    print(f"{name}: {len(off)} ranges, {ln.sum() / 2**30:.1f} GiB: {best * 1e3:.2f} ms  {ln.sum() / best / 1e9:.0f} GB/s  ok={ok}")


off = np.arange(nf, dtype=np.uint64) * flen
run("64 MiB files", off, np.full(nf, flen, dtype=np.uint64), [0, nf // 2, nf - 1])
n = (tg << 30) // (1 << 20)
run("1 MiB files", np.arange(n, dtype=np.uint64) << 20, np.full(n, (1 << 20) - 3, dtype=np.uint64), [0, 7, n - 1])
n = (tg << 30) // 16384
run("16 KiB files", np.arange(n, dtype=np.uint64) * 16384, np.full(n, 16000, dtype=np.uint64), [0, 5, n - 1])
run("one 8 GiB file", np.array([0], dtype=np.uint64), np.array([min(8, tg) << 30], dtype=np.uint64), [])
