#!/usr/bin/env python
"""K6 alone: CRC-32 of N equal device-resident ranges; prints GB/s."""
import sys, time, zlib
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import pbs_plus_b200 as pg
eng = pg.Engine(0)
tg, rk = 16, 4096
n = (tg << 30) // (rk << 10)
buf = torch.empty(tg << 30, dtype=torch.uint8, device="cuda")
eng.corpus_fill(pg.corpus(seed=9, file_len=tg << 30), 0, 1, buf, tg << 30)
off = np.arange(n, dtype=np.uint64) * (rk << 10); ln = np.full(n, rk << 10, dtype=np.uint64)
eng.crc32_batch(buf, off, ln); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); c = eng.crc32_batch(buf, off, ln); best = min(best, time.perf_counter() - t0)
ok = int(c[3]) == zlib.crc32(buf[int(off[3]): int(off[3]) + int(ln[3])].cpu().numpy().tobytes())
print(f"crc32: {n} x {rk} KiB: {ln.sum()/best/1e9:.1f} GB/s ({best*1e3:.1f} ms) ok={ok}")
