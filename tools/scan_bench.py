#!/usr/bin/env python
"""K1 alone: boundaries of N x 64 MiB files (pbsgpu_scan_batch), wall time of the call minus nothing --
the resolve/sort part is < 0.3 ms.  usage: PBSGPU_SCAN_LANES=0|1 python tools/scan_bench.py [total_gib]
Prints a checksum of the boundary list so two variants can be compared byte for byte."""
import os, sys, time, hashlib
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import pbs_plus_b200 as pg

tg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = pg.Engine(0)
cfg = pg.make_config(4 << 20)
flen = 64 << 20
n = (tg << 30) // flen
buf = torch.empty(tg << 30, dtype=torch.uint8, device="cuda")
eng.corpus_fill(pg.corpus(seed=11, file_len=flen), 0, n, buf, flen)
off = np.arange(n, dtype=np.uint64) * flen
ln = np.full(n, flen, dtype=np.uint64)
ends, first = eng.scan_batch(cfg, buf, off, ln)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); ends, first = eng.scan_batch(cfg, buf, off, ln); dt = time.perf_counter() - t0
    best = min(best, dt)
h = hashlib.sha256(ends.tobytes() + first.tobytes()).hexdigest()[:16]
print(f"scan_lanes={os.environ.get('PBSGPU_SCAN_LANES','0')} part={os.environ.get('PBSGPU_PARTITION_SMS','default')} "
      f"{n} x 64 MiB: {best*1e3:.2f} ms  {(tg<<30)/best/1e9:.0f} GB/s  chunks={len(ends)} sum={h}")
