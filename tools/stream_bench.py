#!/usr/bin/env python
"""Streaming form (one io.Reader-like stream): GiB/s from pageable and pinned host memory, through write() and through
the zero-copy reserve/commit ring.  An 8 GiB stream ends with the serial SHA-256 chain of its last window's longest chunk
(~0.3 s for 16 MiB) fully exposed, so a 64 GiB stream (the 8 GiB source 8 times) is timed too, and the time at which the
last write returned is printed beside the total."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import pbs_plus_b200 as pg
eng = pg.Engine(0)
n = 8 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.corpus_fill(pg.corpus(seed=4, file_len=n), 0, 1, dev, n)
pinned = eng.host_alloc(n)
torch.from_numpy(np.asarray(pinned)).copy_(dev)
pageable = np.array(pinned)
cfg = pg.buzhash.NewConfig(4096)
ref = eng.chunk_digest_batch(cfg, dev, [0], [n])
del dev
torch.cuda.empty_cache()


def run(name, src, piece, repeat, mode):
    st = eng.stream(cfg)
    t0 = time.perf_counter(); got = []
    for _ in range(repeat):
        for i in range(0, n, piece):
            if mode == "write":
                st.write(src[i:i + piece])
            else:
                pos = i
                while pos < i + piece:
                    slot = st.reserve(); k = min(len(slot), i + piece - pos)
                    slot[:k] = src[pos:pos + k]; st.commit(k); pos += k
            got.append(st.poll())
    t_fed = time.perf_counter() - t0
    got.append(st.finish()); dt = time.perf_counter() - t0
    rec = np.concatenate(got); st.close()
    eq = rec.tobytes() == ref.tobytes() if repeat == 1 else int(rec["end_off"][-1]) == repeat * n
    print(f"stream {name:8s} {mode:7s} pieces of {piece >> 20:3d} MiB, {repeat * n >> 30:3d} GiB: {repeat * n / dt / 2**30:6.2f} GiB/s "
          f"(fed after {t_fed:.2f} s = {repeat * n / t_fed / 2**30:.1f} GiB/s, done after {dt:.2f} s), chunks {len(rec)}, ok={eq}", flush=True)


run("pinned", pinned, 64 << 20, 1, "write")      # warm-up (pools, ring)
for name, src in (("pinned", pinned), ("pageable", pageable)):
    for piece in (64 << 20, 4 << 20):
        run(name, src, piece, 1, "write")
run("pageable", pageable, 32 << 20, 1, "reserve")
run("pinned", pinned, 64 << 20, 8, "write")
run("pageable", pageable, 32 << 20, 8, "reserve")
