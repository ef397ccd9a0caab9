#!/usr/bin/env python
"""Streaming form (one io.Reader-like stream): GiB/s from pageable and pinned host memory."""
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
for name, src in (("pinned", pinned), ("pageable", pageable)):
    for piece in (64 << 20, 4 << 20):
        st = eng.stream(cfg)
        t0 = time.perf_counter(); got = []
        for i in range(0, n, piece):
            st.write(src[i:i + piece]); got.append(st.poll())
        got.append(st.finish()); dt = time.perf_counter() - t0
        rec = np.concatenate(got); st.close()
        print(f"stream {name} writes of {piece >> 20} MiB: {n / dt / 2**30:.2f} GiB/s, chunks {len(rec)}, equal_to_batch={rec.tobytes() == ref.tobytes()}")
