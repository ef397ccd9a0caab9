"""Generates tests/golden/*.json(l).  Run from the repo root:  python tests/golden/make_golden.py

The vectors come from the RESTATED oracle (oracle/oracle.c) -- the reference's own
arithmetic (Go module github.com/pbs-plus/pxar v0.19.2) is absent and there is no Go
toolchain, so these freeze the oracle's behaviour; they do not pin parity with Go.
Format = SURVEY.md section 8c (one JSON record per case) so that a ~40-line Go
program calling the real buzhash/transfer packages can emit the same file for a diff.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

OUT = Path(__file__).parent

t = oracle.default_table()
(OUT / "table_fingerprint.json").write_text(json.dumps({
    "what": "sha256 over the 256 table entries as little-endian u32",
    "sha256_le_u32": hashlib.sha256(t.astype("<u4").tobytes()).hexdigest(),
    "first": hex(int(t[0])), "last": hex(int(t[255])),
}, indent=1) + "\n")

cases = [  # (seed, file_id, len, avg, block_len)
    (1, 0, 1 << 20, 4096, 1 << 16),
    (1, 1, (1 << 20) + 12345, 4096, 1 << 16),
    (2, 0, 1 << 22, 65536, 1 << 16),
    (2, 5, 300_000, 1024, 1 << 12),
    (3, 0, 8 << 20, 1 << 20, 1 << 20),
    (1, 0, 48 << 20, 4 << 20, 4 << 20),   # the production configuration: 4 MiB average
    (4, 0, 100, 256, 8),
]
with open(OUT / "chunks_oracle.jsonl", "w") as f:
    for seed, fid, ln, avg, bl in cases:
        c = oracle.corpus(seed=seed, file_len=ln, block_len=bl)
        data = oracle.corpus_file(c, fid)
        rec = oracle.chunk_digest(oracle.config(avg), data)
        f.write(json.dumps({
            "gen": "pbsgpu-corpus-v1(fmix64)", "oracle": "restatement", "seed": seed, "file_id": fid,
            "len": ln, "avg": avg, "block_len": bl,
            "data_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
            "cuts": rec["end_off"].tolist(), "digests": [bytes(x).hex() for x in rec["digest"]],
        }) + "\n")
print("golden written")
