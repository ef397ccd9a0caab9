"""Generates tests/golden/xxh3_libxxhash.jsonl.  Run from the repo root:  python tests/golden/make_xxh3_golden.py

Unlike chunks_oracle.jsonl these vectors come from an INDEPENDENT implementation: python-xxhash (the libxxhash
binding) computes XXH3-64 (seed 0) -- the function `xxh3.New() ... Sum64()` of the reference's commit walk
(internal/pxarmount/commit.go:717-725) -- over inputs from the repository's deterministic corpus generator.
They pin the oracle's restatement and the CUDA kernel (K7) for row f2 also where python-xxhash is absent.
"""
import json
import sys
from pathlib import Path

import xxhash

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402  (only its corpus generator is used here)

OUT = Path(__file__).parent
LENS = ([0, 1, 2, 3, 4, 5, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 96, 97, 128, 129, 130, 239, 240, 241, 255, 256, 300,
         511, 512, 1023, 1024, 1025, 1087, 1088, 1089, 2047, 2048, 2049, 3000, 4096, 65535, 65536, 65537]
        + [(1 << 20) - 1, 1 << 20, (1 << 20) + 1, 3_000_003, 8 << 20])
with open(OUT / "xxh3_libxxhash.jsonl", "w") as f:
    for i, n in enumerate(LENS):
        for lead in (0, 3):
            c = oracle.corpus(seed=40 + i, file_len=max(n + lead, 1), block_len=4096)
            data = oracle.corpus_file(c, 0)[lead:lead + n]
            f.write(json.dumps({"gen": "pbsgpu-corpus-v1(fmix64)", "by": f"python-xxhash {xxhash.VERSION} / libxxhash "
                                f"{xxhash.XXHASH_VERSION}", "seed": 40 + i, "block_len": 4096, "lead": lead, "len": n,
                                "xxh3_64": f"{xxhash.xxh3_64_intdigest(data.tobytes()):016x}"}) + "\n")
print("xxh3 golden written")
