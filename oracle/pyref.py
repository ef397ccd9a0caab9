"""oracle/pyref.py -- second, independent restatement in numpy / pure Python.

TEST INFRASTRUCTURE.  Used only to cross-check oracle.c on small inputs: it is
written from the closed form of the window hash (vectorised over all positions),
not from the rolling recurrence, so a slip in either shows up as a mismatch.

Restates (see oracle.c header for provenance; PARITY UNPINNED vs. the Go module):
  H(i)  = XOR_{j=0..63} rotl32(T[b[i-j]], j mod 32)
  cut   : smallest L in [max(min,65),max] with L == max or (H(start+L-1) & mask) >= mask-2
          (65: upstream scan() never tests while the 64-byte window is being filled)
  digest: hashlib.sha256 of the raw chunk bytes
"""
from __future__ import annotations

import hashlib

import numpy as np


def rotl32(x: np.ndarray, r: int) -> np.ndarray:
    r &= 31
    if r == 0:
        return x.copy()
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def window_hashes(table: np.ndarray, data: np.ndarray) -> np.ndarray:
    """H(i) for every i in [0, len); entries with i < 63 are meaningless (window not full)."""
    t = table.astype(np.uint32)[data]                      # T[b[i]]
    n = len(data)
    h = np.zeros(n, dtype=np.uint32)
    for j in range(64):
        shifted = np.zeros(n, dtype=np.uint32)
        if j < n:
            shifted[j:] = t[: n - j]                         # T[b[i-j]]
        h ^= rotl32(shifted, j)
    return h


def candidates(table: np.ndarray, data: np.ndarray, mask: int) -> np.ndarray:
    """Positions i >= 63 whose window hash passes the break test."""
    h = window_hashes(table, data)
    ok = (h & np.uint32(mask)) >= np.uint32(mask - 2)
    ok[:63] = False
    return np.nonzero(ok)[0].astype(np.uint64)


def resolve(cands: np.ndarray, length: int, cmin: int, cmax: int) -> list[int]:
    """Sequential min/max rule over sorted candidate positions -> chunk END offsets."""
    ends: list[int] = []
    start = 0
    k = 0
    nc = len(cands)
    while start < length:
        lo = start + max(cmin, 65) - 1  # first position scan() tests: len >= min and window rolled once
        while k < nc and int(cands[k]) < lo:
            k += 1
        end = start + cmax
        if k < nc and int(cands[k]) + 1 <= end:
            end = int(cands[k]) + 1
        if end > length:
            end = length
        ends.append(end)
        start = end
    return ends


def chunk_ends(table: np.ndarray, data: np.ndarray, avg: int) -> list[int]:
    mask = 2 * avg - 1
    return resolve(candidates(table, data, mask), len(data), avg >> 2, avg << 2)


def chunk_digests(data: np.ndarray, ends: list[int]) -> list[bytes]:
    out, s = [], 0
    b = data.tobytes()
    for e in ends:
        out.append(hashlib.sha256(b[s:e]).digest())
        s = e
    return out


def splitmix64_stream(seed: int, nwords: int) -> np.ndarray:
    """Plain sequential splitmix64 (for documentation / golden format of SURVEY 8c)."""
    M = (1 << 64) - 1
    out = np.empty(nwords, dtype=np.uint64)
    x = seed & M
    for i in range(nwords):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        out[i] = z ^ (z >> 31)
    return out


DIDX_MAGIC = bytes([28, 145, 78, 165, 25, 186, 179, 205])


def didx_build(lengths, digests, uuid: bytes = b"\0" * 16, ctime: int = 0) -> bytes:
    """PBS dynamic index image (restated from upstream pbs-datastore dynamic_index.rs / file_formats.rs;
    UNVERIFIED against the Go module -- see oracle.c header): 4096-byte header {magic, uuid, ctime i64 LE,
    index_csum = SHA-256 over the entry table, zero padding} + entries {u64 end LE, digest[32]}."""
    body = bytearray()
    end = 0
    for ln, d in zip(lengths, digests):
        end += int(ln)
        body += end.to_bytes(8, "little") + bytes(d)
    hdr = bytearray(4096)
    hdr[0:8] = DIDX_MAGIC
    hdr[8:24] = bytes(uuid).ljust(16, b"\0")[:16]
    hdr[24:32] = int(ctime).to_bytes(8, "little", signed=True)
    hdr[32:64] = hashlib.sha256(bytes(body)).digest()
    return bytes(hdr) + bytes(body)
