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


# -- f3: DataBlobs (upstream pbs-datastore data_blob.rs / file_formats.rs, restated; the magics are
#    sha256("Proxmox Backup uncompressed blob v1.0")[:8] and sha256("Proxmox Backup zstd compressed blob v1.0")[:8],
#    which tests/test_oracle.py re-derives) -----------------------------------------------------------
BLOB_MAGIC_UNCOMPRESSED = bytes([66, 171, 56, 7, 190, 131, 112, 161])
BLOB_MAGIC_COMPRESSED = bytes([49, 185, 88, 66, 111, 182, 163, 127])
ZFRAME_BLOCK = 128 * 1024


def zstd_frame_rle_raw(data: bytes) -> bytes:
    """A standard zstd frame (RFC 8878 3.1.1) without match / entropy stages: Frame_Header = magic 0xFD2FB528 LE,
    descriptor 0xE0 (8-byte Frame_Content_Size, Single_Segment), content size; then one block per 128 KiB of input --
    an RLE_Block (type 1, Block_Size = repeat count, 1 byte) when the block is one repeated byte, else a Raw_Block
    (type 0).  Block_Header = Last_Block | type << 1 | Block_Size << 3, 3 bytes LE.  Empty input: one empty raw block."""
    data = bytes(data)
    out = bytearray(b"\x28\xb5\x2f\xfd\xe0" + len(data).to_bytes(8, "little"))
    nb = max(1, (len(data) + ZFRAME_BLOCK - 1) // ZFRAME_BLOCK)
    for b in range(nb):
        blk = data[b * ZFRAME_BLOCK:(b + 1) * ZFRAME_BLOCK]
        rle = len(blk) > 0 and blk.count(blk[:1]) == len(blk)
        hdr = int(b + 1 == nb) | ((1 if rle else 0) << 1) | (len(blk) << 3)
        out += hdr.to_bytes(3, "little") + (blk[:1] if rle else blk)
    return bytes(out)


def blob_encode(data: bytes, compress: bool = True) -> bytes:
    """DataBlob { magic[8], crc32 LE over the payload, payload }; the zstd payload only when it is smaller than the raw
    bytes (upstream `DataBlob::encode`)."""
    import zlib
    data = bytes(data)
    if compress and data:
        fr = zstd_frame_rle_raw(data)
        if len(fr) < len(data):
            return BLOB_MAGIC_COMPRESSED + zlib.crc32(fr).to_bytes(4, "little") + fr
    return BLOB_MAGIC_UNCOMPRESSED + zlib.crc32(data).to_bytes(4, "little") + data
