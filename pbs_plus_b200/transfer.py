"""Mirror of the reference's dedup split-archive writer surface, for the hot path only.

Reference (Go, module github.com/pbs-plus/pxar/transfer, call sites in
internal/pxarmount/commit.go):
    writer, err := transfer.NewRemoteDedupSplitArchiveWriter(ctx, session, meta, payload, origPayloadIdx)   :329
    err = writer.WriteEntryReader(entry, reader, size)                                                      :720, :858
    err = writer.Finish()                                                                                   :383
Inside WriteEntryReader the module pulls the reader's bytes, runs the buzhash scan,
cuts, SHA-256s every chunk, asks the session whether the digest is known and uploads
only new chunks.  This mirror keeps the names and the pull-from-a-reader contract but
batches the files: entries are queued and pushed through the GPU engine in one batch on
Flush()/Finish().  pxar framing, blob encoding and the
HTTP/2 upload are out of scope (section 8 a5, f3): `upload` is a callback.
"""
from __future__ import annotations

import io
from dataclasses import dataclass, field
from typing import BinaryIO, Callable

import numpy as np

from ._lib import CHUNK_DTYPE, CHUNK_KNOWN, Cfg
from .engine import DigestSet, Engine


@dataclass
class Entry:
    """The two fields of pxar.Entry the hot path reads (commit.go:710-715)."""
    Path: str
    FileSize: int


@dataclass
class IndexRecord:
    """One dynamic-index entry the writer appends per chunk: (end offset, digest)."""
    path: str
    end_off: int
    digest: bytes
    known: bool


@dataclass
class DedupWriter:
    engine: Engine
    config: Cfg
    known: DigestSet | None = None
    upload: Callable[[bytes, bytes], None] | None = None   # (digest, chunk bytes) for NEW chunks
    batch_bytes: int = 1 << 30
    index: list[IndexRecord] = field(default_factory=list)
    # relPath -> XXH3-64 of the file as uploaded: commitWalkState.backedHashes (commit.go:187, :725), computed on the
    # GPU from the same staged bytes instead of through an io.TeeReader on the host
    backed_hashes: dict[str, int] = field(default_factory=dict)
    _pending: list[tuple[Entry, np.ndarray]] = field(default_factory=list)
    _pending_bytes: int = 0
    _finished: bool = False

    def WriteEntryReader(self, entry: Entry, reader: BinaryIO, size: int) -> None:
        """Pulls exactly `size` bytes from `reader` (error if it yields fewer, like io.ReadFull)."""
        if self._finished:
            raise RuntimeError("transfer: writer already finished")
        data = reader.read(size) if size else b""
        if len(data) != size:
            raise IOError(f"transfer: short read for {entry.Path}: got {len(data)} of {size} bytes (unexpected EOF)")
        self._pending.append((entry, np.frombuffer(data, dtype=np.uint8)))
        self._pending_bytes += size
        if self._pending_bytes >= self.batch_bytes:
            self.Flush()

    def WriteEntry(self, entry: Entry, data: bytes) -> None:
        self.WriteEntryReader(entry, io.BytesIO(data), len(data))

    def Flush(self) -> None:
        if not self._pending:
            return
        entries, arrs = zip(*self._pending)
        rec, hashes = self._batch(arrs)
        for e, h in zip(entries, hashes):
            self.backed_hashes[e.Path] = int(h)
        for r in rec:
            i = int(r["stream"])
            known = bool(r["flags"] & CHUNK_KNOWN)
            self.index.append(IndexRecord(entries[i].Path, int(r["end_off"]), bytes(r["digest"]), known))
        if self.upload is not None:
            starts: dict[int, int] = {}
            for r in rec:
                i = int(r["stream"])
                s = starts.get(i, 0)
                e = int(r["end_off"])
                if not (r["flags"] & CHUNK_KNOWN):
                    self.upload(bytes(r["digest"]), arrs[i][s:e].tobytes())
                starts[i] = e
        self._pending, self._pending_bytes = [], 0

    def _batch(self, arrs):
        lens = np.array([len(a) for a in arrs], dtype=np.uint64)
        offs = np.zeros(len(arrs), dtype=np.uint64)
        pos = 0
        for i, a in enumerate(arrs):
            offs[i] = pos
            pos += (len(a) + 255) & ~255
        buf = np.zeros(max(pos, 1), dtype=np.uint8)
        for a, o in zip(arrs, offs):
            buf[int(o): int(o) + len(a)] = a
        return self.engine.chunk_digest_batch_xxh3(self.config, buf, offs, lens, self.known)

    def Finish(self) -> list[IndexRecord]:
        self.Flush()
        self._finished = True
        return self.index


def NewRemoteDedupSplitArchiveWriter(engine: Engine, config: Cfg, known: DigestSet | None = None,
                                     orig_payload_idx: bytes | None = None, upload=None) -> DedupWriter:
    """commit.go:329.  `orig_payload_idx` = bytes of the previous .ppxar.didx (commit.go:324-328):
    its digests seed the known set, so unchanged chunks are not uploaded again."""
    if orig_payload_idx:
        if known is None:
            known = engine.digest_set()
        known.seed_didx(orig_payload_idx)
    return DedupWriter(engine, config, known, upload)


def verifyBackedFileHashes(engine: Engine, open_file: Callable[[str], BinaryIO], hashes: dict[str, int],
                           batch_bytes: int = 1 << 30) -> None:
    """Mirror of verifyBackedFileHashes (reference internal/pxarmount/commit.go:957-976): every backed file is read
    again and its XXH3-64 must equal the value recorded at upload time.  The reference hashes on the host through a
    64 KiB copy buffer, one file at a time; here the files are packed into batches and hashed by ONE
    pbsgpu_xxh3_batch call each.  Errors keep the reference's wording:
    `open backed file "<p>" for verification: ...` and `backed file "<p>" content hash differs`."""
    names: list[str] = []
    arrs: list[np.ndarray] = []
    held = 0

    def flush() -> None:
        nonlocal names, arrs, held
        if not names:
            return
        lens = np.array([len(a) for a in arrs], dtype=np.uint64)
        offs = np.zeros(len(arrs), dtype=np.uint64)
        pos = 0
        for i, a in enumerate(arrs):
            offs[i] = pos
            pos += (len(a) + 255) & ~255
        buf = np.zeros(max(pos, 1), dtype=np.uint8)
        for a, o in zip(arrs, offs):
            buf[int(o): int(o) + len(a)] = a
        got = engine.xxh3_batch(buf, offs, lens)
        for name, h in zip(names, got):
            if int(h) != hashes[name]:
                raise IOError(f'backed file "{name}" content hash differs')
        names, arrs, held = [], [], 0

    for rel in hashes:
        try:
            with open_file(rel) as f:
                data = f.read()
        except OSError as ex:
            raise IOError(f'open backed file "{rel}" for verification: {ex}') from ex
        names.append(rel)
        arrs.append(np.frombuffer(data, dtype=np.uint8))
        held += len(data)
        if held >= batch_bytes:
            flush()
    flush()
