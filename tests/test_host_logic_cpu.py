"""CPU-only tests of the host-side mirror of the reference surface (pbs_plus_b200/transfer.py,
buzhash.py): queueing, io.ReadFull error behaviour, index order, upload-only-new.  The GPU engine is
replaced by a stand-in that answers with the ORACLE (tests may use it as the checker); the real
engine runs the same flows in tests/test_gpu_parity.py."""
import hashlib
import io

import numpy as np
import pytest

import oracle
from pbs_plus_b200 import buzhash, transfer
from pbs_plus_b200._lib import CHUNK_KNOWN


class OracleEngine:
    """Same call contract as Engine.chunk_digest_streams / digest_set, answered by the oracle."""

    def __init__(self):
        self.calls = 0

    def chunk_digest_streams(self, cfg, streams, digest_set=None):
        self.calls += 1
        rec = oracle.chunk_digest_streams(oracle.config(cfg.avg), list(streams))
        if digest_set is not None:
            rec["flags"] = digest_set.insert(rec["digest"]) * CHUNK_KNOWN
        return rec

    def chunk_digest_batch_xxh3(self, cfg, buf, off, length, digest_set=None):
        streams = [np.asarray(buf[int(o): int(o + l)]) for o, l in zip(off, length)]
        rec = self.chunk_digest_streams(cfg, streams, digest_set)
        return rec, np.array([oracle.xxh3_64(s) for s in streams], dtype=np.uint64)

    def xxh3_batch(self, buf, off, length):
        return np.array([oracle.xxh3_64(np.asarray(buf[int(o): int(o + l)])) for o, l in zip(off, length)], dtype=np.uint64)

    def blob_encode_batch_z(self, buf, off, length):
        from oracle import pyref
        import zlib
        blobs = [pyref.blob_encode(np.asarray(buf[int(o): int(o + l)]).tobytes()) for o, l in zip(off, length)]
        return blobs, np.array([zlib.crc32(b[12:]) for b in blobs], dtype=np.uint32)

    def digest_set(self, hint=0):
        class S:
            def __init__(s): s.s = oracle.DigestSet()
            def insert(s, d): return s.s.probe(np.asarray(d, dtype=np.uint8).reshape(-1, 32), insert=True)
            def seed_didx(s, img):
                a = np.frombuffer(img, dtype=np.uint8)[4096:].reshape(-1, 40)[:, 8:]
                s.insert(np.ascontiguousarray(a)); return len(a)
        return S()


def rnd(n, seed):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def test_newconfig_matches_reference_call():
    c = buzhash.NewConfig(4096)                         # commit.go:303
    assert (c.avg, c.min, c.max) == (4 << 20, 1 << 20, 16 << 20)
    assert buzhash.NewConfigBytes(4096).avg == 4096
    for bad in (0, -1, 1 << 20):
        with pytest.raises(ValueError):
            buzhash.NewConfig(bad)


def test_dedup_writer_flow_index_order_and_upload_only_new():
    eng = OracleEngine()
    uploaded = []
    w = transfer.NewRemoteDedupSplitArchiveWriter(eng, buzhash.NewConfigBytes(1024), known=eng.digest_set(),
                                                  upload=lambda d, b: uploaded.append((d, b)))
    files = [("a", rnd(50_000, 1)), ("b", rnd(10, 2)), ("empty", rnd(0, 3)), ("a2", rnd(50_000, 1))]
    for name, data in files:
        w.WriteEntryReader(transfer.Entry(name, len(data)), io.BytesIO(data.tobytes()), len(data))
    idx = w.Finish()
    assert eng.calls == 1                               # one batched GPU call for the whole walk
    ref = oracle.chunk_digest_streams(oracle.config(1024), [d for _, d in files])
    assert [(r.end_off, r.digest) for r in idx] == [(int(r["end_off"]), bytes(r["digest"])) for r in ref]
    assert [r.path for r in idx] == [files[int(r["stream"])][0] for r in ref]
    assert all(r.known for r in idx if r.path == "a2") and not any(r.known for r in idx if r.path == "a")
    assert all(hashlib.sha256(b).digest() == d for d, b in uploaded)
    assert {d for d, _ in uploaded} == {r.digest for r in idx}      # every distinct chunk uploaded exactly once
    assert len(uploaded) == len({r.digest for r in idx})
    assert w.backed_hashes == {name: oracle.xxh3_64(data) for name, data in files}   # ow.backedHashes, commit.go:725
    with pytest.raises(RuntimeError):
        w.WriteEntry(transfer.Entry("late", 1), b"x")


def test_dedup_writer_renders_datablobs_for_new_chunks_only():
    """upload_blob: the POST /dynamic_chunk bodies of the chunks the server lacks, one batched render per flush; zero runs
    come out as compressed blobs."""
    from oracle import pyref
    eng = OracleEngine()
    blobs = []
    w = transfer.NewRemoteDedupSplitArchiveWriter(eng, buzhash.NewConfigBytes(65536), known=eng.digest_set(),
                                                  upload_blob=lambda d, b: blobs.append((d, b)))
    sparse = rnd(3_000_000, 4)
    sparse[500_000:2_200_000] = 0
    files = [("disk.img", sparse), ("copy.img", sparse.copy()), ("small", rnd(100, 5))]
    for name, data in files:
        w.WriteEntry(transfer.Entry(name, len(data)), data.tobytes())
    idx = w.Finish()
    new = [r for r in idx if not r.known]
    assert [d for d, _ in blobs] == [r.digest for r in new]            # in index order, new chunks only
    assert all(r.known for r in idx if r.path == "copy.img")
    ends = {}
    n_comp = 0
    for r, (_, blob) in zip(new, blobs):
        data = dict(files)[r.path]
        s0 = ends.get(r.path, 0)
        for q in idx:                                                    # start = previous end of the same file
            if q.path == r.path and q.end_off < r.end_off:
                s0 = max(s0, q.end_off)
        chunk = data[s0: r.end_off].tobytes()
        assert hashlib.sha256(chunk).digest() == r.digest
        assert blob == pyref.blob_encode(chunk)
        n_comp += blob[:8] == pyref.BLOB_MAGIC_COMPRESSED
    assert n_comp >= 1 and sum(len(b) for _, b in blobs) < len(sparse) - 1_000_000


def test_short_reader_is_an_error_like_io_readfull():
    w = transfer.DedupWriter(OracleEngine(), buzhash.NewConfigBytes(1024))
    with pytest.raises(IOError, match="unexpected EOF"):
        w.WriteEntryReader(transfer.Entry("short", 100), io.BytesIO(b"abc"), 100)


def test_batching_threshold_flushes_and_previous_index_seeds_the_known_set():
    eng = OracleEngine()
    data = rnd(40_000, 7)
    first = oracle.chunk_digest(oracle.config(1024), data)
    didx = bytearray(4096)
    for r in first:
        didx += int(r["end_off"]).to_bytes(8, "little") + bytes(r["digest"])
    w = transfer.NewRemoteDedupSplitArchiveWriter(eng, buzhash.NewConfigBytes(1024), orig_payload_idx=bytes(didx))
    w.batch_bytes = 30_000
    w.WriteEntry(transfer.Entry("same", len(data)), data.tobytes())          # crosses the threshold -> flush
    assert eng.calls == 1
    w.WriteEntry(transfer.Entry("new", 5000), rnd(5000, 8).tobytes())
    idx = w.Finish()
    assert eng.calls == 2
    assert all(r.known for r in idx if r.path == "same") and not any(r.known for r in idx if r.path == "new")


def test_verify_backed_file_hashes_mirrors_the_reference_errors(tmp_path):
    """verifyBackedFileHashes (commit.go:957-976): unchanged files pass, a changed file and a missing file fail with
    the reference's messages; several files share one batched hash call."""
    eng = OracleEngine()
    files = {"a/x.bin": rnd(70_000, 11), "b.bin": rnd(10, 12), "empty": rnd(0, 13), "big": rnd(300_000, 14)}
    for k, v in files.items():
        (tmp_path / k).parent.mkdir(parents=True, exist_ok=True)
        (tmp_path / k).write_bytes(v.tobytes())
    hashes = {k: oracle.xxh3_64(v) for k, v in files.items()}
    opener = lambda rel: open(tmp_path / rel, "rb")
    transfer.verifyBackedFileHashes(eng, opener, hashes)                       # all unchanged
    transfer.verifyBackedFileHashes(eng, opener, hashes, batch_bytes=50_000)   # split over several batches
    (tmp_path / "b.bin").write_bytes(b"changed!!!")
    with pytest.raises(IOError, match='backed file "b.bin" content hash differs'):
        transfer.verifyBackedFileHashes(eng, opener, hashes)
    (tmp_path / "b.bin").write_bytes(files["b.bin"].tobytes())
    (tmp_path / "big").unlink()
    with pytest.raises(IOError, match='open backed file "big" for verification'):
        transfer.verifyBackedFileHashes(eng, opener, hashes)


# ---- the layout-faithful writer (host logic only; the real engine runs the same flows in tests/test_gpu_round2.py) ----
class OracleStream:
    """Stand-in for engine.Stream: buffers the bytes and the suggested boundaries, answers finish() with the oracle."""

    def __init__(self, cfg, known):
        self.cfg, self.known, self.buf, self.sug, self.done = cfg, known, bytearray(), [], False
        self._slot = np.zeros(4096, dtype=np.uint8)

    position = property(lambda self: len(self.buf))

    def write(self, data): self.buf += bytes(np.asarray(data, dtype=np.uint8))
    def suggest(self, off):
        assert off >= len(self.buf) and (not self.sug or off > self.sug[-1])
        self.sug.append(off)
    def reserve(self): return self._slot
    def commit(self, n): self.buf += self._slot[:n].tobytes()
    def poll(self, cap=0): return np.zeros(0, dtype=oracle.CHUNK_DTYPE)
    def finish(self):
        data = np.frombuffer(bytes(self.buf), dtype=np.uint8)
        rec = oracle.chunk_digest_forced(oracle.config(self.cfg.avg), data, np.array([s for s in self.sug if 0 < s < len(data)], np.uint64))
        if self.known is not None:
            rec["flags"] = self.known.insert(rec["digest"]) * CHUNK_KNOWN
        return rec
    def close(self): self.done = True


class OracleStreamEngine(OracleEngine):
    def stream(self, cfg, known=None):
        return OracleStream(cfg, known)


def _payload_stream(files):
    parts, pos, starts = [transfer.PXAR_PAYLOAD_START_MARKER.to_bytes(8, "little") + (16).to_bytes(8, "little")], 16, []
    for f in files:
        starts.append(pos); parts += [transfer.payload_header(len(f)), f.tobytes()]; pos += 16 + len(f)
    parts.append(transfer.PXAR_PAYLOAD_TAIL_MARKER.to_bytes(8, "little") + (16).to_bytes(8, "little"))
    return np.frombuffer(b"".join(parts), dtype=np.uint8), starts


def test_payload_header_layout():
    h = transfer.payload_header(1000)
    assert len(h) == 16 and int.from_bytes(h[:8], "little") == transfer.PXAR_PAYLOAD and int.from_bytes(h[8:], "little") == 1016


def test_payload_stream_writer_frames_entries_and_returns_payload_offsets():
    eng = OracleStreamEngine()
    files = [rnd(n, 20 + i) for i, n in enumerate([0, 5, 30_000, 4096, 77_777])]
    w = transfer.PayloadStreamWriter(eng, buzhash.NewConfigBytes(1024), eng.digest_set())
    stream, starts = _payload_stream(files)
    for i, f in enumerate(files):
        assert w.WriteEntryReader(transfer.Entry(f"f{i}", len(f)), io.BytesIO(f.tobytes()), len(f)) == starts[i]
    idx = w.Finish()
    ref = oracle.chunk_digest_forced(oracle.config(1024), stream, np.array(starts, np.uint64))
    assert [(e, d) for e, d, _ in idx] == [(int(r["end_off"]), bytes(r["digest"])) for r in ref]
    assert idx[-1][0] == len(stream) and w.payload_offsets == {f"f{i}": s for i, s in enumerate(starts)}
    with pytest.raises(RuntimeError):
        w.WriteEntryReader(transfer.Entry("late", 1), io.BytesIO(b"x"), 1)
    with pytest.raises(IOError, match="unexpected EOF"):
        transfer.PayloadStreamWriter(eng, buzhash.NewConfigBytes(1024)).WriteEntryReader(transfer.Entry("s", 9), io.BytesIO(b"abc"), 9)


def test_write_entry_ref_splices_the_previous_index_and_rejects_descending_offsets():
    """commit.go:752 / :848-860: chunk reuse by reference; the error text carries "not strictly greater"."""
    eng = OracleStreamEngine()
    cfg = buzhash.NewConfigBytes(1024)
    files = [rnd(n, 60 + i) for i, n in enumerate([40_000, 30_000, 50_000])]
    w0 = transfer.PayloadStreamWriter(eng, cfg)
    offs = [w0.WriteEntryReader(transfer.Entry(f"f{i}", len(f)), io.BytesIO(f.tobytes()), len(f)) for i, f in enumerate(files)]
    prev = w0.Finish()
    ends = np.array([e for e, _, _ in prev], dtype=np.uint64)
    digs = np.array([np.frombuffer(d, dtype=np.uint8) for _, d, _ in prev])
    w1 = transfer.PayloadStreamWriter(eng, cfg, prev_index=(ends, digs))
    head = rnd(7000, 99)
    w1.WriteEntryReader(transfer.Entry("new", len(head)), io.BytesIO(head.tobytes()), len(head))
    pos_before = w1.position
    o1 = w1.WriteEntryRef(transfer.Entry("f1", len(files[1])), offs[1])
    lo = int(np.searchsorted(ends, offs[1], side="right"))
    hi = int(np.searchsorted(ends, offs[1] + 16 + len(files[1]), side="left"))
    first_start = int(ends[lo - 1]) if lo else 0
    assert o1 == pos_before + (offs[1] - first_start)                    # same distance from the first injected chunk's start
    assert w1.position == pos_before + int(ends[hi]) - first_start       # the injected chunks' bytes are part of the new stream
    with pytest.raises(transfer.NotStrictlyGreater, match="not strictly greater"):
        w1.WriteEntryRef(transfer.Entry("f0", len(files[0])), offs[0])
    with pytest.raises(IOError, match="outside the previous index"):
        w1.WriteEntryRef(transfer.Entry("f2", 10**9), offs[2])
    idx = w1.Finish()
    injected = [(d, k) for e, d, k in idx if pos_before < e <= w1.position - 16]
    assert [d for d, _ in injected][: hi - lo + 1] == [bytes(x) for x in digs[lo: hi + 1]] and all(k for _, k in injected[: hi - lo + 1])
    e = [x[0] for x in idx]
    assert e == sorted(e) and len(set(e)) == len(e)
    with pytest.raises(IOError, match="without a previous payload index"):
        transfer.PayloadStreamWriter(eng, cfg).WriteEntryRef(transfer.Entry("x", 1), 5)
