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
