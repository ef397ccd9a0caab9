"""oracle -- ctypes front-end of the CPU restatement (oracle/oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT.  Importable only from tests/, from
``__graft_entry__.smoke()`` and from the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  The product package ``pbs_plus_b200`` never imports it.

PARITY UNPINNED for chunk boundaries (see the header of oracle.c): the reference
(pbs-plus) delegates the arithmetic to the absent Go module
github.com/pbs-plus/pxar v0.19.2 (reference go.mod:28, call sites
internal/pxarmount/commit.go:302-305,:329,:720) and holds no golden vectors.
SHA-256 *is* pinned (NIST vectors + hashlib).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    """Compile liboracle.so with gcc (building the checker is not using it)."""
    src = _HERE / "oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(
        src.stat().st_mtime, (_HERE / "buzhash_table.h").stat().st_mtime
    ):
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Cfg(C.Structure):
    _fields_ = [
        ("avg", C.c_uint32), ("min", C.c_uint32), ("max", C.c_uint32), ("mask", C.c_uint32),
        ("break_min", C.c_uint32), ("window", C.c_uint32), ("table", C.c_uint32 * 256),
    ]


class Chunk(C.Structure):
    _fields_ = [("stream", C.c_uint32), ("flags", C.c_uint32), ("end_off", C.c_uint64), ("digest", C.c_uint8 * 32)]


CHUNK_DTYPE = np.dtype([("stream", "<u4"), ("flags", "<u4"), ("end_off", "<u8"), ("digest", "u1", (32,))])
assert CHUNK_DTYPE.itemsize == C.sizeof(Chunk) == 48


class Corpus(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("file_len", C.c_uint64), ("block_len", C.c_uint64),
        ("run_blocks", C.c_uint32), ("dup_permille", C.c_uint32),
        ("edit_mode", C.c_uint32), ("edit_thresh16", C.c_uint32), ("edit_seed", C.c_uint64),
    ]


class Chunker(C.Structure):
    _fields_ = [("h", C.c_uint32), ("window_size", C.c_uint32), ("chunk_size", C.c_uint64), ("window", C.c_uint8 * 64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_config.argtypes = [C.c_uint32, u32p, C.POINTER(Cfg)]
        L.orc_config_kib.argtypes = [C.c_uint32, u32p, C.POINTER(Cfg)]
        L.orc_default_table.restype = u32p
        L.orc_chunker_reset.argtypes = [C.POINTER(Chunker)]
        L.orc_chunker_scan.argtypes = [C.POINTER(Cfg), C.POINTER(Chunker), C.c_void_p, C.c_uint64]
        L.orc_chunker_scan.restype = C.c_uint64
        L.orc_chunk_buffer.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_chunk_buffer.restype = C.c_uint64
        L.orc_chunk_buffer_closed_form.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_chunk_buffer_closed_form.restype = C.c_uint64
        L.orc_window_hash.argtypes = [u32p, C.c_void_p]
        L.orc_window_hash.restype = C.c_uint32
        L.orc_sha256.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_force_portable_sha.argtypes = [C.c_int]
        L.orc_chunk_digest.argtypes = [C.POINTER(Cfg), C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_chunk_digest.restype = C.c_uint64
        L.orc_set_create.argtypes = [C.c_uint64]
        L.orc_set_create.restype = C.c_void_p
        L.orc_set_destroy.argtypes = [C.c_void_p]
        L.orc_set_count.argtypes = [C.c_void_p]
        L.orc_set_count.restype = C.c_uint64
        L.orc_set_probe_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.orc_corpus_fill.argtypes = [C.POINTER(Corpus), C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_chunk_digest_mt.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_xxh3_64.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_xxh3_64.restype = C.c_uint64
        L.orc_corpus_fill_mt.argtypes = [C.POINTER(Corpus), C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_chunk_digest_forced.argtypes = [C.POINTER(Cfg), C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                              C.c_void_p, C.c_uint64]
        L.orc_chunk_digest_forced.restype = C.c_uint64
        L.orc_corpus_chunk_digest_mt.argtypes = [C.POINTER(Cfg), C.POINTER(Corpus), C.c_uint64, C.c_uint32, C.c_uint32,
                                                 C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def default_table() -> np.ndarray:
    return np.ctypeslib.as_array(lib().orc_default_table(), shape=(256,)).copy()


def config(avg_bytes: int, table: np.ndarray | None = None) -> Cfg:
    """Restates buzhash.NewConfig (reference commit.go:302-305); avg in BYTES."""
    cfg = Cfg()
    tp = None
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.shape == (256,)
        tp = table.ctypes.data_as(C.POINTER(C.c_uint32))
    rc = lib().orc_config(avg_bytes, tp, C.byref(cfg))
    if rc != 0:
        raise ValueError(f"orc_config({avg_bytes}) -> {rc}")
    return cfg


def _buf(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a


def chunk_ends(cfg: Cfg, data, feed: int = 0) -> np.ndarray:
    """Chunk END offsets of one stream via the streaming scan(), fed `feed` bytes at a time."""
    a = _buf(data)
    cap = len(a) // cfg.min + 2
    ends = np.empty(cap, dtype=np.uint64)
    n = lib().orc_chunk_buffer(C.byref(cfg), a.ctypes.data, len(a), feed, ends.ctypes.data, cap)
    assert n <= cap
    return ends[:n].copy()


def chunk_ends_closed_form(cfg: Cfg, data) -> np.ndarray:
    a = _buf(data)
    cap = len(a) // cfg.min + 2
    ends = np.empty(cap, dtype=np.uint64)
    n = lib().orc_chunk_buffer_closed_form(C.byref(cfg), a.ctypes.data, len(a), ends.ctypes.data, cap)
    return ends[:n].copy()


def window_hash(table: np.ndarray, win64) -> int:
    t = np.ascontiguousarray(table, dtype=np.uint32)
    w = _buf(win64)
    assert len(w) == 64
    return int(lib().orc_window_hash(t.ctypes.data_as(C.POINTER(C.c_uint32)), w.ctypes.data))


def sha256(data) -> bytes:
    a = _buf(data)
    out = np.empty(32, dtype=np.uint8)
    lib().orc_sha256(a.ctypes.data if len(a) else None, len(a), out.ctypes.data)
    return out.tobytes()


def xxh3_64(data) -> int:
    """XXH3-64, seed 0 (what xxh3.New()...Sum64() returns at commit.go:717-725)."""
    b = _buf(data)
    return int(lib().orc_xxh3_64(b.ctypes.data, b.size))


def force_portable_sha(on: bool) -> None:
    lib().orc_force_portable_sha(1 if on else 0)


def chunk_digest(cfg: Cfg, data, stream: int = 0) -> np.ndarray:
    """stream -> chunks -> digests for one stream; returns CHUNK_DTYPE records."""
    a = _buf(data)
    cap = len(a) // cfg.min + 2
    out = np.zeros(cap, dtype=CHUNK_DTYPE)
    n = lib().orc_chunk_digest(C.byref(cfg), stream, a.ctypes.data, len(a), out.ctypes.data, cap)
    return out[:n].copy()


def chunk_digest_streams(cfg: Cfg, streams, threads: int = 1) -> np.ndarray:
    """Many streams (list of uint8 arrays) on `threads` host threads; records in stream order."""
    arrs = [_buf(s) for s in streams]
    n = len(arrs)
    if n == 0:
        return np.zeros(0, dtype=CHUNK_DTYPE)
    cap = max(len(a) for a in arrs) // cfg.min + 2
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = np.array([len(a) for a in arrs], dtype=np.uint64)
    out = np.zeros(n * cap, dtype=CHUNK_DTYPE)
    n_out = np.zeros(n, dtype=np.uint64)
    rc = lib().orc_chunk_digest_mt(C.byref(cfg), ptrs, lens.ctypes.data, n, threads, out.ctypes.data, cap,
                                   n_out.ctypes.data)
    assert rc == 0
    return np.concatenate([out[i * cap: i * cap + int(n_out[i])] for i in range(n)])


def chunk_digest_forced(cfg: Cfg, data, forced, stream: int = 0) -> np.ndarray:
    """One stream with suggested boundaries (strictly increasing offsets), see orc_chunk_digest_forced."""
    a = _buf(data)
    f = np.ascontiguousarray(forced, dtype=np.uint64)
    cap = len(a) // cfg.min + len(f) + 2
    out = np.zeros(cap, dtype=CHUNK_DTYPE)
    n = lib().orc_chunk_digest_forced(C.byref(cfg), stream, a.ctypes.data, len(a), f.ctypes.data, len(f),
                                      out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def corpus_chunk_digest(cfg: Cfg, c: Corpus, first_file: int, n_files: int, threads: int | None = None) -> np.ndarray:
    """Chunk records of whole corpus files, generated on the fly per worker (no 64 GiB host copy)."""
    threads = threads or os.cpu_count() or 1
    cap = c.file_len // cfg.min + 2
    out = np.zeros(n_files * cap, dtype=CHUNK_DTYPE)
    n_out = np.zeros(n_files, dtype=np.uint64)
    rc = lib().orc_corpus_chunk_digest_mt(C.byref(cfg), C.byref(c), first_file, n_files, threads, out.ctypes.data, cap,
                                          n_out.ctypes.data)
    assert rc == 0, rc
    return np.concatenate([out[i * cap: i * cap + int(n_out[i])] for i in range(n_files)])


class DigestSet:
    """Known-digest set (reference: dedup seed commit.go:286-294,:324-329)."""

    def __init__(self, capacity: int = 1024):
        self._h = lib().orc_set_create(capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_set_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(lib().orc_set_count(self._h))

    def probe(self, digests: np.ndarray, insert: bool) -> np.ndarray:
        d = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
        hit = np.zeros(len(d), dtype=np.uint8)
        lib().orc_set_probe_insert(self._h, d.ctypes.data, len(d), hit.ctypes.data, 1 if insert else 0)
        return hit


def corpus(seed: int, file_len: int, block_len: int = 4 << 20, run_blocks: int = 8, dup_permille: int = 0,
           edit_mode: int = 0, edit_thresh16: int = 655, edit_seed: int = 5) -> Corpus:
    assert block_len % 8 == 0 and block_len > 0 and run_blocks >= 1
    return Corpus(seed, file_len, block_len, run_blocks, dup_permille, edit_mode, edit_thresh16, edit_seed)


def corpus_file(c: Corpus, file_id: int, off: int = 0, n: int | None = None) -> np.ndarray:
    if n is None:
        n = c.file_len - off
    out = np.empty(n, dtype=np.uint8)
    lib().orc_corpus_fill(C.byref(c), file_id, off, out.ctypes.data, n)
    return out


def corpus_files(c: Corpus, first_file: int, n_files: int, threads: int | None = None) -> list[np.ndarray]:
    threads = threads or os.cpu_count() or 1
    arrs = [np.empty(c.file_len, dtype=np.uint8) for _ in range(n_files)]
    ptrs = (C.c_void_p * n_files)(*[a.ctypes.data for a in arrs])
    rc = lib().orc_corpus_fill_mt(C.byref(c), first_file, ptrs, n_files, threads)
    assert rc == 0
    return arrs
