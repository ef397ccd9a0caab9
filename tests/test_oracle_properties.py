"""Property-based checks (hypothesis) of the oracle's chunker: the invariants that must hold for ANY input
and that the GPU tests then inherit by comparing against this oracle."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import pyref

AVGS = st.sampled_from([256, 512, 1024, 4096])


def data_strategy():
    kinds = st.sampled_from(["random", "zeros", "lowentropy", "periodic"])
    return st.tuples(kinds, st.integers(0, 60_000), st.integers(0, 2**31))


def make(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "random":
        return rng.integers(0, 256, size=n, dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "lowentropy":
        return rng.integers(0, 3, size=n, dtype=np.uint8)
    return np.resize(rng.integers(0, 256, size=max(1, seed % 97 + 1), dtype=np.uint8), n)


@settings(max_examples=60, deadline=None)
@given(AVGS, data_strategy())
def test_cuts_cover_the_stream_and_respect_min_max(avg, d):
    data = make(*d)
    cfg = oracle.config(avg)
    ends = oracle.chunk_ends(cfg, data).tolist()
    if len(data) == 0:
        assert ends == []
        return
    assert ends[-1] == len(data) and ends == sorted(set(ends))
    lens = np.diff([0] + ends)
    assert (lens <= cfg.max).all() and (lens[:-1] >= max(cfg.min, 65)).all()
    assert ends == pyref.chunk_ends(oracle.default_table(), data, avg) == oracle.chunk_ends_closed_form(cfg, data).tolist()


@settings(max_examples=40, deadline=None)
@given(AVGS, data_strategy(), st.lists(st.integers(1, 9000), min_size=1, max_size=12))
def test_streaming_scan_is_split_invariant(avg, d, pieces):
    data = make(*d)
    cfg = oracle.config(avg)
    ref = oracle.chunk_ends(cfg, data).tolist()
    for feed in pieces:
        assert oracle.chunk_ends(cfg, data, feed=feed).tolist() == ref


@settings(max_examples=30, deadline=None)
@given(AVGS, data_strategy(), st.integers(0, 59_999), st.integers(0, 255))
def test_an_edit_only_moves_nearby_cuts(avg, d, pos, val):
    """Content-defined chunking resynchronises: a one-byte edit cannot change a cut that lies more than
    max + window bytes after it once a common cut has been reached, and never changes cuts before it."""
    data = make(*d)
    if len(data) == 0:
        return
    pos %= len(data)
    cfg = oracle.config(avg)
    a = oracle.chunk_ends(cfg, data).tolist()
    edited = data.copy(); edited[pos] = val
    b = oracle.chunk_ends(cfg, edited).tolist()
    before_a = [e for e in a if e <= pos]
    before_b = [e for e in b if e <= pos]
    assert before_a == before_b                                # cuts before the edit are untouched
    common = sorted(set(a) & set(b) & set(range(pos + 65, len(data) + 1)))
    if common:                                                 # after the first common cut everything agrees again
        c = common[0]
        assert [e for e in a if e >= c] == [e for e in b if e >= c]


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=0, max_size=5000), st.integers(min_value=0, max_value=7))
def test_xxh3_oracle_equals_libxxhash_on_arbitrary_bytes(data, lead):
    xxhash = pytest.importorskip("xxhash")
    buf = np.frombuffer(b"\x00" * lead + data, dtype=np.uint8)[lead:]
    assert oracle.xxh3_64(buf) == xxhash.xxh3_64_intdigest(data)
