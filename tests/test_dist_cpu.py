"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: file sharding and the ONE exchange
step (all-gather of per-rank digest lists into a globally ordered list).  The known-set merge
itself runs on the GPU in the product; here the oracle's set is the checker for the ordering."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pbs_plus_b200 import dist as pdist  # noqa: E402


def test_shard_files_is_contiguous_and_balanced():
    for n, w in ((1024, 8), (10, 4), (3, 8), (0, 2), (1025, 8)):
        parts = [pdist.shard_files(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, _) in zip(parts, parts[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_shard_by_size_balances_ragged_files():
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 10_000, size=200)
    parts = pdist.shard_by_size(lens, 8)
    assert sorted(i for p in parts for i in p) == list(range(200))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(lens.max())


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        # every rank "produced" a different number of digests; some digests repeat across ranks
        rng = np.random.default_rng(7)
        pool = rng.integers(0, 256, size=(50, 32), dtype=np.uint8)
        counts_true = [13, 0, 21, 7][:world]
        idx = [rng.integers(0, 50, size=c) for c in counts_true]
        local = pool[idx[rank]]
        allg, counts = pdist.allgather_digests(local)
        assert counts == counts_true
        expect = np.concatenate([pool[i] for i in idx]) if sum(counts_true) else np.zeros((0, 32), np.uint8)
        assert (allg.numpy() == expect).all()
        # global-order KNOWN flags: identical on every rank, own slice extracted by offset
        flags_all = oracle.DigestSet().probe(expect, insert=True)

        class FakeSet:      # stands in for the GPU set: same insert() contract
            def __init__(self): self.s = oracle.DigestSet()
            def insert(self, d): return self.s.probe(d, insert=True)
        mine = pdist.global_known_flags(FakeSet(), allg, counts, rank)
        start = sum(counts[:rank])
        assert (mine == flags_all[start:start + counts[rank]]).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_allgather_digests_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
