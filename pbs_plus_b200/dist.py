"""Multi-GPU plumbing: one process per GPU, files sharded across ranks, ONE exchange step.

The path shards naturally (streams are independent): every rank runs
K1..K3 on its own files with no data-path collective.  The only exchange is the all-gather
of each step's chunk digests, after which every rank inserts the same globally ordered
list into its replicated digest set, so the KNOWN flags equal the single-GPU run.
torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_files(n_files: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, size-balanced shard of a uniform-file corpus: (first_file, count).
    Contiguous ranges keep rank order == global (file, chunk) order."""
    base, extra = divmod(n_files, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_by_size(lengths, world: int) -> list[list[int]]:
    """Greedy size-balanced assignment of whole files to ranks (largest first) for ragged corpora."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    loads = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        out[r].append(int(i))
        loads[r] += int(lengths[i])
    return [sorted(x) for x in out]


def allgather_digests(local, group=None, device=None):
    """All-gather variable-length digest lists.  `local`: (n,32) uint8 numpy array or torch tensor.
    Returns (all_digests tensor (total,32) in rank order on `device`, counts list)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if isinstance(local, np.ndarray):
        local = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8).reshape(-1, 32))
    if device is not None:
        local = local.to(device)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    if m == 0:
        return torch.zeros((0, 32), dtype=torch.uint8, device=local.device), counts
    padded = torch.zeros((m, 32), dtype=torch.uint8, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world, m, 32), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(gathered.view(-1), padded.view(-1), group=group)
    parts = [gathered[r, : counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0).contiguous(), counts


def global_known_flags(digest_set, all_digests, counts, rank: int) -> np.ndarray:
    """Insert the globally ordered digest list into this rank's replicated set (on the GPU) and
    return the KNOWN flags of this rank's own chunks."""
    if hasattr(all_digests, "is_cuda") and all_digests.is_cuda:
        import torch
        # the gather ran on torch's stream; the library's kernels run on its own streams
        torch.cuda.current_stream(all_digests.device).synchronize()
        n = int(all_digests.shape[0])
        hit = np.zeros(n, dtype=np.uint8)
        eng = digest_set._eng
        eng._ck(eng._L.pbsgpu_set_insert(digest_set._h, all_digests.data_ptr() if n else None, n,
                                         hit.ctypes.data if n else None))
    else:
        hit = digest_set.insert(all_digests.cpu().numpy() if hasattr(all_digests, "cpu") else all_digests)
    start = sum(counts[:rank])
    return hit[start: start + counts[rank]]
