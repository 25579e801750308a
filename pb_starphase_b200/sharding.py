"""Host-side partitioning of the scoring path over the GPUs of one box and the merge of the
per-rank top-k pair lists.  Pure Python/torch.distributed plumbing: works with the `nccl` backend on GPUs and
with `gloo` on CPU (tests/test_sharding_cpu.py runs it at world_size 2).

K1 shards the allele (pattern) index range, K2 shards row ranges of the upper-triangular pair space; the only
data-path exchanges are a broadcast of the read set, an all-gather of the distance-matrix shards and an all-gather
of k <= 64 fixed-size records per rank.  Because every reduction is an integer sum/min and the order key
(score, score2, i, j) is total, the merged answer is identical for any world size.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

Record = Tuple[int, ...]  # (score, score2, i, j, c1)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous allele shard [lo, hi) of rank `rank`, plus the padded shard size (equal on every rank, so
    the shards can live in one all-gather buffer)."""
    size = (n + world - 1) // world
    return min(rank * size, n), min((rank + 1) * size, n), size


def triangle_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Row range [lo, hi) of the pairs i <= j < n with (nearly) equal pair count per rank."""
    def edge(b: int) -> int:
        return int(round(n * (1.0 - (1.0 - b / world) ** 0.5)))
    return edge(rank), (n if rank == world - 1 else edge(rank + 1))


def merge_topk(lists: Sequence[Sequence[Record]], k: int) -> List[Record]:
    """Merge per-shard record lists by the library's order: (score, score2, i, j) ascending."""
    rows = [tuple(int(x) for x in r) for lst in lists for r in lst]
    rows.sort(key=lambda r: (r[0], r[1], r[2], r[3]))
    return rows[:k]


def all_gather_topk(recs: Sequence[Record], k: int, device=None, group=None) -> List[Record]:
    """All-gather this rank's (<= k) records and merge; every rank returns the same list."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return merge_topk([recs], k)
    world = dist.get_world_size(group)
    buf = torch.full((k, 5), -1, dtype=torch.int64, device=device)
    if len(recs):
        buf[: len(recs)] = torch.tensor([list(r) for r in recs], dtype=torch.int64, device=device)
    allb = torch.empty((world * k, 5), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allb, buf, group=group)
    rows = [tuple(r) for r in allb.cpu().tolist() if r[2] >= 0]
    return merge_topk([rows], k)


def broadcast_bytes(arr: np.ndarray, src: int = 0, device=None, group=None) -> np.ndarray:
    """Broadcast a uint8/int64 numpy array (the packed read set) from rank `src`; returns the received copy."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device) if device is not None else torch.from_numpy(np.ascontiguousarray(arr).copy())
    dist.broadcast(t, src, group=group)
    return t.cpu().numpy()
