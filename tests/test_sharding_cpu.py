"""CPU tests of the multi-GPU host logic: shard ranges, and a world_size-2 `gloo` run of the broadcast +
per-rank top-k + all-gather + merge path (the GPU kernels are replaced by a numpy brute force here; the
same sharding.py functions are what bench.py runs over NCCL)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pb_starphase_b200 import sharding  # noqa: E402


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 64, 12451):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert len({s[2] for s in spans}) == 1 and spans[0][2] * world >= n


def test_triangle_rows_partition_and_balance():
    for n in (1, 5, 100, 5695):
        for world in (1, 2, 4, 8):
            rows = [sharding.triangle_rows(n, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            if n >= 1000:
                areas = [sum(n - i for i in range(lo, hi)) for lo, hi in rows]
                assert max(areas) / (sum(areas) / world) < 1.02


def brute_topk(D, k, lo, hi, D2=None):
    recs = []
    A = D.shape[1]
    for i in range(lo, hi):
        for j in range(i, A):
            s = int(np.minimum(D[:, i], D[:, j]).sum())
            s2 = int(np.minimum(D2[:, i], D2[:, j]).sum()) if D2 is not None else 0
            le = (D[:, i] < D[:, j]) | ((D[:, i] == D[:, j]) & ((D2[:, i] <= D2[:, j]) if D2 is not None else True))
            recs.append((s, s2, i, j, int(le.sum())))
    recs.sort(key=lambda r: r[:4])
    return recs[:k]


def test_merge_topk_equals_global():
    rng = np.random.default_rng(0)
    D = rng.integers(0, 6, size=(9, 23)).astype(np.int32)  # many ties on purpose
    D2 = rng.integers(0, 4, size=(9, 23)).astype(np.int32)
    whole = brute_topk(D, 10, 0, 23, D2)
    for world in (2, 3, 8):
        parts = [brute_topk(D, 10, *sharding.triangle_rows(23, r, world), D2) for r in range(world)]
        assert sharding.merge_topk(parts, 10) == whole


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)  # rank 0's matrix is THE input; the others get it by broadcast
        D = rng.integers(0, 9, size=(12, 41)).astype(np.int32)
        D = sharding.broadcast_bytes(D.view(np.uint8).reshape(-1), 0).view(np.int32).reshape(12, 41)
        lo, hi = sharding.triangle_rows(41, rank, world)
        mine = brute_topk(D, 8, lo, hi)
        merged = sharding.all_gather_topk(mine, 8)
        np.save(os.path.join(out_dir, f"merged_{rank}.npy"), np.array(merged, dtype=np.int64))
        if rank == 0:
            np.save(os.path.join(out_dir, "D.npy"), D)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_broadcast_gather_merge(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    D = np.load(tmp_path / "D.npy")
    want = np.array(brute_topk(D, 8, 0, 41), dtype=np.int64)
    for r in range(2):
        got = np.load(tmp_path / f"merged_{r}.npy")
        assert (got == want).all()
