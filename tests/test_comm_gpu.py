"""GPU suite, world size 2: the multi-GPU C ABI (sp_comm_*: NCCL inside libstarphase_gpu.so) against the single-GPU path.
Needs two GPUs (gpurun --gpus 2); on a one-GPU box only the world-1 degenerate communicator runs."""
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _workload():
    from pb_starphase_b200 import synth

    alleles, reads, src, cdna = synth.hla_gene(5, "HLA-A", n_alleles=301, n_reads=70, with_cdna=True)
    ctargets = [cdna[int(s)] for s in src]
    return alleles, cdna, reads, ctargets


def _rank_main(rank, world, uid, q):
    try:
        import pb_starphase_b200 as sp
        from pb_starphase_b200 import binding

        alleles, cdna, reads, ctargets = _workload()
        ctx = sp.Context(rank)
        comm = sp.Comm(ctx, uid, rank, world)
        T = comm.bcast_targets(reads if rank == 0 else None, 0)
        Tc = comm.bcast_targets(ctargets if rank == 0 else None, 0)
        idx_d = binding.shard_plan([len(a) for a in alleles], world, rank)
        idx_c = binding.shard_plan([len(a) for a in cdna], world, rank)
        Pd = ctx.patterns([alleles[i] for i in idx_d])
        Pc = ctx.patterns([cdna[i] for i in idx_c])
        for rep in range(2):  # second round re-uses the cached shard layout
            Dd = comm.score_allgather(T, Pd, idx_d, len(alleles), 16)
            Dc = comm.score_allgather(Tc, Pc, idx_c, len(cdna), 16)
            top = comm.pair_minsum_topk(Dc, 12, d2=Dd)
            hd, hc = Dd.to_host(), Dc.to_host()
        comm.barrier()
        q.put((rank, hd, hc, top, ctx.launch_count()))
        comm.close()
        ctx.close()
    except Exception as e:  # surface the failure in the parent
        import traceback

        q.put((rank, "error", f"{type(e).__name__}: {e}\n{traceback.format_exc()}", None, 0))


def _single_gpu_answer():
    import pb_starphase_b200 as sp

    alleles, cdna, reads, ctargets = _workload()
    with sp.Context(0) as ctx:
        T, Tc, Pd, Pc = ctx.targets(reads), ctx.targets(ctargets), ctx.patterns(alleles), ctx.patterns(cdna)
        Dd, Dc = ctx.score_device(T, Pd, 16), ctx.score_device(Tc, Pc, 16)
        return Dd.to_host(), Dc.to_host(), ctx.pair_minsum_topk(Dc, 12, d2=Dd)


def _run_world(world):
    from pb_starphase_b200 import binding

    uid = binding.comm_unique_id() if world > 1 else bytes(128)
    mpx = mp.get_context("spawn")
    q = mpx.Queue()
    procs = [mpx.Process(target=_rank_main, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return sorted(res, key=lambda r: r[0])


def test_world1_comm_equals_plain_calls(oracle):
    want_d, want_c, want_top = _single_gpu_answer()
    (rank, hd, hc, top, launches), = _run_world(1)
    assert not isinstance(hd, str), hc
    assert (hd == want_d).all() and (hc == want_c).all() and top == want_top


def test_world2_nccl_equals_single_gpu(oracle):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    want_d, want_c, want_top = _single_gpu_answer()
    alleles, cdna, reads, ctargets = _workload()
    assert (want_d == oracle.score_batch(reads, alleles)).all()
    res = _run_world(2)
    for rank, hd, hc, top, launches in res:
        assert not isinstance(hd, str), hc
        assert (hd == want_d).all(), f"rank {rank}: gathered DNA matrix differs from the single-GPU matrix"
        assert (hc == want_c).all(), f"rank {rank}: gathered cDNA matrix differs"
        assert top == want_top, f"rank {rank}: merged top-k differs"
        assert launches > 0
