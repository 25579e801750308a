"""GPU parity tests for K2 (pair min-sum + deterministic top-k), through the C ABI."""
import numpy as np
import pytest

import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,A", [(1, 1), (3, 2), (64, 64), (65, 63), (130, 200), (37, 257)])
def test_topk_host_matrix(ctx, oracle, R, A):
    rng = np.random.default_rng(R * 1000 + A)
    D = rng.integers(0, 40, size=(R, A)).astype(np.int32)  # small range => many ties => tie-break matters
    for k in (1, 10, 64):
        got = ctx.pair_minsum_topk(D, k)
        assert got == oracle.pair_minsum_topk(D, k)


def test_full_matrix(ctx, oracle):
    rng = np.random.default_rng(5)
    D = rng.integers(0, 70000, size=(90, 150)).astype(np.int32)  # values beyond u16: chain-pair path
    assert (ctx.pair_minsum_full(D) == oracle.pair_minsum_full(D)).all()


def test_device_pipeline_and_row_sharding(ctx, oracle):
    """K1 -> K2 on the device (u16 matrix), then the allele-pair row blocks used for multi-GPU
    sharding: merging per-shard top-k lists by (score, i, j) equals the unsharded list."""
    alleles, reads, _ = synth.hla_gene(3, "HLA-B", n_alleles=200, n_reads=40)
    P, T = ctx.patterns(alleles), ctx.targets(reads)
    dm = ctx.score_device(T, P, elem_bits=16)
    D = dm.to_host()
    want = oracle.pair_minsum_topk(D, 16)
    assert ctx.pair_minsum_topk(dm, 16) == want
    merged = []
    for lo, hi in ((0, 50), (50, 64), (64, 129), (129, 200)):
        merged += ctx.pair_minsum_topk(dm, 16, lo, hi)
    merged.sort(key=lambda r: (r[0], r[1], r[2]))
    assert merged[:16] == want


@pytest.mark.parametrize("R,A", [(5, 7), (70, 130), (33, 65)])
def test_topk_dual_matrix(ctx, oracle, R, A):
    """(primary, secondary) lexicographic key = (cDNA, DNA) order of HlaMappingScore
    (src/hla/mapping.rs:111-117) carried to pair sums; many primary ties force the secondary."""
    rng = np.random.default_rng(R * 31 + A)
    D1 = rng.integers(0, 4, size=(R, A)).astype(np.int32)
    D2 = rng.integers(0, 60, size=(R, A)).astype(np.int32)
    for k in (1, 16, 64):
        assert ctx.pair_minsum_topk(D1, k, d2=D2) == oracle.pair_minsum_topk(D1, k, D2=D2)


def test_dual_device_matrices(ctx, oracle):
    alleles, reads, _, cdna = synth.hla_gene(11, "HLA-A", n_alleles=130, n_reads=30, with_cdna=True)
    rng = np.random.default_rng(4)
    ctargets = [cdna[int(rng.integers(0, len(cdna)))] for _ in reads]
    T, Tc = ctx.targets(reads), ctx.targets(ctargets)
    P, Pc = ctx.patterns(alleles), ctx.patterns(cdna)
    dd, dc = ctx.score_device(T, P, elem_bits=16), ctx.score_device(Tc, Pc, elem_bits=16)
    got = ctx.pair_minsum_topk(dc, 12, d2=dd)
    assert got == oracle.pair_minsum_topk(dc.to_host(), 12, D2=dd.to_host())


def test_host_matrix_value_range_is_checked(ctx):
    """K2 adds 32 reads in 32 bits before widening: host matrices with values outside [0, 2^27) are refused, not wrapped around."""
    import pb_starphase_b200 as sp

    D = np.full((40, 6), 7, dtype=np.int32)
    assert ctx.pair_minsum_topk(D, 3)[0][0] == 40 * 7
    for bad in (1 << 27, -1):
        E = D.copy()
        E[17, 3] = bad
        with pytest.raises(sp.SpError):
            ctx.pair_minsum_topk(E, 3)
        with pytest.raises(sp.SpError):
            ctx.pair_minsum_full(E)
    D[:, :] = (1 << 27) - 1  # the largest admissible value, 40 reads: beyond 32 bits in the total, exact
    assert ctx.pair_minsum_topk(D, 1)[0][0] == 40 * ((1 << 27) - 1)
