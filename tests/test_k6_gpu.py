"""K6 sp_variant_match (row N3 of SURVEY.md 8f): the haplotype loop of Cyp2d6Extractor::assign_haplotype
(src/cyp2d6/haplotyper.rs:470-517) on the GPU, through the C ABI, against the loop restated in oracle/starphase_oracle.py; and
the C++ host's assign_haplotypes_from_alleles (arg-max, tie handling, RegionVariant list) against the oracle flow."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import starphase_oracle as so  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pb_starphase_b200 as sp

    with sp.Context(0) as c:
        yield c


def random_case(rng, n_seq, n_hap, n_var, p_alt=0.03):
    hap = (rng.random((n_hap, n_var)) < p_alt).astype(np.uint8)
    seq = np.zeros((n_seq, n_var), dtype=np.uint8)
    for s in range(n_seq):
        src = hap[rng.integers(0, n_hap)] if n_hap else np.zeros(n_var, np.uint8)
        seq[s] = src
        noise = rng.random(n_var)
        seq[s][noise < 0.02] ^= 1
        seq[s][(noise >= 0.02) & (noise < 0.04)] = 2
        seq[s][(noise >= 0.04) & (noise < 0.08)] = 3
    vi = (rng.random(n_var) < 0.3).astype(np.uint8)
    return seq, hap, vi


@pytest.mark.parametrize("n_seq,n_hap,n_var", [(1, 1, 1), (3, 5, 31), (3, 5, 32), (2, 7, 33), (4, 9, 64), (6, 40, 387), (1, 300, 1000),
                                               (30, 540, 387), (0, 4, 10), (4, 0, 10), (3, 3, 0)])
def test_variant_match_vs_oracle(ctx, n_seq, n_hap, n_var):
    rng = np.random.default_rng([n_seq, n_hap, n_var])
    seq, hap, vi = random_case(rng, n_seq, n_hap, n_var)
    vm, am = ctx.variant_match(seq, hap, vi)
    assert vm.shape == am.shape == (n_seq, n_hap)
    for s in range(n_seq):
        for h in range(n_hap):
            assert (int(vm[s, h]), int(am[s, h])) == so.variant_match(seq[s].tolist(), hap[h].tolist(), vi.tolist())


def test_variant_match_extremes(ctx):
    n_var = 387  # the size of the reference's CYP2D6 variant list (src/cyp2d6/haplotyper.rs:927)
    vi = np.ones(n_var, np.uint8)
    hap = np.stack([np.zeros(n_var, np.uint8), np.ones(n_var, np.uint8)])
    seq = np.stack([np.full(n_var, v, np.uint8) for v in (0, 1, 2, 3)])
    vm, am = ctx.variant_match(seq, hap, vi)
    assert am.tolist() == [[n_var, 0], [0, n_var], [n_var, n_var], [0, 0]] and (vm == am).all()
    vm, _ = ctx.variant_match(seq, hap, np.zeros(n_var, np.uint8))
    assert not vm.any()


def test_variant_match_rejects_bad_states(ctx):
    import pb_starphase_b200 as sp

    ok = np.zeros((1, 8), np.uint8)
    with pytest.raises(sp.SpError):
        ctx.variant_match(np.full((1, 8), 4, np.uint8), ok, ok[0])   # the reference panics: "Unexpected seq_value=4" (:496)
    with pytest.raises(sp.SpError):
        ctx.variant_match(ok, np.full((1, 8), 2, np.uint8), ok[0])   # assert!(hap_value == 0 || hap_value == 1) (:489)
    vm, am = ctx.variant_match(ok, ok, ok[0])                        # the context stays usable
    assert am.tolist() == [[8]] and vm.tolist() == [[0]]


def test_assign_haplotypes_host_vs_oracle():
    from pb_starphase_b200 import _starphase_host as host

    gpu = host.GpuAligner(0)
    rng = np.random.default_rng(12)
    n_var = 387
    seq, hap, vi = random_case(rng, 24, 150, n_var)
    hap[17] = hap[3]                      # two star alleles with the same definition: an exact tie
    hap[0] = 0                            # *1-like: all REF
    seq[0] = hap[3]                       # ties between stars[3] and stars[17]
    seq[1] = 3                            # nothing set: every haplotype scores (0, 0) and ties with the initial Unknown entry
    seq[2] = hap[0]
    stars = [f"{1 + k // 3}.{k % 3:03d}" for k in range(hap.shape[0])]
    lookup = {s: hap[k].tolist() for k, s in enumerate(stars)}
    labels = [f"rs{1000 + v}" for v in range(n_var)]
    meta = [(labels[v], bool(vi[v])) for v in range(n_var)]
    for force in (False, True):
        got = host.assign_haplotypes_from_alleles(gpu, seq.tolist(), lookup, meta, force)
        for s in range(seq.shape[0]):
            star, rv, score = so.assign_haplotype_from_alleles(seq[s].tolist(), lookup, labels, vi.tolist(), force)
            g_star, g_rv, g_score, g_full = got[s]
            assert (g_star, tuple(g_score)) == (star, score)
            assert g_rv == (None if rv is None else so.serde_pretty(rv))
            assert g_full == ("UNKNOWN" if star is None else f"CYP2D6*{star}")
    unforced = host.assign_haplotypes_from_alleles(gpu, seq.tolist(), lookup, meta, False)
    forced = host.assign_haplotypes_from_alleles(gpu, seq.tolist(), lookup, meta, True)
    assert unforced[0][0] is None and forced[0][0] == min(stars[3], stars[17], key=lambda x: f"CYP2D6*{x}".encode())
    assert unforced[1][0] is None and forced[1][3] == min(["UNKNOWN"] + [f"CYP2D6*{s}" for s in stars], key=lambda x: x.encode())
    assert forced[2][0] == stars[0] and json.loads(forced[2][1]) == []   # REF everywhere: nothing to report
