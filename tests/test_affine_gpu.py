"""GPU suite: K9 (sp_align_affine_resident: banded two-piece affine local alignment, the reference's cost model) through the C ABI
against the unbanded CPU model (oracle/sp_oracle_affine.c): score, nm, clips, spans and every CIGAR entry, for both cost sets of
the reference (map-hifi a=1, allele scoring a=5)."""
import numpy as np
import pytest

import oracle_util as ou
from pb_starphase_b200 import synth
from test_k3_gpu import noisy, rnd

pytestmark = pytest.mark.gpu


def check(ctx, texts, pats, pairs, costs, band, centres=None):
    aff = ou.AffineOracle(costs)
    T, P = ctx.targets(texts), ctx.targets(pats)
    got = ctx.align_affine(T, P, pairs, costs, band=band, centres=centres)
    for (t, p), g in zip(pairs, got):
        want = aff.align(pats[p], texts[t])
        if want["score"] == 0:
            want = dict(want, dist=len(pats[p]))
        assert g == want, (t, p, len(texts[t]), len(pats[p]), {k: (g[k], want[k]) for k in g if g[k] != want[k] and k != "cigar"})
    T.close(); P.close()
    return got


@pytest.mark.parametrize("costs", [ou.COSTS_MAP_HIFI, ou.COSTS_ALLELE_SCORING])
def test_affine_small_full_band(ctx, costs):
    """Band wider than both sequences: the banded DP is the full DP, ties included (unrelated and related short pairs)."""
    rng = np.random.default_rng(31)
    pats = [rnd(rng, m) for m in (1, 2, 7, 31, 32, 33, 60, 90)] + [b"NNNN", b"A" * 30, b"ACGT" * 12]
    texts = [b"A", rnd(rng, 17), b"A" * 40, b"ACGT" * 15]
    for p in pats[:8]:
        texts.append(rnd(rng, int(rng.integers(0, 20))) + noisy(rng, p, int(rng.integers(0, 5))) + rnd(rng, int(rng.integers(0, 20))))
    pairs = [(t, p) for t in range(len(texts)) for p in range(len(pats))]
    centres = [0] * len(pairs)
    check(ctx, texts, pats, pairs, costs, 140, centres)


@pytest.mark.parametrize("costs", [ou.COSTS_MAP_HIFI, ou.COSTS_ALLELE_SCORING])
def test_affine_hla_shaped(ctx, oracle, costs):
    """Alleles against consensus-like targets with flanks, band centred on the unit-cost placement (what the host does with K4's
    t_start - p_start): clipped ends, consolidated gaps, the two-piece switch."""
    alleles, reads, src, cdna = synth.hla_gene(synth.DEFAULT_SEED, "HLA-A", n_alleles=24, n_reads=3, with_cdna=True)
    rng = np.random.default_rng(2)
    gap = alleles[4][:1500] + alleles[4][1540:]                       # a 40-base deletion: second gap piece
    pats = list(alleles[:16]) + [rnd(rng, 60) + alleles[3], alleles[5] + rnd(rng, 45), gap]
    def banded(texts, patterns, all_pairs, slack):
        """pairs whose unit-cost placement fits a band of half-width <= 255 around the middle of its start and end diagonals"""
        keep, centres, width = [], [], 0
        for t, p in all_pairs:
            u = oracle.align(patterns[p], texts[t])
            d0, d1 = u["t_start"] - u["p_start"], u["t_end"] - u["p_end"]
            w = (abs(d1 - d0) + 1) // 2 + u["nm"] + slack
            if w <= 255:
                keep.append((t, p)); centres.append((d0 + d1) // 2); width = max(width, w)
        return keep, centres, width

    pairs, centres, w = banded(list(reads), pats, [(t, p) for t in range(3) for p in range(len(pats))], 24)
    assert len(pairs) >= 15
    check(ctx, list(reads), pats, pairs, costs, w, centres)
    ctargets = [cdna[int(s)] for s in src]
    pairs, centres, w = banded(ctargets, list(cdna[:16]), [(t, p) for t in range(3) for p in range(16)], 24)
    assert len(pairs) >= 15
    check(ctx, ctargets, list(cdna[:16]), pairs, costs, w, centres)


def test_affine_windows_and_nothing_to_align(ctx):
    rng = np.random.default_rng(5)
    pat = rnd(rng, 300)
    text = rnd(rng, 400) + noisy(rng, pat, 4) + rnd(rng, 500)
    T, P = ctx.targets([text, b"", b"TTTT"]), ctx.targets([pat, b"", b"GGGG"])
    aff = ou.AffineOracle(ou.COSTS_MAP_HIFI)
    got = ctx.align_affine(T, P, [(0, 0)], ou.COSTS_MAP_HIFI, band=48, centres=[0], windows=[(380, 730)])
    want = aff.align(pat, text[380:730])
    assert got[0] == want
    got = ctx.align_affine(T, P, [(1, 0), (0, 1), (2, 2)], ou.COSTS_MAP_HIFI, band=16)
    assert [g["score"] for g in got] == [0, 0, 0] and [g["dist"] for g in got] == [300, 0, 4] and all(g["cigar"] == [] for g in got)
    import pb_starphase_b200 as sp

    with pytest.raises(sp.SpError):
        ctx.align_affine(T, P, [(0, 0)], ou.COSTS_MAP_HIFI, band=256)
    T.close(); P.close()


def test_affine_band_per_pair(ctx, oracle):
    """pair_band: every pair inside its own band (all four width classes in one call, run side by side), each equal to the CPU
    model banded the same way -- including bands too narrow for the optimum, where the band decides the answer."""
    alleles, reads, src, cdna = synth.hla_gene(synth.DEFAULT_SEED, "HLA-B", n_alleles=20, n_reads=4, with_cdna=True)
    texts, pats = list(reads), list(alleles) + list(cdna[:6])
    pairs = [(t, p) for t in range(len(texts)) for p in range(len(pats))]
    widths = [3, 20, 31, 32, 63, 64, 100, 127, 128, 200, 255]
    bands = [widths[k % len(widths)] for k in range(len(pairs))]
    centres = []
    for t, p in pairs:
        u = oracle.align(pats[p], texts[t])
        centres.append((u["t_start"] - u["p_start"] + u["t_end"] - u["p_end"]) // 2)
    aff = ou.AffineOracle(ou.COSTS_ALLELE_SCORING)
    T, P = ctx.targets(texts), ctx.targets(pats)
    got = ctx.align_affine(T, P, pairs, ou.COSTS_ALLELE_SCORING, centres=centres, bands=bands)
    for (t, p), c, b, g in zip(pairs, centres, bands, got):
        want = aff.align(pats[p], texts[t], centre=c, band=b)
        if want["score"] == 0:
            want = dict(want, dist=len(pats[p]))
        assert g == want, (t, p, c, b, {k: (g[k], want[k]) for k in g if g[k] != want[k] and k != "cigar"})
    import pb_starphase_b200 as sp

    with pytest.raises(sp.SpError):
        ctx.align_affine(T, P, pairs[:2], ou.COSTS_ALLELE_SCORING, bands=[10, 256])
    T.close(); P.close()
