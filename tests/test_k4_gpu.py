"""GPU parity tests for K4 (sp_align_pairs: traceback alignment of selected pairs), through the C ABI, against the
CPU oracle's full-matrix traceback (sp_oracle_align).  Bit-exact: spans, nm, clips and every CIGAR entry."""
import numpy as np
import pytest

from test_k3_gpu import noisy, rnd

pytestmark = pytest.mark.gpu


def check_against_oracle(ctx, oracle, texts, pats, pairs):
    got = ctx.align_pairs(texts, pats, pairs)
    assert len(got) == len(pairs)
    for (t, p), g in zip(pairs, got):
        want = oracle.align(pats[p], texts[t])
        assert g == want, (t, p, len(texts[t]), len(pats[p]), {k: (g[k], want[k]) for k in g if g[k] != want[k] and k != "cigar"})
    return got


def test_align_small_known(ctx):
    texts = [b"TTACGTTT", b"ACGT", b"", b"GGACGAACGTGG"]
    pats = [b"ACGT", b"", b"ACGTACGT", b"TTTT"]
    r = ctx.align_pairs(texts, pats, [(0, 0), (1, 0), (2, 0), (0, 1), (3, 2), (1, 3)])
    assert r[0] == dict(dist=0, nm=0, p_start=0, p_end=4, t_start=2, t_end=6, cigar=[(4, 7)])
    assert r[1] == dict(dist=0, nm=0, p_start=0, p_end=4, t_start=0, t_end=4, cigar=[(4, 7)])
    assert r[2] == dict(dist=4, nm=0, p_start=4, p_end=4, t_start=0, t_end=0, cigar=[])      # empty text: all clipped
    assert r[3] == dict(dist=0, nm=0, p_start=0, p_end=0, t_start=0, t_end=0, cigar=[])      # empty pattern
    assert r[4]["dist"] == 1 and sum(l for l, op in r[4]["cigar"] if op in (7, 8, 1)) == r[4]["p_end"] - r[4]["p_start"]
    assert r[5]["dist"] == 3 and r[5]["nm"] + 4 - (r[5]["p_end"] - r[5]["p_start"]) == 3


def test_align_vs_oracle_edge_lengths(ctx, oracle):
    rng = np.random.default_rng(11)
    pats = [rnd(rng, m) for m in (1, 2, 31, 32, 33, 100, 511, 512, 513, 700, 1025, 2049)] + [b"", b"NNNN", b"A" * 40]
    texts = [b"", b"A", rnd(rng, 17), b"A" * 100]
    for p in pats:
        texts.append(rnd(rng, int(rng.integers(0, 30))) + noisy(rng, p, int(rng.integers(0, 6))) + rnd(rng, int(rng.integers(0, 30))))
    texts.append(pats[3] + rnd(rng, 10) + pats[3])
    pairs = [(t, p) for t in range(len(texts)) for p in range(len(pats))]
    check_against_oracle(ctx, oracle, texts, pats, pairs)


def test_align_hla_shaped(ctx, oracle):
    """score_read shape (src/hla/caller.rs:1413-1500): alleles (patterns) of a mutation tree against two consensuses
    with flanks; the clipped ends and CIGARs feed HlaProcessedMatch.add_mapping."""
    from pb_starphase_b200 import synth

    alleles, reads, src, cdna = synth.hla_gene(synth.DEFAULT_SEED, "HLA-A", n_alleles=24, n_reads=2, with_cdna=True)
    rng = np.random.default_rng(2)
    # clipped ends: alleles that overhang the consensus on either side, and an unrelated one
    pats = list(alleles[:20]) + [rnd(rng, 60) + alleles[3], alleles[5] + rnd(rng, 45), rnd(rng, 1200)]
    pairs = [(t, p) for t in range(2) for p in range(len(pats))]
    check_against_oracle(ctx, oracle, list(reads), pats, pairs)
    # without flanks on the consensus the overhanging allele ends cannot be paired with anything: they are clipped
    bare = [alleles[3], alleles[5]]
    got = check_against_oracle(ctx, oracle, bare, pats, [(0, 20), (1, 21)])
    assert got[0]["p_start"] == 60 and got[0]["nm"] == 0 and got[0]["dist"] == 60
    assert got[1]["dist"] == 45 and got[1]["nm"] + len(pats[21]) - got[1]["p_end"] == 45
    ctargets = [cdna[int(s)] for s in src]
    check_against_oracle(ctx, oracle, ctargets, list(cdna[:24]), [(t, p) for t in range(2) for p in range(24)])


def test_align_long_text_window(ctx, oracle):
    """realign_record shape (src/hla/realigner.rs:116-146): a long read against one allele; only the window around
    the placement is kept for the traceback."""
    rng = np.random.default_rng(8)
    allele = rnd(rng, 3300)
    read = rnd(rng, 5200) + noisy(rng, allele, 25) + rnd(rng, 4100)
    read2 = rnd(rng, 300) + noisy(rng, allele[:2000], 12)  # the allele's tail hangs off the read
    check_against_oracle(ctx, oracle, [read, read2], [allele], [(0, 0), (1, 0)])


def test_align_windows_vs_oracle(ctx, oracle):
    """sp_align_windows: the pair is aligned inside T[begin, end) only and the record is relative to begin -- identical to
    aligning the copied sub-string (what the template search of find_base_type_in_sequence did before)."""
    rng = np.random.default_rng(23)
    pats = [rnd(rng, m) for m in (40, 300, 700, 1500)]
    texts = [rnd(rng, 500) + noisy(rng, pats[k % 4], 5) + rnd(rng, 800) + noisy(rng, pats[(k + 1) % 4], 9) + rnd(rng, 300) for k in range(6)]
    pairs, wins = [], []
    for t in range(len(texts)):
        for p in range(len(pats)):
            n = len(texts[t])
            b = int(rng.integers(0, n // 2))
            e = int(rng.integers(b, n + 1))
            pairs.append((t, p)); wins.append((b, e))
    pairs += [(0, 1), (1, 2), (2, 0)]
    wins += [(0, len(texts[0])), (100, 100), (len(texts[2]), len(texts[2]))]   # the whole text, two empty windows
    got = ctx.align_pairs(texts, pats, pairs, windows=wins)
    for (t, p), (b, e), g in zip(pairs, wins, got):
        assert g == oracle.align(pats[p], texts[t][b:e]), (t, p, b, e)
    import pb_starphase_b200 as sp

    for bad in ((-1, 10), (10, 5), (0, len(texts[0]) + 1)):
        with pytest.raises(sp.SpError):
            ctx.align_pairs(texts, pats, [(0, 0)], windows=[bad])


def test_align_cyp_sized_pattern(ctx, oracle):
    rng = np.random.default_rng(9)
    d6 = rnd(rng, 6165)
    read = rnd(rng, 150) + noisy(rng, d6, 60) + rnd(rng, 90)
    check_against_oracle(ctx, oracle, [read], [d6], [(0, 0)])


def test_align_matches_k1_and_spans(ctx):
    """dist / t_end agree with K1's distance and end column; [t_start, t_end) is a valid optimal span."""
    rng = np.random.default_rng(4)
    pats = [rnd(rng, int(m)) for m in rng.integers(50, 900, 12)]
    texts = [rnd(rng, 20) + noisy(rng, pats[i % 12], 7) + rnd(rng, 25) for i in range(6)]
    D, E = ctx.score_batch(texts, pats, want_end_col=True)
    pairs = [(t, p) for t in range(6) for p in range(12)]
    for (t, p), r in zip(pairs, ctx.align_pairs(texts, pats, pairs)):
        assert r["dist"] == D[t, p] and r["t_end"] == E[t, p]
        n_t = sum(l for l, op in r["cigar"] if op in (7, 8, 2))
        n_p = sum(l for l, op in r["cigar"] if op in (7, 8, 1))
        assert n_t == r["t_end"] - r["t_start"] and n_p == r["p_end"] - r["p_start"]
        assert r["nm"] == sum(l for l, op in r["cigar"] if op != 7)


def test_align_bad_arguments(ctx):
    import pb_starphase_b200 as sp

    with pytest.raises(sp.SpError):
        ctx.align_pairs([b"ACGT"], [b"AC"], [(1, 0)])
    with pytest.raises(sp.SpError):
        ctx.align_pairs([b"ACGT"], [b"A" * 24577], [(0, 0)])
    assert ctx.align_pairs([b"ACGT"], [b"AC"], []) == []
