"""GPU parity tests for K3 (CYP2D6 candidate scoring = K1 with the roles of the call site, plus the
text span of the optimal placement), through the C ABI, against the CPU oracle.  Bit-exact integers."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"

# template lengths of src/cyp2d6/definitions.rs:13-14, :137-172 (D6, D7, *5 signature, REP6/7, spacer, link)
CYP_LENS = dict(d6=6165, d7=5938, star5=3500, rep=2772, spacer=1564, link=2919)


def rnd(rng, n):
    return bytes(rng.choice(list(b"ACGT"), n).tolist())


def noisy(rng, s: bytes, n_edits: int) -> bytes:
    b = bytearray(s)
    for _ in range(n_edits):
        pos = int(rng.integers(0, max(len(b), 1)))
        r = rng.random()
        if r < 0.4 and b:
            b[pos] = int(rng.choice(list(b"ACGTN")))
        elif r < 0.7 and b:
            del b[pos]
        else:
            b.insert(pos, int(rng.choice(list(b"ACGT"))))
    return bytes(b)


def test_spans_small_known(ctx):
    D, S, E = ctx.score_spans([b"TTACGTTT", b"ACGT", b"", b"GGGG"], [b"ACGT", b"", b"ACGTACGT"])
    assert D[0, 0] == 0 and (S[0, 0], E[0, 0]) == (2, 6)
    assert D[1, 0] == 0 and (S[1, 0], E[1, 0]) == (0, 4)
    assert D[2, 0] == 4 and (S[2, 0], E[2, 0]) == (0, 0)          # empty text: nothing placed
    assert D[:, 1].tolist() == [0, 0, 0, 0] and (S[:, 1] == E[:, 1]).all()  # empty pattern
    assert D[1, 2] == 4 and E[1, 2] - S[1, 2] == 4                # pattern longer than the text


def test_spans_vs_oracle_edge_lengths(ctx, oracle):
    rng = np.random.default_rng(5)
    pats = [rnd(rng, m) for m in (1, 2, 31, 32, 33, 100, 511, 512, 513, 700, 1025)] + [b"", b"NNNN", b"A" * 40]
    texts = [b"", b"A", rnd(rng, 17)]
    for p in pats[::2]:
        texts.append(rnd(rng, int(rng.integers(0, 30))) + noisy(rng, p, int(rng.integers(0, 5))) + rnd(rng, int(rng.integers(0, 30))))
    texts.append(pats[3] + rnd(rng, 10) + pats[3])  # the same pattern twice: leftmost end, start paired with it
    texts.append(b"A" * 100)                        # homopolymer: many co-optimal placements
    got = ctx.score_spans(texts, pats)
    ref = oracle.score_spans(texts, pats)
    for g, r, name in zip(got, ref, "DSE"):
        assert (g == r).all(), (name, np.argwhere(g != r)[:10])


def test_weight_sequence_shapes(ctx, oracle):
    """weight_sequence (src/cyp2d6/chaining.rs:28-103): segments (patterns) against consensuses (texts); the
    overlap score 1 - (clip_start + clip_end)/|C| is formed on the host from the returned span."""
    rng = np.random.default_rng(31)
    d6 = rnd(rng, CYP_LENS["d6"])
    d7 = noisy(rng, d6, 185)[: CYP_LENS["d7"]]
    cons = [d6, d7, noisy(rng, d6, 3), noisy(rng, d7, 4), rnd(rng, CYP_LENS["rep"]), rnd(rng, CYP_LENS["spacer"]),
            rnd(rng, CYP_LENS["link"]), rnd(rng, CYP_LENS["star5"])]
    segs = []
    for k in range(20):
        c = cons[k % len(cons)]
        lo = int(rng.integers(0, len(c) // 4))
        hi = len(c) - int(rng.integers(0, len(c) // 4))
        segs.append(noisy(rng, c[lo:hi], int(rng.integers(0, 25))))
    D, S, E = ctx.score_spans(cons, segs)
    Dr, Sr, Er = oracle.score_spans(cons, segs)
    assert (D == Dr).all() and (S == Sr).all() and (E == Er).all()
    # the span length can differ from |segment| only by the number of indels, at most D
    seg_len = np.array([len(s) for s in segs])[None, :]
    assert (np.abs((E - S) - seg_len) <= D).all()
    assert (S >= 0).all() and (E <= np.array([len(c) for c in cons])[:, None]).all() and (S <= E).all()
    # sp_score_spans_filtered: the same numbers, but pairs further apart than 35 % of the segment get start = -1 (no reverse pass):
    # the unrelated (segment, consensus) pairs of this set, for which the reference's aligner would report no mapping
    Df, Sf, Ef = ctx.score_spans(cons, segs, max_dist_permille=350)
    far = D * 1000 > seg_len * 350
    assert far.any() and (~far).any()
    assert (Df == D).all() and (Ef == E).all() and (Sf[~far] == S[~far]).all() and (Sf[far] == -1).all()
    D0, S0, E0 = ctx.score_spans(cons, segs, max_dist_permille=0)   # only exact placements keep a span
    assert ((S0 == -1) == (D > 0)).all() and (S0[D == 0] == S[D == 0]).all()
    # golden: exact copy wins, N at the differing base ties all three (src/cyp2d6/chaining.rs:1050-1080)
    g = json.loads((GOLDEN / "weight_sequence.json").read_text())
    Dg, Sg, Eg = ctx.score_spans([c.encode() for c in g["consensuses"]], [q.encode() for q in g["queries"]])
    assert Dg[:, 0].tolist() == [0, 1, 1] and Dg[:, 1].tolist() == [1, 1, 1]
    assert ((Eg - Sg) == len(g["queries"][0])).all()


def test_template_search_shapes(ctx, oracle):
    """find_base_type_in_sequence (src/cyp2d6/haplotyper.rs:142-315): 39 templates (patterns) against whole
    reads (texts); the read span target_start..target_end is the returned [start, end)."""
    rng = np.random.default_rng(32)
    d6 = rnd(rng, CYP_LENS["d6"])
    d7 = noisy(rng, d6, 185)[: CYP_LENS["d7"]]
    templates = [d6, d7, rnd(rng, CYP_LENS["star5"]), rnd(rng, CYP_LENS["rep"]), rnd(rng, CYP_LENS["rep"]),
                 rnd(rng, CYP_LENS["spacer"]), rnd(rng, CYP_LENS["link"])]
    for k in range(32):  # exon / intron hybrids: D6 prefix + D7 suffix and vice versa
        cut = 300 + 170 * k
        templates.append((d6[:cut] + d7[cut:]) if k % 2 == 0 else (d7[:cut] + d6[cut:]))
    assert len(templates) == 39
    reads = []
    for k in range(6):
        body = noisy(rng, templates[k % 2], 10) + templates[6] + noisy(rng, templates[3], 4)
        reads.append(rnd(rng, 500) + body + rnd(rng, 700))
    D, S, E = ctx.score_spans(reads, templates)
    Dr, Sr, Er = oracle.score_spans(reads, templates)
    assert (D == Dr).all() and (S == Sr).all() and (E == Er).all()
    # reads 0, 2, 4 carry a noisy D6: its span starts right after the 500-bp flank
    for r in (0, 2, 4):
        assert D[r, 0] <= 10 and abs(int(S[r, 0]) - 500) <= 10


def _random_chain_case(rng, n_haps, n_chains, n_reads, max_len=15, max_w=6):
    chains = [rng.integers(0, n_haps, size=int(rng.integers(1, max_len + 1))).tolist() for _ in range(n_chains)]
    reads = []
    for _ in range(n_reads):
        w = int(rng.integers(1, max_w + 1))
        truth = rng.integers(0, n_haps, size=w)
        W = rng.integers(20, 400, size=(w, n_haps)).astype(np.uint32)
        W[np.arange(w), truth] = rng.integers(0, 6, size=w)  # the true consensus of each segment is close
        reads.append(W)
    return chains, reads


def test_chain_windows_vs_oracle(ctx, oracle):
    """containment_score's window scan (src/cyp2d6/chaining.rs:683-731) per (read, chain), incl. chains shorter
    than the read (2 x worst sentinel), single-segment reads and single-member chains."""
    rng = np.random.default_rng(41)
    chains, reads = _random_chain_case(rng, n_haps=7, n_chains=90, n_reads=130)
    chains += [[0], [6, 6, 6], list(range(7))]
    B = ctx.chain_window_scores(chains, reads, 7)
    got = B.to_host()
    want = oracle.chain_windows(chains, reads, 7)
    assert got.shape == (len(reads), len(chains)) and (got == want).all()
    # pair sums on the device matrix == oracle pair sums on the oracle matrix
    S = ctx.pair_minsum_full(B)
    assert (S == oracle.pair_minsum_full(want)).all()
    top = ctx.pair_minsum_topk(B, 10)
    assert top == oracle.pair_minsum_topk(want, 10)


def test_chain_pairs_config4_shape(ctx, oracle):
    """P = 2,000 chains x R_c = 2,000 reads, w <= 6.  Full B against the oracle; the 2e6 pair
    sums against numpy on a seeded sample of pairs plus the oracle's top-10."""
    rng = np.random.default_rng(42)
    chains, reads = _random_chain_case(rng, n_haps=24, n_chains=2000, n_reads=2000)
    B = ctx.chain_window_scores(chains, reads, 24)
    got = B.to_host()
    want = oracle.chain_windows(chains, reads, 24)
    assert (got == want).all()
    S = ctx.pair_minsum_full(B)
    for _ in range(300):
        i = int(rng.integers(0, 2000)); j = int(rng.integers(i, 2000))
        assert int(S[i, j]) == int(np.minimum(want[:, i], want[:, j]).sum())
    assert (np.tril(S, -1) == 0).all()
    assert ctx.pair_minsum_topk(B, 10) == oracle.pair_minsum_topk(want, 10)


def test_chain_windows_empty_and_errors(ctx):
    import pb_starphase_b200 as sp

    B = ctx.chain_window_scores([], [np.zeros((2, 3), np.uint32)], 3)
    assert B.to_host().shape == (1, 0)
    B = ctx.chain_window_scores([[0, 1]], [], 3)
    assert B.to_host().shape == (0, 1)
    with pytest.raises(sp.SpError):
        ctx.chain_window_scores([[0, 5]], [np.zeros((1, 3), np.uint32)], 3)  # hap index outside [0, n_haps)
