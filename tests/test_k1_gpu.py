"""GPU parity tests for K1 (batched infix edit distance), through the C ABI, against the CPU
oracle on the same seeded inputs.  Bar: bit-exact integers."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

import pb_starphase_b200 as sp
from pb_starphase_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def rnd(rng, n):
    return bytes(rng.choice(list(b"ACGT"), n).tolist())


def noisy_copy(rng, s: bytes, n_edits: int) -> bytes:
    b = bytearray(s)
    for _ in range(n_edits):
        if not b:
            break
        pos = int(rng.integers(0, len(b)))
        r = rng.random()
        if r < 0.4:
            b[pos] = int(rng.choice(list(b"ACGTN")))
        elif r < 0.7:
            del b[pos]
        else:
            b.insert(pos, int(rng.choice(list(b"ACGT"))))
    return bytes(b)


def edge_case_sets(rng):
    lens = [0, 1, 2, 7, 8, 9, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 512, 513, 1000, 1023, 1024, 1025, 2047, 2049]
    patterns = [rnd(rng, m) for m in lens]
    patterns += [b"N" * 5, b"ACGTNNNNACGT", b"A" * 70, b"acgtacgtacgt"]
    texts = [b"", b"A", b"ACGT", b"N" * 17, rnd(rng, 8), rnd(rng, 15), rnd(rng, 16), rnd(rng, 17)]
    for p in patterns[3::3]:
        texts.append(rnd(rng, int(rng.integers(0, 40))) + noisy_copy(rng, p, int(rng.integers(0, 6))) + rnd(rng, int(rng.integers(0, 40))))
    texts.append(patterns[-1].upper() + b"TT")
    return texts, patterns


@pytest.mark.parametrize("prefix", [False, True])
def test_edge_cases_vs_dp(ctx, oracle, prefix):
    rng = np.random.default_rng(101)
    texts, patterns = edge_case_sets(rng)
    mode = sp.SP_PREFIX if prefix else sp.SP_INFIX
    D, E = ctx.score_batch(texts, patterns, mode=mode, want_end_col=True)
    Dref, Eref = oracle.score_batch(texts, patterns, prefix=prefix, impl="dp", want_end_col=True)
    assert D.shape == (len(texts), len(patterns))
    assert (D == Dref).all(), np.argwhere(D != Dref)[:10]
    assert (E == Eref).all(), np.argwhere(E != Eref)[:10]


@pytest.mark.parametrize("force_u", [str(u) for u in range(4, 17)])
def test_every_lane_width(ctx, oracle, force_u, monkeypatch):
    """Each compiled lane width (U words of 32 rows per lane) must give identical distances."""
    monkeypatch.setenv("SP_FORCE_U", force_u)
    rng = np.random.default_rng(int(force_u))
    patterns = [rnd(rng, int(m)) for m in rng.integers(1, 2200, size=40)]
    texts = [rnd(rng, 30) + noisy_copy(rng, patterns[int(rng.integers(0, 40))], int(rng.integers(0, 30))) + rnd(rng, 30)
             for _ in range(24)]
    D, E = ctx.score_batch(texts, patterns, want_end_col=True)
    Dref, Eref = oracle.score_batch(texts, patterns, impl="myers", want_end_col=True)
    assert (D == Dref).all() and (E == Eref).all()


def test_golden_hla_faux(ctx):
    """Reference known-answer test src/hla/caller.rs:1709-1773: exact-copy reads => distance 0
    against their own allele (DNA and cDNA-of-read), HLA-B through the reverse complement."""
    g = json.loads((GOLDEN / "hla_faux.json").read_text())
    ids = sorted(g["hla_sequences"])
    dna = [g["hla_sequences"][i]["dna_sequence"].encode() for i in ids]
    cdna = [g["hla_sequences"][i]["cdna_sequence"].encode() for i in ids]
    reads = []
    for case in g["expected"]["test_reference_alleles"]:
        a = g["hla_sequences"][case["hla_id"]]["dna_sequence"].encode()
        read = a.translate(COMP)[::-1] if case["read_is_revcomp"] else a
        reads.append(read.translate(COMP)[::-1] if case["read_is_revcomp"] else read)  # gene-strand orientation
    D = ctx.score_batch(reads, dna)
    for k, case in enumerate(g["expected"]["test_reference_alleles"]):
        col = ids.index(case["hla_id"])
        assert D[k, col] == 0
        assert (np.delete(D[k], col) > 0).all()
    # cDNA targets: the allele's own cDNA scored against itself is (len, 0, 0)
    Dc = ctx.score_batch(cdna, cdna)
    assert (np.diag(Dc) == 0).all()
    # test_score_bad_read (src/hla/caller.rs:1783-1809): nothing matches a 4-bp read
    Db = ctx.score_batch([b"ACGT"], dna)
    for col, s in enumerate(dna):
        assert len(s) - 4 <= Db[0, col] <= len(s)


def test_golden_weight_sequence(ctx):
    """src/cyp2d6/chaining.rs:1050-1080 with the roles of that call site: P = segment, T = consensus."""
    g = json.loads((GOLDEN / "weight_sequence.json").read_text())
    cons = [c.encode() for c in g["consensuses"]]
    queries = [q.encode() for q in g["queries"]]
    D = ctx.score_batch(cons, queries)  # D[consensus, query]
    assert D[:, 0].tolist() == [0, 1, 1]
    assert D[:, 1].tolist() == [1, 1, 1]


def test_hla_like_vs_myers(ctx, oracle):
    """Mutation-tree alleles of HLA-B-like lengths x HiFi-like reads, full matrix vs the oracle."""
    alleles, reads, src = synth.hla_gene(synth.DEFAULT_SEED, "HLA-B", n_alleles=300, n_reads=48)
    D, E = ctx.score_batch(reads, alleles, want_end_col=True)
    Dref, Eref = oracle.score_batch(reads, alleles, impl="myers", want_end_col=True)
    assert (D == Dref).all()
    assert (E == Eref).all()
    # every read is closest (or tied) to an allele within a few edits of its source
    assert (D[np.arange(len(reads)), src] <= 40).all()


def test_device_matrix_u16_and_i32_agree(ctx, oracle):
    alleles, reads, _ = synth.hla_gene(5, "HLA-A", n_alleles=120, n_reads=70)
    P, T = ctx.patterns(alleles), ctx.targets(reads)
    d16 = ctx.score_device(T, P, elem_bits=16)
    d32 = ctx.score_device(T, P, elem_bits=32)
    a, b = d16.to_host(), d32.to_host()
    assert (a == b).all()
    assert (a == oracle.score_batch(reads, alleles)).all()
    assert d16.ld % 64 == 0 and d16.elem_bits == 16 and d16.device_ptr != 0
    # 16-bit read-back into pinned memory from sp_host_alloc, and int32 read-back into a caller buffer
    pin = ctx.pinned_empty(a.shape, np.uint16)
    assert (d16.to_host_u16(pin) == a).all() and (d16.to_host_u16() == a).all()
    out = ctx.pinned_empty(a.shape, np.int32)
    assert (d32.to_host(out=out) == a).all()
    with pytest.raises(sp.SpError):
        d32.to_host_u16()


def test_sharding_invariance(ctx):
    """Allele-range shards (the multi-GPU partition of K1) reproduce the unsharded matrix."""
    alleles, reads, _ = synth.hla_gene(9, "HLA-A", n_alleles=150, n_reads=20)
    whole = ctx.score_batch(reads, alleles)
    parts = [ctx.score_batch(reads, alleles[lo:hi]) for lo, hi in ((0, 40), (40, 111), (111, 150))]
    assert (np.concatenate(parts, axis=1) == whole).all()


def test_empty_inputs(ctx):
    assert ctx.score_batch([], [b"ACGT"]).shape == (0, 1)
    assert ctx.score_batch([b"ACGT"], []).shape == (1, 0)
    assert ctx.score_batch([b"ACGT", b""], [b"", b"AC"]).tolist() == [[0, 0], [0, 2]]


def test_too_long_pattern_is_an_error(ctx):
    with pytest.raises(sp.SpError) as ei:
        ctx.score_batch([b"ACGT"], [b"A" * 24577])
    assert ei.value.status == 3


def test_class_ii_sized_alleles(ctx, oracle):
    """DRB1-sized genomic alleles (11-17 kb) next to ordinary ones.  Lane widths 20 / 24 hold patterns up to 24,576 rows;
    K1 distances and end columns, K3 spans and K4 tracebacks stay bit-exact against the oracle."""
    rng = np.random.default_rng(41)
    big = rnd(rng, 17011)
    pats = [big, noisy_copy(rng, big, 40)[:16385], noisy_copy(rng, big, 9), rnd(rng, 3100), big[2000:14000], rnd(rng, 24576)]
    texts = [rnd(rng, 300) + noisy_copy(rng, big, 25) + rnd(rng, 200), big[500:16000], rnd(rng, 5000)]
    D, E = ctx.score_batch(texts, pats, want_end_col=True)
    Dw, Ew = oracle.score_batch(texts, pats, want_end_col=True)
    assert (D == Dw).all() and (E == Ew).all()
    got = ctx.score_spans(texts[:2], pats[:3])
    for g, w, name in zip(got, oracle.score_spans(texts[:2], pats[:3]), ("distance", "start", "end")):
        assert (g == w).all(), name
    pairs = [(0, 0), (0, 1), (1, 2), (0, 3), (1, 4)]
    for (t, p), g in zip(pairs, ctx.align_pairs(texts, pats, pairs)):
        assert g == oracle.align(pats[p], texts[t]), (t, p)


def test_cyp2d6_shapes_roles_swapped(ctx, oracle):
    """K3: CYP2D6 template / segment lengths (src/cyp2d6/definitions.rs:137-172): patterns up to 6,165 bp."""
    rng = np.random.default_rng(22)
    d6 = rnd(rng, 6165)
    d7 = noisy_copy(rng, d6, 180)[:5938]
    templates = [d6, d7, rnd(rng, 3500), rnd(rng, 2772), rnd(rng, 1564), rnd(rng, 2919)]
    reads = [rnd(rng, 400) + noisy_copy(rng, templates[i % 6], 12) + rnd(rng, 300) for i in range(12)]
    D = ctx.score_batch(reads, templates)
    assert (D == oracle.score_batch(reads, templates)).all()
    # roles swapped (weight_sequence, src/cyp2d6/chaining.rs:48-94): segments are the patterns
    segs = [noisy_copy(rng, d6, 5), noisy_copy(rng, d7, 9)]
    D2 = ctx.score_batch(templates, segs)
    assert (D2 == oracle.score_batch(templates, segs)).all()


@pytest.mark.skipif(os.environ.get("SP_SKIP_LARGE") == "1", reason="large case skipped")
def test_large_sampled_parity(ctx, oracle):
    """HLA-A-like set at a size where the full oracle matrix is too slow: exact-copy property on all
    pairs plus oracle equality on a seeded sample of rows."""
    alleles, reads, src = synth.hla_gene(77, "HLA-A", n_alleles=1500, n_reads=96)
    reads = list(reads)
    for k in range(0, 96, 8):  # error-free reads: distance to the source allele must be exactly 0
        reads[k] = rnd(np.random.default_rng(k), 100) + alleles[int(src[k])] + rnd(np.random.default_rng(k + 1), 100)
    D = ctx.score_batch(reads, alleles)
    for k in range(0, 96, 8):
        assert D[k, int(src[k])] == 0
    rows = [0, 5, 17, 40, 95]
    Dref = oracle.score_batch([reads[r] for r in rows], alleles)
    assert (D[rows] == Dref).all()
    assert (D >= 0).all() and (D <= np.array([len(a) for a in alleles])[None, :]).all()


def test_targets_derive_splice_and_revcomp(ctx, oracle):
    """sp_targets_derive: spliced / reverse-complemented sequences built on the device from a resident set equal the host's
    (splice_read's exon concatenation, reverse_complement: src/hla/caller.rs:1518-1576, :1337-1368), and score like them."""
    import pb_starphase_b200 as sp

    rng = np.random.default_rng(41)
    srcs = [bytes(rng.choice(list(b"ACGTN"), n, p=[.24, .24, .24, .24, .04]).tolist()) for n in (700, 1, 0, 1300, 64)]
    srcs.append(srcs[0][:200].lower())
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    pieces = [(0, [(10, 80), (200, 200), (300, 512), (690, 700)]), (3, [(0, 1300)]), (3, [(5, 6)]), (1, [(0, 1)]), (2, []), (4, [(0, 64), (0, 64)]),
              (5, [(3, 190)]), (0, [])]
    revcomp = [False, True, True, False, False, True, True, True]
    want = []
    for (s, ivs), rc in zip(pieces, revcomp):
        x = b"".join(srcs[s][b:e] for b, e in ivs)
        want.append(x[::-1].translate(comp) if rc else x)
    T = ctx.targets(srcs)
    D = T.derive(pieces, revcomp)
    assert D.read() == want and D.n == len(want) and D.total_len == sum(map(len, want))
    assert T.derive(pieces[:2]).read() == [b"".join(srcs[s][b:e] for b, e in ivs) for s, ivs in pieces[:2]]  # revcomp NULL
    pats = [srcs[0][300:512], want[1][100:400], b"ACGTACGT"]
    P = ctx.patterns(pats)
    M = ctx.score_device(D, P, 32)
    got = M.to_host()
    assert (got == oracle.score_batch(want, pats)).all()
    M.close(); P.close()
    for bad in ([(9, [(0, 1)])], [(0, [(5, 3)])], [(0, [(0, 701)])]):
        with pytest.raises(sp.SpError):
            T.derive(bad)
    D.close(); T.close()


def test_share_device_mode_is_bit_identical(oracle):
    """sp_ctx_share_device: K1 as one CTA per item (short launches on the context stream, launches of several rounds on the
    lowest-priority side stream) gives the matrix of the persistent grid, with two contexts scoring from two host threads."""
    import threading

    alleles, reads, _ = synth.hla_gene(1234, "HLA-A", n_alleles=1200, n_reads=40)  # ~38 pattern groups x ~40 text tiles: the bulk route
    small_p, small_t = alleles[:9], reads[:5]                                     # a handful of items: stays on the main stream
    with sp.Context(0) as plain:
        want, want_e = plain.score_batch(reads, alleles, want_end_col=True)
        want_small = plain.score_batch(small_t, small_p)
    assert (want[:6, :40] == oracle.score_batch(reads[:6], alleles[:40])).all()
    got = {}

    def run(tag):
        with sp.Context(0) as c:
            c.share_device(True)
            for _ in range(2):
                got[tag] = (c.score_batch(reads, alleles, want_end_col=True), c.score_batch(small_t, small_p))

    threads = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for k in range(2):
        (D, E), Ds = got[k]
        assert (D == want).all() and (E == want_e).all() and (Ds == want_small).all()
    with sp.Context(0) as c:  # switching the mode off again restores the persistent grid
        c.share_device(True)
        c.share_device(False)
        assert (c.score_batch(small_t, small_p) == want_small).all()
