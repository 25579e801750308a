"""CPU suite: pins the oracle (oracle/sp_oracle.c) against itself (DP vs Myers) and against the
golden vectors re-expressed from the reference's own tests (tests/golden)."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def rnd(rng, n):
    return bytes(rng.choice(list(b"ACGT"), n).tolist())


def test_dp_equals_myers_random(oracle):
    rng = np.random.default_rng(7)
    for _ in range(250):
        m, n = int(rng.integers(0, 260)), int(rng.integers(0, 330))
        P = bytearray(rnd(rng, m))
        T = bytearray(rnd(rng, n // 2)) + P + bytearray(rnd(rng, n // 2))
        for _ in range(int(rng.integers(0, 12))):
            if T:
                T[int(rng.integers(0, len(T)))] = int(rng.choice(list(b"ACGTN")))
        for prefix in (False, True):
            assert oracle.infix(bytes(P), bytes(T), prefix, "dp") == oracle.infix(bytes(P), bytes(T), prefix, "myers")


def test_known_small_cases(oracle):
    # exact substring, one substitution, one deletion from the pattern, N never matches
    assert oracle.infix(b"ACGT", b"TTACGTTT") == (0, 6)
    assert oracle.infix(b"ACGT", b"TTAGGTTT")[0] == 1
    assert oracle.infix(b"ACGT", b"TTAGTTT")[0] == 1
    assert oracle.infix(b"ACGT", b"TTACNTTT")[0] == 1
    assert oracle.infix(b"ACNT", b"TTACNTTT")[0] == 1
    assert oracle.infix(b"", b"ACGT") == (0, 0)
    assert oracle.infix(b"ACGT", b"") == (4, 0)
    # prefix mode: placement has to start at text position 0
    assert oracle.infix(b"ACGT", b"TTACGT", prefix=True)[0] == 2
    assert oracle.infix(b"ACGT", b"ACGTTT", prefix=True) == (0, 4)


def test_block_boundaries(oracle):
    rng = np.random.default_rng(11)
    for m in (1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025):
        P = rnd(rng, m)
        T = rnd(rng, 50) + P + rnd(rng, 50)
        for impl in ("myers", "dp"):
            d, e = oracle.infix(P, T, impl=impl)
            assert d == 0 and e <= 50 + m
            if m >= 31:  # a chance earlier occurrence is impossible in practice
                assert e == 50 + m


def test_golden_hla_faux_reference_alleles(oracle):
    """src/hla/caller.rs:1709-1773: an exact-copy read scores (len, 0, 0) against its own allele,
    for HLA-A (forward) and HLA-B (read given as the reverse complement, gene on the reverse strand)."""
    g = json.loads((GOLDEN / "hla_faux.json").read_text())
    for case in g["expected"]["test_reference_alleles"]:
        a = g["hla_sequences"][case["hla_id"]]
        dna, cdna = a["dna_sequence"].encode(), a["cdna_sequence"].encode()
        read = revcomp(dna) if case["read_is_revcomp"] else dna
        target = revcomp(read) if case["read_is_revcomp"] else read  # score_read puts the read on the gene strand
        assert oracle.infix(dna, target, impl="myers")[0] == 0
        assert oracle.infix(dna, target, impl="dp")[0] == 0
        # the other gene's allele must not be a perfect hit
        for other_id, other in g["hla_sequences"].items():
            if other_id != case["hla_id"]:
                assert oracle.infix(other["dna_sequence"].encode(), target, impl="myers")[0] > 0
        assert len(cdna) in (1098, 1089)


def test_golden_bad_read(oracle):
    """src/hla/caller.rs:1783-1809: a 4-bp read matches nothing; the distance is ~|allele| so the
    score (nm+unmapped)/len is at the worst value (>= 0.99, reported as 1.0 when minimap2 maps nothing)."""
    g = json.loads((GOLDEN / "hla_faux.json").read_text())
    for a in g["hla_sequences"].values():
        dna = a["dna_sequence"].encode()
        d, _ = oracle.infix(dna, b"ACGT", impl="myers")
        assert len(dna) - 4 <= d <= len(dna)


def test_golden_weight_sequence(oracle):
    """src/cyp2d6/chaining.rs:1050-1080: P = read segment (fully explained), T = consensus."""
    g = json.loads((GOLDEN / "weight_sequence.json").read_text())
    cons = [c.encode() for c in g["consensuses"]]
    q0, q1 = (q.encode() for q in g["queries"])
    s0 = [oracle.infix(q0, c)[0] for c in cons]
    assert s0[0] == 0 and s0[1] == 1 and s0[2] == 1
    s1 = [oracle.infix(q1, c)[0] for c in cons]
    assert s1[0] == s1[1] == s1[2] == 1


def test_pair_minsum_against_numpy(oracle):
    rng = np.random.default_rng(3)
    D = rng.integers(0, 50, size=(37, 23)).astype(np.int32)
    S = oracle.pair_minsum_full(D)
    ref = np.zeros_like(S)
    for i in range(23):
        for j in range(i, 23):
            ref[i, j] = np.minimum(D[:, i], D[:, j]).sum()
    assert (S == ref).all()
    top = oracle.pair_minsum_topk(D, 10)
    flat = sorted((int(ref[i, j]), i, j) for i in range(23) for j in range(i, 23))[:10]
    assert [(s, i, j) for s, i, j, _ in top] == flat
    for s, i, j, c1 in top:
        assert c1 == int((D[:, i] <= D[:, j]).sum())


def test_span_convention(oracle):
    """sp_oracle_span: end = leftmost end of a best placement, start = rightmost start of a best placement
    ending there; checked against brute force over all substrings on small inputs."""
    rng = np.random.default_rng(3)

    def lev(a: bytes, b: bytes) -> int:
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[-1] + 1, prev[j - 1] + (0 if ca == cb and ca in b"ACGT" else 1)))
            prev = cur
        return prev[-1]

    for _ in range(60):
        P = rnd(rng, int(rng.integers(1, 9)))
        T = bytes(rng.choice(list(b"ACGTN"), int(rng.integers(0, 14))).tolist())
        D, S, E = oracle.score_spans([T], [P])
        d, s, e = int(D[0, 0]), int(S[0, 0]), int(E[0, 0])
        best = min(lev(P, T[a:b]) for a in range(len(T) + 1) for b in range(a, len(T) + 1))
        assert d == best
        ends = sorted({b for a in range(len(T) + 1) for b in range(a, len(T) + 1) if lev(P, T[a:b]) == best})
        assert e == ends[0]
        starts = [a for a in range(e + 1) if lev(P, T[a:e]) == best]
        assert s == max(starts)


# ---- sp_oracle_align: the traceback checker of sp_align_pairs (K4) ---------------------------------
def _replay(pattern: bytes, text: bytes, a: dict):
    """Walks the CIGAR over the two sequences; returns (pattern bases used, text bases used, edits)."""
    i, j, edits = a["p_start"], a["t_start"], 0
    for ln, op in a["cigar"]:
        for _ in range(ln):
            if op in (7, 8):
                same = pattern[i:i + 1].upper() == text[j:j + 1].upper() and pattern[i:i + 1].upper() in (b"A", b"C", b"G", b"T")
                assert same == (op == 7), (i, j, op)
                edits += op == 8
                i += 1; j += 1
            elif op == 1:
                i += 1; edits += 1
            else:
                assert op == 2
                j += 1; edits += 1
    return i, j, edits


def test_align_oracle_invariants(oracle):
    rng = np.random.default_rng(17)
    for trial in range(60):
        m, n = int(rng.integers(0, 300)), int(rng.integers(0, 400))
        p = rnd(rng, m)
        t = rnd(rng, n)
        if trial % 3 == 0 and m:  # a noisy copy with flanks, sometimes with N
            b = bytearray(p)
            for _ in range(int(rng.integers(0, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.choice(list(b"ACGTN")))
            t = rnd(rng, int(rng.integers(0, 40))) + bytes(b) + rnd(rng, int(rng.integers(0, 40)))
        a = oracle.align(p, t)
        d, e = oracle.infix(p, t)
        assert a["dist"] == d and a["t_end"] == e
        i, j, edits = _replay(p, t, a)
        assert (i, j) == (a["p_end"], a["t_end"]) and edits == a["nm"]
        assert a["nm"] + a["p_start"] + (len(p) - a["p_end"]) == d
        assert all(op in (1, 2, 7, 8) and ln > 0 for ln, op in a["cigar"])
        assert all(a["cigar"][k][1] != a["cigar"][k + 1][1] for k in range(len(a["cigar"]) - 1))
        if a["cigar"]:  # clips were stripped: the aligned part neither starts nor ends with an insertion
            assert a["cigar"][0][1] != 1 and a["cigar"][-1][1] != 1


def test_align_oracle_known(oracle):
    assert oracle.align(b"ACGT", b"TTACGTTT") == dict(dist=0, nm=0, p_start=0, p_end=4, t_start=2, t_end=6, cigar=[(4, 7)])
    # one mismatch in the middle
    assert oracle.align(b"ACGTACGT", b"GGACGAACGTGG")["cigar"] == [(3, 7), (1, 8), (4, 7)]
    # homopolymer deletion in the pattern is left-aligned: text AAAAA vs pattern AAAA inside unique flanks
    a = oracle.align(b"CGTAAAATGC", b"CGTAAAAATGC")
    assert a["dist"] == 1 and a["cigar"] == [(3, 7), (1, 2), (7, 7)]
    # pattern base missing from the text: insertion, also left-aligned
    a = oracle.align(b"CGTAAAAATGC", b"CGTAAAATGC")
    assert a["dist"] == 1 and a["cigar"] == [(3, 7), (1, 1), (7, 7)]
    # nothing aligns: everything clipped
    assert oracle.align(b"ACGT", b"") == dict(dist=4, nm=0, p_start=4, p_end=4, t_start=0, t_end=0, cigar=[])
    assert oracle.align(b"", b"ACGT") == dict(dist=0, nm=0, p_start=0, p_end=0, t_start=0, t_end=0, cigar=[])


def test_align_oracle_feeds_process_mm_cigar(oracle):
    """The CIGAR + clips drive the restated HlaProcessedMatch.add_mapping (src/hla/processed_match.rs:53-100): the
    last prefix-edit count equals nm + unmapped whenever the clipped allele ends fit inside the consensus."""
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
    import starphase_oracle as so

    rng = np.random.default_rng(23)
    allele = rnd(rng, 400)
    cons = rnd(rng, 50) + allele[:200] + b"T" + allele[200:390] + rnd(rng, 60)
    a = oracle.align(allele, cons)
    m = so.Mapping(a["p_start"], a["p_end"], len(allele), a["t_start"], a["t_end"], len(cons), a["nm"], True, a["cigar"])
    pm = so.HlaProcessedMatch.worst_match(0)
    pm.add_mapping(m)
    pc = pm.processed_cigars[0]
    assert len(pc) == len(cons) + 1 and pc[0] == 0 and pc[-1] == a["dist"]
    assert pm.full_mapping_stats[0] == so.MappingStats(len(allele), a["nm"], len(allele) - (a["p_end"] - a["p_start"]))
