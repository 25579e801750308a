"""CPU suite: the affine-gap model oracle (oracle/sp_oracle_affine.c) is pinned against an independent pure-Python Gotoh DP and
against hand-computed cases, and the measured distance between the product's unit-cost numbers and that cost model is held:
the call-level flip counts recorded in DESIGN.md §3 / profiles/r02_affine_divergence.json (0 best-allele flips, 0 diplotype
flips, 0 weight_sequence argmin flips on the seeded cases below) may not rise.  Parity with minimap2 itself stays unpinned."""
import json
from pathlib import Path

import numpy as np
import pytest

import affine_model as am
import flow_oracle as fo
import oracle_util as ou
from pb_starphase_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
NEG = -10 ** 9


def py_affine_local_score(P: bytes, T: bytes, costs):
    """Independent restatement: best local alignment score under match a / mismatch -b / ambiguous -1 / gap -min(q+k*e, q2+k*e2)."""
    a, b, q, e, q2, e2 = costs
    m, n = len(P), len(T)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    H = [[0] * (n + 1) for _ in range(m + 1)]
    E1 = [[NEG] * (n + 1) for _ in range(m + 1)]
    E2 = [[NEG] * (n + 1) for _ in range(m + 1)]
    F1 = [[NEG] * (n + 1) for _ in range(m + 1)]
    F2 = [[NEG] * (n + 1) for _ in range(m + 1)]
    best = 0
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            E1[i][j] = max(H[i][j - 1] - q - e, E1[i][j - 1] - e)
            E2[i][j] = max(H[i][j - 1] - q2 - e2, E2[i][j - 1] - e2)
            F1[i][j] = max(H[i - 1][j] - q - e, F1[i - 1][j] - e)
            F2[i][j] = max(H[i - 1][j] - q2 - e2, F2[i - 1][j] - e2)
            cp, ct = code.get(P[i - 1], 4), code.get(T[j - 1], 4)
            s = -1 if 4 in (cp, ct) else (a if cp == ct else -b)
            H[i][j] = max(0, H[i - 1][j - 1] + s, E1[i][j], E2[i][j], F1[i][j], F2[i][j])
            best = max(best, H[i][j])
    return best


def path_score(r, P, T, costs):
    """Score of the reported path, recomputed from its CIGAR and the two sequences (an N column costs 1, not b)."""
    a, b, q, e, q2, e2 = costs
    s, i, j = 0, r["p_start"], r["t_start"]
    for ln, op in r["cigar"]:
        if op in (7, 8):
            for k in range(ln):
                amb = P[i + k] not in b"ACGT" or T[j + k] not in b"ACGT"
                assert (op == 7) == (not amb and P[i + k] == T[j + k])
                s += a if op == 7 else (-1 if amb else -b)
            i, j = i + ln, j + ln
        else:
            s -= min(q + ln * e, q2 + ln * e2)
            i, j = (i + ln, j) if op == 1 else (i, j + ln)
    assert (i, j) == (r["p_end"], r["t_end"])
    return s


@pytest.mark.parametrize("costs", [ou.COSTS_MAP_HIFI, ou.COSTS_ALLELE_SCORING])
def test_affine_oracle_equals_independent_dp(costs):
    rng = np.random.default_rng(3)
    aff = ou.AffineOracle(costs)
    for case in range(120):
        n = int(rng.integers(1, 60))
        T = bytes(rng.choice(list(b"ACGT"), n).tolist())
        if case % 3 == 0:
            P = bytes(rng.choice(list(b"ACGTN"), int(rng.integers(1, 40))).tolist())
        else:  # a mutated copy of a piece of T: the regime the path works in
            lo = int(rng.integers(0, n))
            P = bytes(synth.mutate(rng, np.frombuffer(T[lo:lo + int(rng.integers(1, 40))], dtype=np.uint8), int(rng.integers(0, 3)), int(rng.integers(0, 2))).tolist())
            if not P:
                P = b"A"
        r = aff.align(P, T)
        assert r["score"] == py_affine_local_score(P, T, costs), (P, T)
        if r["score"] > 0:
            assert path_score(r, P, T, costs) == r["score"]  # the reported path carries the optimum
            assert sum(ln for ln, op in r["cigar"] if op in (7, 8, 1)) == r["p_end"] - r["p_start"]
            assert sum(ln for ln, op in r["cigar"] if op in (7, 8, 2)) == r["t_end"] - r["t_start"]
            assert r["nm"] == sum(ln for ln, op in r["cigar"] if op != 7)
            assert r["dist"] == r["nm"] + len(P) - (r["p_end"] - r["p_start"])
        else:
            assert r["cigar"] == [] and r["dist"] == len(P)


def test_affine_oracle_hand_cases():
    T = b"TTTT" + b"ACGTTGCAAGCTTCGGATCCATGGTACCGAGCTCGAATTCACTGGCCGTCGTTTTACAACG" + b"GGGG"
    core = T[4:-4]
    hifi, a5 = ou.AffineOracle(ou.COSTS_MAP_HIFI), ou.AffineOracle(ou.COSTS_ALLELE_SCORING)
    r = hifi.align(core, T)
    assert (r["score"], r["dist"], r["cigar"], r["t_start"], r["t_end"]) == (len(core), 0, [(len(core), 7)], 4, 4 + len(core))
    # a mismatch two bases before the pattern end: a = 1 clips three bases (1 - 4 < 0), a = 5 keeps them (src/hla/caller.rs:1381-1387)
    bad = bytearray(core)
    bad[-3] = ord("A") if bad[-3] != ord("A") else ord("C")
    r1, r5 = hifi.align(bytes(bad), T), a5.align(bytes(bad), T)
    assert (r1["nm"], r1["p_end"], r1["dist"]) == (0, len(core) - 3, 3)
    assert (r5["nm"], r5["p_end"], r5["dist"]) == (1, len(core), 1)
    # two-piece gap: a 30-base deletion costs min(6 + 60, 26 + 30) = 56 and is one D run
    rng = np.random.default_rng(1)
    big = bytes(rng.choice(list(b"ACGT"), 400).tolist())
    pat = big[:185] + big[215:]
    r = hifi.align(pat, big)
    assert r["cigar"] == [(185, 7), (30, 2), (185, 7)] and r["score"] == 370 - 56 and r["nm"] == 30


def test_call_level_divergence_is_held():
    """Seeded product-vs-model comparison on the regimes of the bench workload; the numbers mirror DESIGN.md §3."""
    orc = ou.Oracle()
    g = json.loads((ROOT / "tests/golden/hla_faux.json").read_text())
    rows = [(k, v["gene_name"], v["star_allele"], v["dna_sequence"], v["cdna_sequence"]) for k, v in g["hla_sequences"].items()]
    dna, cdna = rows[0][3].encode(), rows[0][4].encode()
    snp = bytearray(dna)
    snp[1500] = ord("A") if snp[1500] != ord("A") else ord("C")
    aff5 = am.AffineFlowOracle(orc, ou.COSTS_ALLELE_SCORING)
    res = am.score_read_flips(orc, aff5, rows, "HLA-A", [(dna, cdna), (bytes(snp), cdna), (b"ACGT", b"N")])
    assert res["best_allele_differs"] == 0 and res["best_stats_differ"] == 0 and res["floor_disagrees"] == 0
    # a synthetic 40-allele gene: score_read on two consensus-like targets, the allele-pair call of one het sample
    alleles, reads, src, cd = synth.hla_gene(9, "HLA-A", n_alleles=40, n_reads=6, with_cdna=True)
    db = [(f"HLA:X{a:04d}", "HLA-A", ["01", f"{a:02d}"], alleles[a].decode(), cd[a].decode()) for a in range(40)]
    targets = [(reads[k], cd[int(src[k])]) for k in range(2)]
    res = am.score_read_flips(orc, aff5, db, "HLA-A", targets)
    assert res["best_allele_differs"] == 0, res
    sample = [(f"r{k}", reads[k], cd[int(src[k])]) for k in range(6)]
    res = am.diplotype_flips(orc, aff5, db, "HLA-A", [sample])
    assert res["pair_differs"] == 0 and res["hom_het_differs"] == 0, res
    # CYP2D6: weight_sequence decisions of a few reads of the diploid sample
    c = synth.cyp2d6_diploid_sample(2001, n_reads=6)
    labels = fo.labels_from_rows(c["regions"])
    segs = [r[2] for q in sorted(c["roi"])[:3] for r in c["roi"][q]][:6]
    aff1 = am.AffineFlowOracle(orc, ou.COSTS_MAP_HIFI)
    res = am.weight_sequence_flips(orc, aff1, segs, c["consensuses"], labels)
    assert res["emptiness_differs"] == 0 and res["argmin_set_differs"] == 0, res
    assert res["cut_35pct_hides_model_hit"] == 0 and res["cut_35pct_keeps_model_miss"] == 0, res
