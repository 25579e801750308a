"""CPU suite for the C++ host (pb_starphase_b200/host, pybind11 module _starphase_host): the pure host logic is held
against the known-answer vectors of the reference's own unit tests (same vectors as test_host_logic_cpu.py, which
pins the Python oracle) and against the Python oracle on seeded random inputs.  Nothing here needs a GPU; the
GPU-backed functions are covered by test_host_cpp_gpu.py."""
import math
import random
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import starphase_oracle as so  # noqa: E402


@pytest.fixture(scope="module")
def host():
    from pb_starphase_b200 import build

    build.build()
    build.build_host()
    from pb_starphase_b200 import _starphase_host

    return _starphase_host


CIGAR = [(2, 7), (1, 8), (2, 7), (1, 1), (2, 7), (1, 2), (2, 7)]  # ==X==I==D==


def test_process_mm_cigar(host):  # src/hla/processed_match.rs:269-328
    assert host.process_mm_cigar(CIGAR, 0, 10, 0, 0) == [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 3]
    assert host.process_mm_cigar(CIGAR, 3, 18, 2, 3) == [0, 0, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5, 5, 6, 7, 8, 8, 8]
    assert host.process_mm_cigar([(2, 7)], 2, 4, 100, 0) == [0, 1, 2, 2, 2]
    assert host.process_mm_cigar([(2, 7)], 0, 4, 0, 100) == [0, 0, 0, 1, 2]
    with pytest.raises(host.HostError, match="Unexpected cigar type: 0"):
        host.process_mm_cigar([(4, 0)], 0, 4, 0, 0)


def test_process_mm_cigar_random_vs_oracle(host):
    rnd = random.Random(3)
    for _ in range(300):
        cig, t_used = [], 0
        for _ in range(rnd.randint(0, 8)):
            op = rnd.choice([1, 2, 7, 8])
            if cig and cig[-1][1] == op:
                continue
            ln = rnd.randint(1, 6)
            cig.append((ln, op))
            t_used += ln if op != 1 else 0
        off = rnd.randint(0, 10)
        tlen = off + t_used + rnd.randint(0, 10)
        cs, ce = rnd.randint(0, 15), rnd.randint(0, 15)
        assert host.process_mm_cigar(cig, off, tlen, cs, ce) == so.process_mm_cigar(cig, off, tlen, cs, ce)


def test_mapping_scores(host):  # src/data_types/mapping.rs:211-241, src/hla/mapping.rs:181-229
    assert host.MappingStats(10, 1, 0).mapping_score() == 0.1
    assert host.MappingStats(10, 0, 0).mapping_score() == 0.01
    assert host.MappingStats(10, 1, 5).custom_score(False) == 1 / 5
    s = host.HlaMappingStats()
    assert s.mapping_score() == (1.0, 1.0)
    s.cdna_stats, s.dna_stats = host.MappingStats(10, 1, 0), host.MappingStats(20, 0, 1)
    assert s.mapping_score() == (0.1, 0.05)
    rnd = random.Random(1)
    for _ in range(200):
        l, nm, um = rnd.randint(1, 5000), rnd.randint(0, 50), 0
        um = rnd.randint(0, l - 1)
        for pen in (True, False):
            assert host.MappingStats(l, nm, um).custom_score(pen) == so.MappingStats(l, nm, um).custom_score(pen)


def test_select_best_mapping(host):  # src/util/mapping.rs:22-57
    m1 = host.Mapping(0, 90, 100, 10, 100, 200, nm=2)
    m2 = host.Mapping(0, 100, 100, 0, 100, 200, nm=3)
    assert host.select_best_mapping([m1, m2], False, True) == (1, (100, 3, 0))
    assert host.select_best_mapping([m1, m2], True, False) == (0, (200, 2, 110))
    assert host.select_best_mapping([], False, True) == (None, (1, 1, 0))
    assert host.select_best_mapping([m2, m2], False, True)[0] == 0
    assert host.select_best_mapping([m1], False, True, 400) == (0, (400, 2, 310))


def _pm(host, hla_id, cdna, dna):
    pm = host.HlaProcessedMatch(hla_id)
    pm.add_mapping(cdna)
    pm.add_mapping(dna)
    return pm


def test_is_better_match_overlap_rule(host):  # src/hla/processed_match.rs:103-184
    full = host.Mapping(0, 10, 10, 0, 10, 10, nm=0, cigar=[(10, 7)])
    one_x = host.Mapping(0, 10, 10, 0, 10, 10, nm=1, cigar=[(4, 7), (1, 8), (5, 7)])
    a, b = _pm(host, "A", full, full), _pm(host, "B", one_x, full)
    assert a.is_better_match(b) and not b.is_better_match(a)
    worst = host.HlaProcessedMatch.worst_match(2)
    assert a.is_better_match(worst) and not worst.is_better_match(a)
    assert not _pm(host, "N", None, None).is_better_match(worst)
    short = host.Mapping(0, 6, 6, 2, 8, 10, nm=0, cigar=[(6, 7)])
    assert _pm(host, "C", short, None).is_better_match(_pm(host, "D", one_x, None))
    rev = host.Mapping(0, 10, 10, 0, 10, 10, nm=0, forward=False, cigar=[(10, 7)])
    with pytest.raises(host.HostError, match="Reverse strand"):
        host.HlaProcessedMatch("R").add_mapping(rev)


def test_is_better_match_random_vs_oracle(host):
    rnd = random.Random(9)

    def rand_mapping(tlen):
        qlen = rnd.randint(5, 30)
        ts = rnd.randint(0, tlen - 1)
        cig, t, q, nm = [], ts, 0, 0
        while t < tlen and q < qlen and rnd.random() < 0.9:
            op = rnd.choice([7, 7, 7, 8, 1, 2])
            if cig and cig[-1][1] == op:
                continue
            ln = rnd.randint(1, 4)
            if op in (7, 8, 2):
                ln = min(ln, tlen - t)
            if op in (7, 8, 1):
                ln = min(ln, qlen - q)
            if ln == 0:
                break
            cig.append((ln, op))
            t += ln if op != 1 else 0
            q += ln if op != 2 else 0
            nm += ln if op != 7 else 0
        qs = rnd.randint(0, qlen - q)
        return dict(query_start=qs, query_end=qs + q, query_len=qlen, target_start=ts, target_end=t, target_len=tlen, nm=nm, cigar=cig)

    for _ in range(300):
        tl_c, tl_d = rnd.randint(10, 40), rnd.randint(10, 40)
        pair = []
        for _side in range(2):
            kc = rand_mapping(tl_c) if rnd.random() < 0.85 else None
            kd = rand_mapping(tl_d) if rnd.random() < 0.85 else None
            cpp = _pm(host, "x", host.Mapping(**kc) if kc else None, host.Mapping(**kd) if kd else None)
            py = so.HlaProcessedMatch("x")
            for k in (kc, kd):
                py.add_mapping(so.Mapping(k["query_start"], k["query_end"], k["query_len"], k["target_start"], k["target_end"],
                                          k["target_len"], k["nm"], True, k["cigar"]) if k else None)
            assert cpp.processed_cigars() == py.processed_cigars and [tuple(r) for r in cpp.processed_ranges()] == py.processed_ranges
            pair.append((cpp, py))
        (c1, p1), (c2, p2) = pair
        assert c1.is_better_match(c2) == p1.is_better_match(p2) and c2.is_better_match(c1) == p2.is_better_match(p1)


def test_is_passing_dual(host):  # src/hla/caller.rs:1836-1845
    kw = dict(min_consensus_fraction=0.10, min_cdf=0.001, expected_maf=0.5)
    assert not host.is_passing_dual(3, 20, **kw) and not host.is_passing_dual(20, 3, **kw)
    assert host.is_passing_dual(10, 20, **kw) and host.is_passing_dual(20, 10, **kw)
    for c1 in range(0, 70, 3):
        for c2 in range(1, 70, 5):
            assert host.is_passing_dual(c1, c2) == so.is_passing_dual(c1, c2), (c1, c2)
            assert host.binomial_cdf(c1 + c2, 0.45, min(c1, c2)) == pytest.approx(so.binomial_cdf(c1 + c2, 0.45, min(c1, c2)), rel=1e-9)


def test_statistics(host):  # src/util/stats.rs:45-70
    assert host.multinomial_ln_pmf([1.0], [10]) == pytest.approx(0.0, abs=1e-9)
    assert host.multinomial_ln_pmf([0.25, 0.75], [1, 3]) == pytest.approx(math.log(4.0 * 0.25 * 0.75 ** 3), abs=1e-6)
    assert host.multinomial_ln_pmf([0.25, 0.25, 0.5], [2, 2, 0]) == pytest.approx(math.log(6.0 * 0.25 ** 4), abs=1e-6)
    rnd = random.Random(2)
    for _ in range(100):
        k = rnd.randint(1, 5)
        raw = [rnd.random() + 0.01 for _ in range(k)]
        probs = [x / sum(raw) for x in raw]
        obs = [rnd.randint(0, 400) for _ in range(k)]
        if sum(obs) == 0:
            continue
        assert host.multinomial_ln_pmf(probs, obs) == so.multinomial_ln_pmf(probs, obs)  # same operations in the same order
    for x in (0, 1, 5, 170, 171, 500, 10000):
        assert host.ln_factorial(x) == so.ln_factorial(x)


TYPES = [so.UNKNOWN, so.REP6, so.CYP2D6, so.LINK, so.REP7, so.SPACER, so.CYP2D7, so.DELETION, so.HYBRID, so.FALSE_ALLELE]
SUBS = [None, "4.001", "10", "1e2", "abc", "CYP2D6::CYP2D7::exon2", "CYP2D7::CYP2D6::intron1", "nan", "-3.5", ".5", "5."]


def test_region_labels_vs_oracle(host):  # src/cyp2d6/region_label.rs
    tr = so.Cyp2d6Config.default().cyp_translate
    for t1 in TYPES:
        for s1 in SUBS:
            a = so.RegionLabel(t1, s1)
            for t2 in TYPES:
                b = so.RegionLabel(t2, None)
                for na in (True, False):
                    got = host.label_ops(t1, s1, t2, None, na)
                    assert got == dict(full_allele=a.full_allele(), simple=a.simplify_allele(False, tr), detailed=a.simplify_allele(True, tr),
                                       allowed=a.is_allowed_label(), allowed_pair=a.is_allowed_label_pair(b),
                                       head=a.is_candidate_chain_head(na), normalizing=a.is_normalizing_allele(na)), (t1, s1, t2, na)


def test_convert_chain_to_hap(host):  # src/cyp2d6/caller.rs:971-1006
    rows = [("CYP2D7", None, 0), ("CYP2D6", "1.001", 1), ("CYP2D6", "10", 2), ("CYP2D6", "1.002", 3), ("CYP2D6", "1.002", 4)]
    assert host.convert_chain_to_hap([2, 2, 1, 0], rows, "SubAlleles") == "*1.001 + *10x2"
    assert host.convert_chain_to_hap([3, 1, 0], rows, "SubAlleles") == "*1.001 + *1.002"
    assert host.convert_chain_to_hap([3, 1, 0], rows, "CoreAlleles") == "*1x2"
    assert host.convert_chain_to_hap([3, 4], rows, "SubAlleles") == "*1.002x2"
    labels = [so.RegionLabel(t, s) for t, s, _ in rows]
    tr = so.Cyp2d6Config.default().cyp_translate
    assert host.convert_chain_to_hap([3, 1, 0], rows, "DeepAlleles") == so.convert_chain_to_hap([3, 1, 0], labels, so.DEEP, tr, [0, 1, 2, 3, 4])
    rows5 = [("CYP2D6*5", None, None), ("CYP2D6", "4", None)]
    assert host.convert_chain_to_hap([0], rows5, "CoreAlleles") == "*5" and host.convert_chain_to_hap([0, 1], rows5, "CoreAlleles") == "*4"
    assert host.convert_chain_to_hap([0], rows5, "DeepAlleles") == "(X_CYP2D6*5)"


def test_build_chains_vs_oracle(host):  # src/cyp2d6/caller.rs:430-537
    rnd = random.Random(6)
    for trial in range(60):
        n_haps = rnd.randint(2, 6)
        rw = {}
        for r in range(rnd.randint(1, 12)):
            regions = []
            for _ in range(rnd.randint(0, 4)):
                if rnd.random() < 0.15:
                    regions.append([])
                else:
                    regions.append([(rnd.randint(0, 3), rnd.random()) for _ in range(n_haps)])
            rw[f"read_{r:02d}"] = regions
        try:
            want = so.build_chains(rw, n_haps)
        except RuntimeError:
            with pytest.raises(host.HostError, match="chain collapse"):
                host.build_chains(rw, n_haps)
            continue
        chains, scores, counts = host.build_chains(rw, n_haps)
        assert chains == want[0] and counts == want[2]
        assert {k: [[tuple(x) for x in seg] for seg in v] for k, v in scores.items()} == {k: [[tuple(x) for x in seg] for seg in v] for k, v in want[1].items()}


def test_json_writer(host):  # src/util/file_io.rs:37-52, src/data_types/starphase_json.rs:13-21
    s = host.HlaMappingStats()
    s.dna_stats = host.MappingStats(3502, 2, 0)
    meta = dict(pbstarphase_version="2.0.1-test", cpic_version="v", hla_version="3.62.0", pharmvar_version="6", build_time="t\n\"x\"")
    text = host.starphase_json("2.0.1-test", meta, {"HLA-B": s.to_json(), "HLA-A": s.to_json()})
    want = so.starphase_json("2.0.1-test", meta, {"HLA-A": so.HlaMappingStats(None, so.MappingStats(3502, 2, 0)).to_json(),
                                                  "HLA-B": so.HlaMappingStats(None, so.MappingStats(3502, 2, 0)).to_json()})
    assert text == want
    with pytest.raises(host.HostError, match="lacks"):
        host.starphase_json("v", {}, {})


def test_no_gpu_is_an_error(host):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(host.HostError, match="no CPU fallback"):
        host.GpuAligner(0)


def test_dp_score(host):
    assert host.dp_score([(40, 7)]) == 200 and host.dp_score([(10, 7), (1, 8), (10, 7)]) == 96
    assert host.dp_score([(50, 7), (1, 1), (50, 7)]) == 500 - 8 and host.dp_score([(50, 7), (30, 2), (50, 7)]) == 500 - 56


# ---- hla_debug.json pieces, consensus preparation, hemizygous test (pb_starphase_b200/host/sp_host_debug.cpp) ----
def test_json_f64_matches_ryu_layout(host):
    known = {0.1: "0.1", 1e-5: "0.00001", 1e-6: "1e-6", 1.0: "1.0", 1e15: "1000000000000000.0", 1e16: "1e16", 0.45: "0.45",
             123456.789: "123456.789", 1.5e-7: "1.5e-7", 1 / 3: "0.3333333333333333", 1.2345678901234568e17: "1.2345678901234568e17",
             5e-324: "5e-324", -2.5: "-2.5", 0.0: "0.0", float("nan"): "null", float("inf"): "null"}
    for v, text in known.items():
        assert host.json_f64(v) == text == so.format_f64(v)
    rnd = random.Random(11)
    import struct

    for _ in range(20000):
        v = struct.unpack("d", struct.pack("Q", rnd.getrandbits(64)))[0] if rnd.random() < 0.5 else rnd.uniform(-10, 10) * 10 ** rnd.randint(-8, 20)
        assert host.json_f64(v) == so.format_f64(v)


def test_is_hemizygous_better_reference_vectors(host):  # src/hla/caller.rs:1847-1899
    def run(c1, c2, nc, delta):
        ic = [True] * c1 + [False] * c2
        s1, s2 = [0] * c1 + [delta] * c2, [delta] * c1 + [0] * c2
        a = host.is_hemizygous_better(s1, s2, ic, c2 != 0, 20, nc)
        assert a == so.is_hemizygous_better(s1, s2, ic, c2 != 0, 20, nc)
        return a

    assert run(20, 0, 20.0, 1)
    assert not run(40, 0, 20.0, 1)
    assert run(18, 2, 20.0, 1)
    assert not run(18, 17, 20.0, 1)
    assert not run(15, 6, 20.0, 20)
    # unscored reads take the other score + dual_max_ed_delta (:1597-1599); no coverage -> read count
    rnd = random.Random(5)
    for _ in range(300):
        n = rnd.randint(1, 60)
        s1 = [rnd.choice([None, rnd.randint(0, 40)]) for _ in range(n)]
        s2 = [rnd.randint(0, 40) if a is None or rnd.random() < 0.8 else None for a in s1]
        ic = [rnd.random() < 0.6 for _ in range(n)]
        nc = rnd.choice([None, rnd.uniform(5, 60)])
        dual = rnd.random() < 0.8
        assert host.is_hemizygous_better(s1, s2, ic, dual, 20, nc) == so.is_hemizygous_better(s1, s2, ic, dual, 20, nc)
    with pytest.raises(host.HostError, match="Bad distribution parameters"):
        host.is_hemizygous_better([], [], [], False, 20, None)


def test_dual_passing_stats(host):  # src/hla/caller.rs:1225-1247, tests :1837-1845
    import json

    for c1, c2, passing in ((3, 20, False), (20, 3, False), (10, 20, True), (20, 10, True)):
        got = json.loads(host.dual_passing_stats_json(True, c1, c2, expected_maf=0.5))  # the settings of run_passing_test (:1812-1817)
        want = so.dual_passing_stats(True, c1, c2, expected_maf=0.5)
        assert got["is_passing"] is passing is want["is_passing"]
        assert (got["counts1"], got["counts2"], got["maf"], got["is_dual"]) == (c1, c2, want["maf"], True)
        assert got["cdf"] == want["cdf"]                                 # both restate statrs' beta_reg
        assert got["cdf"] == pytest.approx(so.binomial_cdf_exact(c1 + c2, 0.5, min(c1, c2)), rel=1e-12)
    assert host.dual_passing_stats_json(False, 0, 0) == so.serde_pretty(so.dual_passing_stats(False))
    # the reference's documented example (docs/debug_outputs.md:128-135), byte for byte
    assert host.dual_passing_stats_json(True, 27, 10) == so.serde_pretty(so.dual_passing_stats(True, 27, 10))
    assert '"cdf": 0.019406414321609413' in host.dual_passing_stats_json(True, 27, 10)
    rnd = random.Random(9)
    for _ in range(150):  # beta_reg (continued fraction) against the exact sum
        n = rnd.randint(1, 200)
        k = rnd.randint(0, n)
        p = rnd.choice([0.45, 0.5, rnd.uniform(0.01, 0.99)])
        assert host.binomial_cdf(n, p, k) == so.binomial_cdf(n, p, k)
        assert host.binomial_cdf(n, p, k) == pytest.approx(so.binomial_cdf_exact(n, p, k), rel=1e-10, abs=1e-300)


def test_reverse_complement(host):  # src/util/sequence.rs:25-36
    assert host.reverse_complement("ACCGGGTN") == "NACCCGGT" == so.reverse_complement(b"ACCGGGTN").decode()
    with pytest.raises(host.HostError, match="Unexpected character for reverse-complement: 97"):
        host.reverse_complement("ACa")


def test_cigar_and_md_strings(host):
    #            target  ACGTACGTAC  query ACGAACTTTGTAC: 3= 1X 2= 3I(TTT) 1= ... built by hand below
    target, query = "GGACGTACGTACGG", "ACGAACTTTGAC"
    cigar = [(3, 7), (1, 8), (2, 7), (3, 1), (1, 7), (1, 2), (2, 7)]  # ACG A AC [TTT] G ^T AC
    assert host.cigar_string(cigar) == "3=1X2=3I1=1D2=" == so.cigar_string(cigar)
    assert host.md_string(cigar, target, 2, query, 0) == "3T3^T2" == so.md_string(cigar, target.encode(), 2, query.encode(), 0)
    # docs/debug_outputs.md:96-118: an exact match of a 1,098 bp cDNA inside a 1,535 bp consensus reads "1098=" / "1098"
    q = "".join(random.Random(1).choice("ACGT") for _ in range(1098))
    assert host.cigar_string([(1098, 7)]) == "1098=" and host.md_string([(1098, 7)], "T" * 200 + q + "A" * 237, 200, q, 0) == "1098"
    assert host.md_string([(4, 7)], "ACGT", 0, "ACGT", 0) == "4"
    assert host.md_string([(1, 8), (3, 7)], "ACGT", 0, "TCGT", 0) == "0A3"
    assert host.md_string([(3, 7), (1, 8)], "ACGT", 0, "ACGA", 0) == "3T"       # no trailing zero (write_MD_core)
    assert host.md_string([(2, 2), (2, 7)], "ACGT", 0, "GT", 0) == "0^AC2"
    rnd = random.Random(4)
    for _ in range(300):
        t = "".join(rnd.choice("ACGTN") for _ in range(60))
        cig, ti, q = [], 5, []
        for _ in range(rnd.randint(1, 8)):
            op = rnd.choice([7, 8, 1, 2])
            ln = rnd.randint(1, 5)
            if ti + ln > 55:
                break
            if op == 7:
                q.append(t[ti:ti + ln]); ti += ln
            elif op == 8:
                q.append("".join(rnd.choice([c for c in "ACGT" if c != x]) for x in t[ti:ti + ln])); ti += ln
            elif op == 1:
                q.append("".join(rnd.choice("ACGT") for _ in range(ln)))
            else:
                ti += ln
            cig.append((ln, op))
        qs = "xx" + "".join(q)
        assert host.md_string(cig, t, 5, qs, 2) == so.md_string(cig, t.encode(), 5, qs.encode(), 2)


def test_splice_read(host):  # src/hla/caller.rs:1518-1576
    seq = "AAAACCCCGGGGTTTTACGTACGT"
    # read starts at reference 100: 4 soft-clipped, 8 aligned (100..108), 2 inserted, 3 deleted (108..111), 10 aligned (111..121)
    cigar = [(4, 4), (8, 0), (2, 1), (3, 2), (10, 7)]
    exons = [(90, 104), (106, 112), (118, 130), (200, 210)]
    got = host.splice_read(seq, 100, cigar, exons)
    want = so.splice_read(seq.encode(), 100, cigar, exons)
    assert got == (want[0].decode(), want[1])
    # exon 1 covers reference 100..103 -> read 4..7; exon 2 covers 106, 107 and 111 (108..110 deleted) -> read 10..14 (spans the insertion);
    # exon 3 covers 118..120 -> read 21..23; exon 4 is not covered; offset = 100 - 90
    assert got == (seq[4:8] + seq[10:15] + seq[21:24], 10)
    assert host.splice_read(seq, 100, cigar, [(300, 310)]) == ("", 10)
    s = host.DiplotypeSettings()
    assert host.prepare_score_read_targets(seq, 100, cigar, exons, True, s) == (seq, got[0])
    assert host.prepare_score_read_targets(seq, 100, cigar, exons, False, s) == (host.reverse_complement(seq), host.reverse_complement(got[0]))
    assert host.prepare_score_read_targets(seq, 100, cigar, [(300, 310)], False, s)[1] == "N"
    s.disable_cdna_scoring = True
    assert host.prepare_score_read_targets(seq, 100, cigar, exons, True, s)[1] == "N"
    rnd = random.Random(8)
    for _ in range(200):
        cig, qlen = [], 0
        for _ in range(rnd.randint(1, 7)):
            op = rnd.choice([0, 1, 2, 3, 4, 7, 8])
            ln = rnd.randint(1, 9)
            cig.append((ln, op))
            qlen += ln if op in (0, 1, 4, 7, 8) else 0
        sq = "".join(rnd.choice("ACGT") for _ in range(qlen))
        ex, p = [], 40
        for _ in range(rnd.randint(1, 4)):
            p += rnd.randint(0, 12)
            e = p + rnd.randint(1, 15)
            ex.append((p, e))
            p = e
        w = so.splice_read(sq.encode(), 50, cig, ex)
        assert host.splice_read(sq, 50, cig, ex) == (w[0].decode(), w[1])


def test_alleles_from_traversal(host):  # src/cyp2d6/haplotyper.rs:454-468
    node_to_alleles = {0: [(0, 0), (1, 1)], 2: [(1, 1), (2, 0)], 5: [(2, 1), (3, 1)], 9: [(0, 1)]}
    for nodes in ([0, 2, 5], [5, 2, 0], [1, 3], [0, 9, 9], []):
        assert host.alleles_from_traversal(5, nodes, node_to_alleles) == so.alleles_from_traversal(5, nodes, node_to_alleles)
    assert host.alleles_from_traversal(5, [0, 2, 5], node_to_alleles) == [0, 1, 2, 1, 3]
    assert host.alleles_from_traversal(5, [0, 9], node_to_alleles) == [2, 1, 3, 3, 3]


def test_variant_match_oracle_known_answers():  # the match rule of src/cyp2d6/haplotyper.rs:486-506
    alleles, hap, vi = [0, 1, 2, 3, 1, 0], [0, 1, 1, 1, 0, 1], [True, False, True, True, True, False]
    assert so.variant_match(alleles, hap, vi) == (2, 3)      # sites 0, 1, 2 agree; of those 0 and 2 are VI
    with pytest.raises(ValueError, match="Unexpected seq_value=4"):
        so.variant_match([4], [0], [True])
    star, rv, score = so.assign_haplotype_from_alleles(alleles, {"1": [1, 0, 0, 0, 0, 1], "2": hap}, list("abcdef"), vi, False)
    assert (star, score) == ("2", (2, 3))
    # an all-REF definition ties at (2, 3) (sites 0, 2, 5): Unknown unless forced, then the first by full_allele
    assert so.assign_haplotype_from_alleles(alleles, {"1": [0] * 6, "2": hap}, list("abcdef"), vi, False)[0] is None
    assert so.assign_haplotype_from_alleles(alleles, {"1": [0] * 6, "2": hap}, list("abcdef"), vi, True)[0] == "1"
    assert [(v["label"], v["variant_state"]) for v in rv] == [("b", "Match"), ("c", "AmbiguousMissing"), ("d", "UnknownMissing"),
                                                             ("e", "Unexpected"), ("f", "Missing")]


def test_hpc_reference_vectors(host):  # src/util/homopolymers.rs:71-104, src/hla/realigner.rs:534-556
    assert host.hpc("AACAAAAAAGGGTAACAA") == "ACAGTACA" == so.hpc(b"AACAAAAAAGGGTAACAA").decode()
    seq = "AACCCGTTTT"
    for i, c in enumerate(seq):
        assert host.hpc_pos(seq, i) == "ACGT".index(c) == so.hpc_pos(seq.encode(), i)
    assert host.hpc_pos("ATTGGGGGAACCCGTTTT", 6) == 2 and host.hpc("GAACCCGTTTT") == "GACGT"   # test_hpc_guide
    assert host.hpc("AACCGGTTAACCGGTTAACCGGTT"[4:10]) == "GTA"                                    # test_realigned_record
    assert host.hpc_pos(seq, 100) == 4 == so.hpc_pos(seq.encode(), 100) and host.hpc("") == ""
    assert host.hpc_with_guide("GAACCCGTTTT", "ATTGGGGGAACCCGTTTT", 6) == (b"GACGT", 2)            # test_hpc_guide, :93-101


def test_overlap_score_and_region_variant_display(host):  # src/cyp2d6/haplotyper.rs:935-941, src/data_types/region_variants.rs:80-110
    assert host.overlap_score(0, 1, 1, 2) == 0.0        # no overlap
    assert host.overlap_score(0, 10, 1, 5) == 1.0       # fully contained
    assert host.overlap_score(0, 10, 5, 100) == 0.5     # half shared of first
    assert host.overlap_score(15, 100, 0, 20) == 0.25   # quarter shared of second
    states = ["Unknown", "Match", "Unexpected", "Missing", "AmbiguousUnexpected", "AmbiguousMissing", "UnknownUnexpected", "UnknownMissing"]
    shown = [host.region_variant_string("rs123", True, k) for k in range(len(states))]
    assert shown == ["?rs123", "=rs123", "+rs123", "-rs123", "?rs123", "?rs123", "?rs123", "?rs123"]


def test_cyp2d6_alleles_json_reference_docs_vector(host):
    """docs/debug_outputs.md:27-54: the reference's documented cyp2d6_alleles.json -- the deep / sub-allele / core forms of
    (*41.004 with an unexpected rs28735595) and (*68 + *4.001 with the same extra variant), allele keys {index}_{label}."""
    states = ["Unknown", "Match", "Unexpected", "Missing", "AmbiguousUnexpected", "AmbiguousMissing", "UnknownUnexpected", "UnknownMissing"]
    regions = [("Hybrid", "CYP2D6::CYP2D7::exon2", 0, None), ("CYP2D7", None, 1, None),
               ("CYP2D6", "4.001", 2, [("rs28371738", False, "Match"), ("rs28735595", False, "Unexpected")]),
               ("REP6", None, 3, None),
               ("CYP2D6", "41.004", 4, [("rs16947", True, "Match"), ("rs28735595", False, "Unexpected"), ("rs1", False, "UnknownUnexpected")])]
    best = [[4], [2, 0]]
    got = host.cyp2d6_alleles_json(best, [(t, s, u, None if v is None else [(l, vi, states.index(st)) for l, vi, st in v])
                                          for t, s, u, v in regions])
    doc = __import__("json").loads(got)
    assert doc["hap1"] == {"deep_form": "(4_CYP2D6*41.004 +rs28735595)", "suballele_form": "*41.004", "core_form": "*41"}
    assert doc["hap2"] == {"deep_form": "(0_CYP2D6::CYP2D7::exon2) + (2_CYP2D6*4.001 +rs28735595)", "suballele_form": "*68 + *4.001",
                           "core_form": "*68 + *4"}
    assert list(doc["alleles"]) == ["2_CYP2D6*4.001", "4_CYP2D6*41.004"]
    assert doc["alleles"]["2_CYP2D6*4.001"][0] == {"label": "rs28371738", "is_vi": False, "variant_state": "Match"}
    labels = [so.RegionLabel(t, s) for t, s, _, _ in regions]
    variants = [None if v is None else [dict(label=l, is_vi=vi, variant_state=st) for l, vi, st in v] for _, _, _, v in regions]
    want = so.cyp2d6_alleles_json(best, labels, [u for _, _, u, _ in regions], variants, so.Cyp2d6Config.default().cyp_translate)
    assert got == want
    # the other deltas of deep_label (src/cyp2d6/region.rs:60-95)
    assert so.deep_label(so.RegionLabel("CYP2D6", "1.001"), None, [dict(label="a", variant_state="Missing"), dict(label="b", variant_state="AmbiguousMissing"),
                                                                 dict(label="c", variant_state="Match")]) == "X_CYP2D6*1.001 -a ?b"


def test_diplotype_strings_reference_vectors(host):  # src/data_types/pgx_diplotype.rs:236-256
    assert host.diplotype_strings("B", "A")[0] == "B/A"
    assert host.diplotype_strings("*4", "*1")[1] == "*4/*1"
    assert host.diplotype_strings("*4x2", "*1")[1] == "*4x2/*1"
    assert host.diplotype_strings("*4 + *68", "*1")[1] == "[*4 + *68]/*1"
    assert host.diplotype_strings("*68 + *4.001", "*4.001")[2] == so.serde_pretty(so.diplotype_json("*68 + *4.001", "*4.001"))


def test_harmonic_mean_reference_vector(host):  # src/data_types/mapping.rs:230-241
    assert host.harmonic_mean([0.2, 0.4, 0.2]) == 3.0 / (5.0 + 2.5 + 5.0) == so.harmonic_mean([0.2, 0.4, 0.2])
    assert host.harmonic_mean([]) == 0.0
    with pytest.raises(host.HostError, match="dna_score must be > 0.0"):
        host.harmonic_mean([0.1, 0.0])


def test_is_allowed_allele_def(host):  # src/hla/caller.rs:1673-1701
    assert host.is_allowed_allele_def("HLA-A", True, "HLA-A", True)        # base case
    assert not host.is_allowed_allele_def("HLA-B", True, "HLA-A", True)    # wrong gene
    assert not host.is_allowed_allele_def("HLA-A", False, "HLA-A", True)   # no DNA while DNA is required
    assert host.is_allowed_allele_def("HLA-A", False, "HLA-A", False)      # requirement lifted


def test_score_min(host):  # src/data_types/mapping.rs:220-228, src/hla/mapping.rs:221-229
    def hla(cdna, dna):
        h = host.HlaMappingStats()
        h.cdna_stats, h.dna_stats = cdna, dna
        return h.mapping_score()

    st = lambda nm: host.MappingStats(10, nm, 0)  # noqa: E731
    assert [st(n).mapping_score() for n in (10, 9, 2)] == [1.0, 0.9, 0.2]
    s1, s2, s3 = hla(None, st(5)), hla(st(9), None), hla(None, st(2))   # (1.0, 0.5), (0.9, 1.0), (1.0, 0.2): a missing side scores 1.0
    assert (s1, s2, s3) == ((1.0, 0.5), (0.9, 1.0), (1.0, 0.2))
    assert min(s1, s2) == s2 and min(s1, s3) == s3 and min(s2, s3) == s2  # cDNA first, then DNA
