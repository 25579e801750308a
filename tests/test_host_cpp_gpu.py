"""GPU parity tests for the C++ host above the C ABI (pb_starphase_b200/host): the reference's operator interface
for the hot path -- score_read, HlaRealigner, the allele-pair diplotype, weight_sequence, find_best_chain_pair, the
CYP2D6 chain call and the result JSON -- driven through the pybind11 module, against the same flows restated on the
CPU oracle (tests/flow_oracle.py).  Integers identical, calls identical, JSON byte-identical."""
import json
import math
import random

import numpy as np
import pytest

import flow_oracle as fo
from flow_oracle import so
from test_k3_gpu import noisy, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from pb_starphase_b200 import _starphase_host

    return _starphase_host


@pytest.fixture(scope="module")
def gpu(host):
    return host.GpuAligner(0)


def hla_db(seed=5, n_alleles=14, n_reads=10, genes=("HLA-A", "HLA-B")):
    """Small two-gene database + heterozygous reads per gene (two source alleles), DNA + cDNA."""
    from pb_starphase_b200 import synth

    rows, reads = [], {}
    for g, gene in enumerate(genes):
        alleles, _, _, cdna = synth.hla_gene(seed + g, gene, n_alleles=n_alleles, n_reads=2, with_cdna=True)
        rng = np.random.default_rng([seed, g])
        for a in range(n_alleles):
            star = [f"{1 + a // 4:02d}", f"{1 + a % 4:02d}", "01", f"{a + 1:02d}"]
            rows.append((f"HLA:HLA{g * 1000 + a:05d}", gene, star, alleles[a].decode(), cdna[a].decode()))
        i, j = 2, 9
        rd, src = synth.hifi_reads(rng, [alleles[i], alleles[j]], n_reads)
        reads[gene] = [(f"read_{gene}_{k:02d}", rd[k], cdna[(i, j)[int(src[k])]]) for k in range(n_reads)]
    return rows, reads


def cpp_reads(reads):
    return [(q, d.decode(), c.decode()) for q, d, c in reads]


# ---- score_read: src/hla/caller.rs:1332-1511; reference tests :1709-1809 -------------------------------------
def test_score_read_reference_alleles(host, gpu):
    """test_reference_alleles (src/hla/caller.rs:1709-1773): an exact copy of an allele scores (len, 0, 0) on both
    sequence types and is the best match."""
    rows, _ = hla_db()
    hla_id, gene, star, dna, cdna = rows[3]
    stats, best_id, best_star = host.score_read(gpu, dna, cdna, rows, gene, host.DiplotypeSettings())
    assert stats[hla_id] == ((len(cdna), 0, 0), (len(dna), 0, 0))
    assert best_id == hla_id and best_star == ":".join(star)
    assert set(stats) == {r[0] for r in rows if r[1] == gene}


def test_score_read_golden_hla_faux(host, gpu):
    """The reference's own fixture (test_data/HLA-faux/database.json, committed as tests/golden/hla_faux.json):
    test_reference_alleles (src/hla/caller.rs:1709-1773) for HLA-A (forward gene) and HLA-B (reverse-strand gene: the read is
    the reverse complement and the host puts it back on the gene strand, :1343-1350), and test_score_bad_read (:1783-1809)."""
    from pathlib import Path

    g = json.loads((Path(__file__).resolve().parent / "golden" / "hla_faux.json").read_text())
    rows = [(hid, d["gene_name"], d["star_allele"], d["dna_sequence"], d["cdna_sequence"]) for hid, d in g["hla_sequences"].items()]
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for exp in g["expected"]["test_reference_alleles"]:
        d = g["hla_sequences"][exp["hla_id"]]
        read = d["dna_sequence"].encode()
        if exp["read_is_revcomp"]:
            read = read.translate(comp)[::-1]             # what the BAM holds for a reverse-strand gene
            read = read.translate(comp)[::-1]             # reverse_complement() of score_read puts it on the gene strand
        stats, best_id, best_star = host.score_read(gpu, read.decode(), d["cdna_sequence"], rows, exp["gene"], host.DiplotypeSettings())
        assert best_id == exp["hla_id"] and best_star == exp["star"]
        assert stats[exp["hla_id"]] == ((len(d["cdna_sequence"]), 0, 0), (len(d["dna_sequence"]), 0, 0))
        assert set(stats) == {exp["hla_id"]}              # only the alleles of that gene are scored
    bad = g["expected"]["test_score_bad_read"]
    s = host.DiplotypeSettings()
    s.disable_cdna_scoring = True
    stats, best_id, _ = host.score_read(gpu, bad["read"], "N", rows, bad["gene"], s)
    assert best_id == "" and all(v == (None, None) for v in stats.values())
    h = host.HlaMappingStats()
    assert h.mapping_score() == (bad["worst_score"], bad["worst_score"])


def test_score_bad_read(host, gpu):
    """test_score_bad_read (src/hla/caller.rs:1783-1809): a 4-bp junk read maps nowhere => every allele is worst and
    there is no best id."""
    rows, _ = hla_db()
    stats, best_id, best_star = host.score_read(gpu, "ACGT", "N", rows, "HLA-A", host.DiplotypeSettings())
    assert all(v == (None, None) for v in stats.values()) and best_id == "" and best_star == ""


def test_score_read_vs_oracle(host, gpu, oracle):
    rows, reads = hla_db()
    for gene in ("HLA-A", "HLA-B"):
        for q, dna_t, cdna_t in reads[gene][:2]:
            got = host.score_read(gpu, dna_t.decode(), cdna_t.decode(), rows, gene, host.DiplotypeSettings())
            want = fo.score_read(oracle, dna_t, cdna_t, rows, gene)
            assert (dict(got[0]), got[1], got[2]) == want
    s = host.DiplotypeSettings()
    s.disable_cdna_scoring = True
    q, dna_t, cdna_t = reads["HLA-A"][0]
    got = host.score_read(gpu, dna_t.decode(), cdna_t.decode(), rows, "HLA-A", s)
    assert (dict(got[0]), got[1], got[2]) == fo.score_read(oracle, dna_t, cdna_t, rows, "HLA-A", disable_cdna=True)


def test_score_consensus_vs_oracle(host, gpu, oracle):
    """score_consensus (src/hla/caller.rs:1259-1319): consensus -> reference mapping (K4) -> soft-clipped CIGAR -> splice_read and the
    strand handling -> score_read, for a forward-strand and a reverse-strand gene, against the same flow on the oracle."""
    rng = np.random.default_rng(29)
    ref_start, exon_rel = 10_000, [(300, 600), (1000, 1300), (2000, 2400)]
    for gene, fwd in (("HLA-A", True), ("HLA-B", False)):
        root = rnd(rng, 3000)                       # hg38 orientation
        hg = [root]
        for k in range(1, 9):                       # substitutions only: the exon coordinates stay put
            a = bytearray(root)
            for pos in rng.choice(len(a), size=4 + k, replace=False):
                a[pos] = ord(rng.choice([c for c in "ACGT" if c != chr(a[pos])]))
            hg.append(bytes(a))
        splice = lambda s_: b"".join(s_[a:b] for a, b in exon_rel)
        rows = []
        for k, a in enumerate(hg):
            dna = a if fwd else so.reverse_complement(a)
            cdna = splice(a) if fwd else so.reverse_complement(splice(a))
            rows.append((f"HLA:HLA{k:05d}", gene, ["01", f"{k + 1:02d}", "01", "01"], dna.decode(), cdna.decode()))
        L, R = rnd(rng, 500), rnd(rng, 500)
        reference = L + root + R
        exons = [(ref_start + 500 + a, ref_start + 500 + b) for a, b in exon_rel]
        cases = [("exact_5", L[-150:] + hg[5] + R[:150]), ("noisy_2", noisy(rng, L[-80:] + hg[2] + R[:60], 6)),
                 ("with_junk_ends", rnd(rng, 40) + hg[7] + rnd(rng, 30)), ("empty", b""), ("junk", rnd(rng, 1500))]
        for name, cons in cases:
            got = host.score_consensus(gpu, reference.decode(), ref_start, cons.decode(), rows, gene, exons, fwd, host.DiplotypeSettings())
            want = fo.score_consensus(oracle, reference, ref_start, cons, rows, gene, exons, fwd)
            assert (dict(got[0]), got[1], got[2]) == want[:3], (gene, name)
            assert got[3] == so.serde_pretty(want[3]), (gene, name)
        got = host.score_consensus(gpu, reference.decode(), ref_start, cases[0][1].decode(), rows, gene, exons, fwd, host.DiplotypeSettings())
        assert got[1] == rows[5][0] and got[0][rows[5][0]] == ((len(rows[5][4]), 0, 0), (len(rows[5][3]), 0, 0))
        assert host.score_consensus(gpu, reference.decode(), ref_start, "", rows, gene, exons, fwd, host.DiplotypeSettings())[:3] == ({}, "", "")
        assert host.score_consensus(gpu, reference.decode(), ref_start, cases[4][1].decode(), rows, gene, exons, fwd, host.DiplotypeSettings())[1] == ""


def test_hla_debug_json_vs_oracle(host, gpu, oracle):
    """hla_debug.json (src/hla/debug.rs): per consensus and allele the DetailedMappingStats of the best cDNA / DNA mapping
    (lengths, match_len, nm, unmapped on both sides, EQX CIGAR string, MD string) + DualPassingStats -- the artefact a
    maintainer diffs against `pbstarphase diplotype --debug-folder`.  Byte-identical to the oracle flow."""
    rows, reads = hla_db()
    for gene, fwd in (("HLA-A", True), ("HLA-B", False)):
        cons = [(f"consensus{k + 1}", d, c) for k, (_, d, c) in enumerate(reads[gene][:2])]
        cons.append(("junk", b"ACGT", b"N"))
        got = host.hla_debug_json(gpu, rows, gene, [(q, d.decode(), c.decode()) for q, d, c in cons], True, 7, 5, host.DiplotypeSettings())
        want = fo.hla_debug_json(oracle, rows, gene, cons, True, 7, 5)
        assert got == want  # byte for byte, floats included (statrs' beta_reg restated on both sides, pinned by docs/debug_outputs.md)
        gj = json.loads(got)
        per = gj["read_mapping_stats"][gene]
        assert list(per) == ["consensus1", "consensus2", "junk"] and per["junk"]["best_match_id"] is None
        d = next(iter(per["consensus1"]["mapping_stats"].values()))["dna_mapping"]
        assert d["query_len"] - d["query_unmapped"] > 0 and d["cigar"] and d["md"] and d["match_len"] <= d["query_len"]
    with pytest.raises(host.HostError, match="is already occupied"):
        host.hla_debug_json(gpu, rows, "HLA-A", [("consensus1", "ACGT", "N"), ("consensus1", "ACGT", "N")], False, 0, 0, host.DiplotypeSettings())


# ---- HlaRealigner: src/hla/realigner.rs:98-211 ------------------------------------------------------------------
def test_realign_records_vs_oracle(host, gpu, oracle):
    rows, reads = hla_db()
    rng = np.random.default_rng(3)
    rs = [(q, d) for gene in reads for q, d, _ in reads[gene][:4]]
    rs.append(("junk", rnd(rng, 2500)))                 # maps nowhere well: REFERENCE / ignored
    rs.append(("empty", b""))                           # record without a sequence (realigner.rs:111-114)
    rs.append(("noisy", noisy(rng, rows[1][3].encode(), 160)))  # > 3 % edits: rejected by max_ed_frac
    rs.append(("half", rows[2][3].encode()[:1200]))     # most of the allele unmapped: rejected by max_unmapped_frac
    got = host.realign_records(gpu, ["HLA-A", "HLA-B"], rows, [(q, s.decode()) for q, s in rs]).pretty()
    want = so.serde_pretty(fo.realign_records(oracle, ["HLA-A", "HLA-B"], rows, rs))
    assert got == want
    back = json.loads(got)
    assert [d["is_ignored"] for d in back[-4:]] == [True, True, True, True] and not any(d["is_ignored"] for d in back[:-4])
    assert back[0]["best_star_allele"].startswith("HLA-A*") and back[-4]["best_hla_id"] == "REFERENCE"


def test_realign_records_full_vs_oracle(host, gpu, oracle):
    """realign_record in full (src/hla/realigner.rs:98-350): the best database allele, then the buffered read segment against the
    gene's hg38 sequence and -- when that mapping does not start before the allele mapping -- the allele against hg38, giving the
    segment range and the DNA / HPC offsets; reverse-strand gene (HLA-B) alleles are flipped into hg38 orientation first."""
    rows, _ = hla_db()
    rng = np.random.default_rng(17)
    fw = {}
    for hid, gene, star, dna, cdna in rows:  # hg38-oriented allele sequences
        fw[hid] = dna.encode() if gene == "HLA-A" else so.reverse_complement(dna.encode())
    root = {g: next(fw[r[0]] for r in rows if r[1] == g) for g in ("HLA-A", "HLA-B")}
    flank = {g: (rnd(rng, 1300), rnd(rng, 1300)) for g in root}
    gene_defs = {g: (g == "HLA-A", flank[g][0] + root[g] + flank[g][1]) for g in root}
    reads = []
    for g in ("HLA-A", "HLA-B"):
        ids = [r[0] for r in rows if r[1] == g]
        L, R = flank[g]
        reads.append((f"{g}_flanked", noisy(rng, L[-600:] + fw[ids[3]] + R[:600], 12)))      # hg38 mapping starts first: offsets from it
        reads.append((f"{g}_inner", fw[ids[5]][300:2900]))                                   # starts inside the allele: allele-vs-hg38 path
        reads.append((f"{g}_right_only", noisy(rng, fw[ids[7]][700:] + R[:900], 8)))
        reads.append((f"{g}_long_flanks", L[-1250:] + fw[ids[1]] + R[:1250]))                # buffer of 1000 clips the flanks
    reads.append(("junk", rnd(rng, 3000)))
    reads.append(("empty", b""))
    got = host.realign_records_full(gpu, ["HLA-A", "HLA-B"], rows, {g: (f, r.decode()) for g, (f, r) in gene_defs.items()},
                                    [(q, s_.decode()) for q, s_ in reads])
    want = fo.realign_records_full(oracle, ["HLA-A", "HLA-B"], rows, gene_defs, reads)
    assert len(got) == len(want) == len(reads)
    n_realigned = 0
    for g_, w_, (q, s_) in zip(got, want, reads):
        assert g_["gene_name"] == w_["gene_name"], q
        assert g_["mapping_details"] == so.serde_pretty(w_["mapping_details"]), q
        assert g_["read_mapping_stats"] == so.serde_pretty(w_["read_mapping_stats"]), q
        if w_["realigned_record"] is None:
            assert g_["realigned_record"] is None, q
        else:
            ws = w_["realigned_record"]
            assert g_["realigned_record"] == (ws[0], ws[1], ws[2], ws[3], ws[4].decode(), ws[5].decode()), q
            n_realigned += 1
    assert n_realigned == 8 and got[-1]["realigned_record"] is None and got[-2]["realigned_record"] is None
    by = {q: g_ for g_, (q, _) in zip(got, reads)}
    # the flanked read: the segment covers allele + the 600 bp flanks, its offset is where the left flank piece sits in the gene sequence
    seg = by["HLA-A_flanked"]["realigned_record"]
    assert seg[2] == 1300 - 600 and seg[1] - seg[0] >= len(fw[[r[0] for r in rows if r[1] == "HLA-A"][3]]) + 1100
    # the inner read has no flank: offset = allele start in hg38 (1300) + 300 into the allele, give or take the allele's indels
    # (the best allele may be a shorter one lying inside the read: candidates are ranked by nm + unmapped, like best_n hits by score)
    assert abs(by["HLA-A_inner"]["realigned_record"][2] - 1600) <= 12
    assert json.loads(by["HLA-B_inner"]["mapping_details"])["best_star_allele"].startswith("HLA-B*")


# ---- allele-pair diplotype: north_star (2), src/hla/caller.rs:889-901 -----------------------------------------
def test_diplotype_hla_gene_vs_oracle(host, gpu, oracle):
    rows, reads = hla_db()
    genes = {}
    for gene in ("HLA-A", "HLA-B"):
        got = host.diplotype_hla_gene(gpu, rows, gene, cpp_reads(reads[gene]), host.DiplotypeSettings())
        want = fo.diplotype_hla_gene(oracle, rows, gene, reads[gene])
        assert {k: got[k] for k in got if k != "gene_details"} == {k: want[k] for k in want if k != "gene_details"}
        assert got["gene_details"].pretty() == so.serde_pretty(want["gene_details"])
        genes[gene] = (got["gene_details"], want["gene_details"])
        ids = {got["hla_id1"], got["hla_id2"]}
        base = 0 if gene == "HLA-A" else 1000
        assert ids == {f"HLA:HLA{base + 2:05d}", f"HLA:HLA{base + 9:05d}"}  # the two source alleles of the reads
    meta = dict(pbstarphase_version="2.0.1-6-gdeadbee", cpic_version="cpic-v1.44", hla_version="3.57.0", pharmvar_version="6.1.2.1",
                build_time="2024-08-26T00:00:00Z")
    text = host.starphase_json("2.0.1-6-gdeadbee", meta, {g: v[0] for g, v in genes.items()})
    assert text == so.starphase_json("2.0.1-6-gdeadbee", meta, {g: v[1] for g, v in genes.items()})
    assert text.encode() == text.encode("ascii") and text.startswith('{\n  "pbstarphase_version": "2.0.1-6-gdeadbee",')


def test_diplotype_homozygous_and_skewed(host, gpu, oracle):
    from pb_starphase_b200 import synth

    rows, reads = hla_db()
    a = [r for r in rows if r[1] == "HLA-A"]
    rng = np.random.default_rng(12)
    hom, _ = synth.hifi_reads(rng, [a[5][3].encode()], 8)
    hom_reads = [(f"h{k}", hom[k], a[5][4].encode()) for k in range(8)]
    got = host.diplotype_hla_gene(gpu, rows, "HLA-A", cpp_reads(hom_reads), host.DiplotypeSettings())
    want = fo.diplotype_hla_gene(oracle, rows, "HLA-A", hom_reads)
    assert got["hla_id1"] == got["hla_id2"] == a[5][0] and got["gene_details"].pretty() == so.serde_pretty(want["gene_details"])
    # 1 read of a second allele among 24: minor fraction below min_consensus_fraction => homozygous for the majority
    one, _ = synth.hifi_reads(rng, [a[11][3].encode()], 1)
    maj, _ = synth.hifi_reads(rng, [a[5][3].encode()], 23)
    sk = [(f"s{k:02d}", maj[k], a[5][4].encode()) for k in range(23)] + [("s99", one[0], a[11][4].encode())]
    got = host.diplotype_hla_gene(gpu, rows, "HLA-A", cpp_reads(sk), host.DiplotypeSettings())
    want = fo.diplotype_hla_gene(oracle, rows, "HLA-A", sk)
    assert got["hla_id1"] == got["hla_id2"] == a[5][0] == want["hla_id1"] and got["counts1"] == want["counts1"]
    assert got["gene_details"].pretty() == so.serde_pretty(want["gene_details"])
    empty = host.diplotype_hla_gene(gpu, rows, "HLA-A", [], host.DiplotypeSettings())
    assert empty["hla_id1"] == "NO_READS"


# ---- find_best_chain_pair: the reference's own tests, src/cyp2d6/chaining.rs:949-1195 ---------------------------
def d6(name):
    return ("CYP2D6", name, None)


def test_find_best_chain_pair(host, gpu):
    rows = [d6("A"), d6("B"), d6("C"), d6("D")]
    obs = {"seq_1": [[0, 2]], "seq_2": [[1, 1]]}
    scores = {"seq_1": [[(0, 1.0), (1, 1.0), (1, 1.0), (1, 1.0)], [(1, 1.0), (1, 1.0), (0, 1.0), (1, 1.0)]],
              "seq_2": [[(1, 1.0), (0, 1.0), (1, 1.0), (1, 1.0)], [(1, 1.0), (0, 1.0), (1, 1.0), (1, 1.0)]]}
    r = host.find_best_chain_pair(gpu, obs, scores, rows, False, True, ignore_limits=True)
    assert r["best_chains"] == [[0, 2], [1, 1]] and r["dangling"] == ["3_CYP2D6*D"]


def test_ambiguous_find_best_chain_pair(host, gpu):
    rows = [d6("A"), d6("B")]
    obs = {"seq_0": [[1]], "seq_1": [[1, 0]], "seq_2": [[0, 0]], "seq_3": [[0]], "seq_4": [[1]], "seq_5": [[1, 0]], "seq_6": [[0]]}
    a, b = [(0, 1.0), (10, 1.0)], [(10, 1.0), (0, 1.0)]
    scores = {"seq_0": [b], "seq_1": [b, a], "seq_2": [a, a], "seq_3": [a], "seq_4": [b], "seq_5": [b, a], "seq_6": [a]}
    kw = dict(ignore_limits=True, ln_ed_penalty=-math.log(0.01), unexpected_chain_penalty=0.0, inferred_edge_penalty=2.0)
    r = host.find_best_chain_pair(gpu, obs, scores, rows, False, True, lasso_penalty=0.0, **kw)
    assert r["best_chains"] == [[1], [1, 0, 0, 0]] and r["dangling"] == []
    r = host.find_best_chain_pair(gpu, obs, scores, rows, False, True, lasso_penalty=3.0, **kw)
    assert r["best_chains"] == [[1], [1, 0, 0]] and r["dangling"] == []


def pairwise_chains(num_labels, chains):  # create_pairwise_chains, src/cyp2d6/chaining.rs:918-947
    obs, scores, idx = {}, {}, 0
    for chain in chains:
        for a, b in zip(chain, chain[1:]):
            name = f"read_{idx}"
            obs[name] = [[a, b]]
            w = []
            for h in chain:
                row = [(100, 1.0)] * num_labels
                row[h] = (0, 1.0)
                w.append(row)
            scores[name] = w
            idx += 1
    return obs, scores


def test_inferred_alleles(host, gpu):
    rows = [d6("3"), ("link_region", None, None), ("REP7", None, None), ("spacer", None, None), ("CYP2D7", None, None), d6("4"),
            ("Hybrid", "CYP2D6::CYP2D7::exon2", None)]
    obs, scores = pairwise_chains(len(rows), [[0, 1], [2, 3, 4], [5, 1], [2, 3, 6]])
    r = host.find_best_chain_pair(gpu, obs, scores, rows, False, True)
    assert r["best_chains"] == [[0, 1], [5, 1]]
    assert r["dangling"] == ["2_REP7", "3_spacer", "4_CYP2D7", "6_CYP2D6::CYP2D7::exon2"]
    r = host.find_best_chain_pair(gpu, obs, scores, rows, True, True)
    assert r["best_chains"] == [[0, 1, 2, 3, 4], [5, 1, 2, 3, 6]] and r["dangling"] == []


def test_chaining_errors_and_double5(host, gpu):
    rows = [("CYP2D7", None, None), ("link_region", None, None), ("spacer", None, None), ("UNKNOWN", None, None)]
    with pytest.raises(host.NoChainingHead):
        host.find_best_chain_pair(gpu, {}, {}, rows, False, True)
    with pytest.raises(host.HostError, match="Lasso"):
        host.find_best_chain_pair(gpu, {}, {}, rows, False, True, lasso_penalty=-1.0)
    rows = [("CYP2D6*5", None, None)]
    obs = {f"read{x}": [[0]] for x in range(2)}
    scores = {f"read{x}": [[(0, 1.0)]] for x in range(2)}
    r = host.find_best_chain_pair(gpu, obs, scores, rows, True, False)
    assert r["best_chains"] == [[0], [0]] and r["dangling"] == []


def test_find_best_chain_pair_random_vs_oracle(host, gpu):
    """Bound-ordered search on GPU edit distances == the reference's exhaustive loop with its 10-entry heap."""
    import signal

    class TooSlow(Exception):
        pass

    def on_alarm(*_):
        raise TooSlow()

    signal.signal(signal.SIGALRM, on_alarm)
    py_rng = random.Random(77)
    kinds = [("REP6", None), ("CYP2D6", "1"), ("CYP2D6", "4.001"), ("CYP2D6", "2"), ("link_region", None), ("REP7", None), ("spacer", None),
             ("CYP2D7", None), ("CYP2D6*5", None), ("Hybrid", "CYP2D6::CYP2D7::exon2"), ("CYP2D6", "10"), ("CYP2D6", "36")]
    done = 0
    for trial in range(80):
        n = py_rng.randint(2, 8)
        rows = [py_rng.choice(kinds) + (None,) for _ in range(n)]
        labels = fo.labels_from_rows(rows)
        obs, scores = {}, {}
        for r in range(py_rng.randint(1, 14)):
            w = py_rng.randint(1, 3)
            chain = [py_rng.randrange(n) for _ in range(w)]
            obs[f"r{r:02d}"] = [chain] if py_rng.random() < 0.8 else [chain, [py_rng.randrange(n) for _ in range(w)]]
            seg_rows = []
            for t in range(w):
                row = [(py_rng.randint(3, 40), py_rng.choice([1.0, 0.9, 0.5])) for _ in range(n)]
                row[chain[t]] = (py_rng.randint(0, 2), 1.0)
                seg_rows.append(row)
            scores[f"r{r:02d}"] = seg_rows
        infer, norm_all, ign = py_rng.random() < 0.5, py_rng.random() < 0.5, py_rng.random() < 0.3
        signal.alarm(2)  # random label soups can have an astronomical number of chains: only instances the oracle finishes count
        try:
            want = so.find_best_chain_pair(so.Cyp2d6Config.default(), obs, scores, labels, infer, norm_all, so.ChainPenalties(), ign,
                                           return_debug=True)
        except TooSlow:
            continue
        except (so.NoChainingHead, so.NoChainsFound, so.NoScorePairs) as e:
            signal.alarm(0)
            with pytest.raises(getattr(host, type(e).__name__)):
                host.find_best_chain_pair(gpu, obs, scores, rows, infer, norm_all, ignore_limits=ign)
            continue
        finally:
            signal.alarm(0)
        got = host.find_best_chain_pair(gpu, obs, scores, rows, infer, norm_all, ignore_limits=ign)
        best = want[2]["best"]
        assert got["best_chains"] == want[0] and got["dangling"] == want[1], trial
        assert (got["score"], got["i"], got["j"], got["edit_distance"]) == (best["score"], best["i"], best["j"], best["edit_distance"]), trial
        assert got["n_possible_chains"] == len(want[2]["possible_chains"])
        done += 1
    assert done >= 30


# ---- weight_sequence + the whole CYP2D6 chain call: src/cyp2d6/chaining.rs:28-103, caller.rs:430-739 --------------
def cyp_case(seed=1, n_reads=36, scale=320):
    """Two haplotypes REP6 - D6 - link - REP7 - spacer - D7 (hap B = *4.001 with its own region copies) and HiFi-like
    reads that span 2-4 consecutive regions of one haplotype."""
    rng = np.random.default_rng(seed)
    kinds = ["REP6", "CYP2D6", "link_region", "REP7", "spacer", "CYP2D7"]
    base = [rnd(rng, int(scale * f)) for f in (0.9, 1.6, 0.8, 0.9, 0.6, 1.5)]
    cons, rows = [], []
    for hap, sub in enumerate(("1.001", "4.001")):
        for k, kind in enumerate(kinds):
            seq = bytearray(base[k])
            for pos in rng.choice(len(seq), size=max(4, len(seq) // 60), replace=False):  # ~1.7 % haplotype-specific SNPs
                seq[pos] = b"ACGT"[(b"ACGT".index(seq[pos]) + 1 + hap) % 4]
            cons.append(bytes(seq))
            rows.append((kind, sub if kind == "CYP2D6" else None, len(rows)))
    roi = {}
    for r in range(n_reads):
        hap, start, w = int(rng.integers(0, 2)), int(rng.integers(0, 5)), int(rng.integers(2, 5))
        pos, regs = int(rng.integers(50, 300)), []
        for t in range(start, min(start + w, 6)):
            seg = noisy(rng, cons[hap * 6 + t], int(rng.integers(0, 3))).replace(b"N", b"A")
            regs.append((pos, pos + len(seg), seg))
            pos += len(seg)
        roi[f"m64/{r:03d}/ccs"] = regs
    return cons, rows, roi


def test_weight_sequences_vs_oracle(host, gpu, oracle):
    cons, rows, roi = cyp_case()
    segs = [r[2] for q in sorted(roi) for r in roi[q]][:40] + [rnd(np.random.default_rng(4), 300)]  # + one junk segment => []
    rows2 = list(rows)
    rows2[7] = ("UNKNOWN", None, 7)  # skipped consensus (chaining.rs:52-55)
    got = host.weight_sequences(gpu, [s.decode() for s in segs], [c.decode() for c in cons], rows2)
    want = fo.weight_sequences(oracle, segs, cons, fo.labels_from_rows(rows2))
    assert [[tuple(x) for x in ws] for ws in got] == [[tuple(x) for x in ws] for ws in want]
    assert got[-1] == [] and all(ws[7] == (len(s), 0.0) for ws, s in zip(got[:-1], segs))


def test_call_cyp2d6_chains_vs_oracle(host, gpu, oracle):
    for seed, infer in ((1, False), (2, True)):
        cons, rows, roi = cyp_case(seed)
        cpp_roi = {q: [(a, b, s.decode()) for a, b, s in regs] for q, regs in roi.items()}
        got = host.call_cyp2d6_chains(gpu, [c.decode() for c in cons], rows, cpp_roi, infer, True)
        want = fo.call_cyp2d6_chains(oracle, cons, rows, roi, infer, True)
        assert got["best_chains"] == want["best_chains"] and got["score"] == want["score"] and got["dangling"] == want["dangling"]
        assert got["n_possible_chains"] == want["n_possible_chains"]
        text = got["gene_details"].pretty()
        assert text == so.serde_pretty(want["gene_details"])
        back = json.loads(text)
        assert back["diplotypes"][0]["diplotype"] in ("*1.001/*4.001", "*4.001/*1.001") and back["simple_diplotypes"][0]["diplotype"] in ("*1/*4", "*4/*1")
        assert back["multi_mapping_details"] and back["multi_mapping_details"][0]["read_position"].keys() == {"start", "end"}


# ---- template search: Cyp2d6Extractor::find_base_type_in_sequence, src/cyp2d6/haplotyper.rs:142-315 -------------
def cyp_templates(seed=3, scale=260):
    """Small-scale generate_cyp_hybrids: D6, D7 (97 % identical), two D6::D7 / D7::D6 hybrids, *5 signature, REP6, REP7
    (near identical), spacer, link."""
    rng = np.random.default_rng(seed)
    d6 = rnd(rng, int(scale * 3.0))
    d7 = bytearray(d6)
    for pos in rng.choice(len(d7), size=len(d7) // 33, replace=False):
        d7[pos] = b"ACGT"[(b"ACGT".index(d7[pos]) + 1) % 4]
    d7 = bytes(d7)
    h = len(d6) // 2
    rep6 = rnd(rng, int(scale * 1.4))
    rep7 = rep6[:-12] + rnd(rng, 12)
    star5 = rnd(rng, int(scale * 0.9)) + rnd(rng, int(scale * 0.9))
    return dict(d6=d6, d7=d7, rep6=rep6, rep7=rep7, spacer=rnd(rng, int(scale * 1.1)), link=rnd(rng, int(scale * 1.5)), star5=star5,
                templates=[("CYP2D6", None, d6), ("CYP2D7", None, d7), ("Hybrid", "CYP2D6::CYP2D7::exon2", d6[:h] + d7[h:]),
                           ("Hybrid", "CYP2D7::CYP2D6::exon2", d7[:h] + d6[h:]), ("CYP2D6*5", None, star5), ("REP6", None, rep6), ("REP7", None, rep7),
                           ("spacer", None, None), ("link_region", None, None)])


def test_find_base_type_in_sequences_vs_oracle(host, gpu, oracle):
    c = cyp_templates()
    templates = [(t, s, (seq if seq is not None else c["spacer" if t == "spacer" else "link"])) for t, s, seq in c["templates"]]
    rng = np.random.default_rng(8)
    fl = lambda n: rnd(rng, n)  # noqa: E731
    hap = c["rep6"] + c["d6"] + c["link"] + c["rep7"] + c["spacer"] + c["d7"]
    seqs = [
        fl(150) + noisy(rng, hap, 12).replace(b"N", b"A") + fl(120),                                   # the reference layout
        fl(80) + noisy(rng, c["rep6"] + c["d6"] + c["link"] + c["rep6"] + c["d6"] + c["link"], 10).replace(b"N", b"A") + fl(90),  # duplication: D6 twice
        c["d6"][len(c["d6"]) // 3:] + c["link"] + fl(40),                                              # read starts inside D6: clipped template
        fl(100) + c["star5"] + fl(100),                                                                # *5 signature
        fl(700),                                                                                       # nothing to find
        b"",                                                                                           # :148-151
        fl(50) + templates[2][2] + c["link"] + fl(50),                                                 # a hybrid: must win over D6 / D7
    ]
    names = {}
    for mmf in (0.1, 0.5):
        got = host.find_base_type_in_sequences(gpu, [(t, s, q.decode()) for t, s, q in templates], [s.decode() for s in seqs], False, mmf)
        want = fo.find_base_type_in_sequences(oracle, templates, seqs, mmf)
        assert [[tuple(h[:3]) + (tuple(h[3]),) for h in one] for one in got] == want
        names[mmf] = [[h[0] for h in one] for one in got]
    # a third of D6 is missing from the read: reported when half may be missing, dropped at 10 % (haplotyper.rs:297-312)
    assert names[0.5][2] == ["CYP2D6", "link_region"] and names[0.1][2] == ["link_region"]
    assert got[2][0][3][3:] == (len(c["d6"]) // 3, 0)  # clipped_start, clipped_end
    names = names[0.5]
    assert names[0] == ["REP6", "CYP2D6", "link_region", "REP7", "spacer", "CYP2D7"]
    assert names[1].count("CYP2D6") == 2 and names[1].count("link_region") == 2
    assert names[3] == ["CYP2D6*5"] and names[4] == [] and names[5] == [] and "CYP2D6::CYP2D7::exon2" in names[6]


@pytest.mark.parametrize("is_forward_strand", [True, False])
def test_diplotype_from_records_derives_targets_on_device(host, gpu, is_forward_strand):
    """diplotype_hla_gene_records: the reads go up once, the DNA targets (reverse-complemented for a reverse-strand gene) and the
    cDNA targets (splice_read's exon intervals; "N" when nothing is left) are cut on the device (sp_targets_derive).  Same call and
    same gene-details JSON as the flow fed with host-prepared targets (prepare_score_read_targets, src/hla/caller.rs:1337-1368)."""
    rows, reads = hla_db()
    s = host.DiplotypeSettings()
    index = host.HlaGeneIndex(gpu, rows, "HLA-A", s)
    exons = [(5100, 5400), (5900, 6300), (7000, 7250), (9000, 9100)]
    records, prepared = [], []
    for k, (q, dna, _c) in enumerate(reads["HLA-A"]):
        seq = dna.decode() if is_forward_strand else host.reverse_complement(dna.decode())  # as the aligned record holds it
        n = len(seq)
        if k == 0:
            cigar, pos = [(n, 0)], 20000                                   # covers no exon: cDNA target "N"
        elif k % 3 == 1:
            cigar, pos = [(30, 4), (500, 7), (3, 1), (n - 533 - 12, 0), (12, 4)], 5000 + k   # clips, an insertion
        else:
            cigar, pos = [(900, 0), (7, 2), (n - 900, 8)], 5000 + k                            # a deletion inside exon 2
        records.append((q, seq, pos, cigar))
        d_t, c_t = host.prepare_score_read_targets(seq, pos, cigar, exons, is_forward_strand, s)
        prepared.append((q, d_t, c_t))
    assert prepared[0][2] == "N" and all(len(p[2]) > 300 for p in prepared[1:])
    want = host.diplotype_hla_gene_indexed(gpu, index, prepared, s)
    got = host.diplotype_hla_gene_records(gpu, index, records, exons, is_forward_strand, s)
    assert {k: got[k] for k in got if k != "gene_details"} == {k: want[k] for k in want if k != "gene_details"}
    assert got["gene_details"].pretty() == want["gene_details"].pretty()
    assert host.diplotype_hla_gene_records(gpu, index, [], exons, True, s)["hla_id1"] == "NO_READS"


# ---- a cohort worker pool: several GpuAligners on one GPU, one host thread each (sp_ctx_share_device) -----------------
def test_concurrent_aligners_share_one_gpu(host, gpu, oracle):
    """Three host threads, one GpuAligner each in share_device mode, run HLA calls, the template search and the chain call
    side by side (the pybind layer drops the interpreter lock for the C++ calls); every result equals the one the
    single aligner of this module gives on its own."""
    import threading

    rows, reads = hla_db(seed=21, n_alleles=40, n_reads=16)
    cons, crow, roi = cyp_case(1)
    cpp_roi = {q: [(a, b, s.decode()) for a, b, s in regs] for q, regs in roi.items()}
    c = cyp_templates()
    templates = [(t, s, (seq if seq is not None else c["spacer" if t == "spacer" else "link"]).decode()) for t, s, seq in c["templates"]]
    rng = np.random.default_rng(9)
    seqs = [(rnd(rng, 100) + noisy(rng, c["rep6"] + c["d6"] + c["link"] + c["rep7"] + c["spacer"] + c["d7"], 10).replace(b"N", b"A") + rnd(rng, 80)).decode()
            for _ in range(6)]
    settings = host.DiplotypeSettings()

    def sample(g):
        out = [host.diplotype_hla_gene(g, rows, gene, cpp_reads(reads[gene]), settings)["gene_details"].pretty() for gene in ("HLA-A", "HLA-B")]
        out.append(host.find_base_type_in_sequences(g, templates, seqs, False, 0.5))
        out.append(host.call_cyp2d6_chains(g, [x.decode() for x in cons], crow, cpp_roi, False, True)["gene_details"].pretty())
        return out

    want = sample(gpu)
    results, errors = {}, []

    def work(k):
        try:
            g = host.GpuAligner(0)
            g.share_device(True)
            results[k] = [sample(g) for _ in range(3)]
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in range(3):
        assert all(r == want for r in results[k])
