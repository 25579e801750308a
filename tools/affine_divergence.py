#!/usr/bin/env python3
"""Measures how far the product's unit-cost quantities are from the reference's cost model (CPU only; run in the build
container).  Product numbers = the unit-cost oracle (the GPU kernels are bit-exact against it, tests/test_k*_gpu.py);
model numbers = oracle/sp_oracle_affine.c through tests/affine_model.py.  Writes profiles/r02_affine_divergence.json, the
source of the table in DESIGN.md §3.

Data: (a) tests/golden/hla_faux.json (BASELINE configs[0]); (b) a seeded sample of the bench workload (configs[1]);
(c) the real IMGT/HLA 3.57.0 allele set, read from /root/reference/data/v0.14.1 when that checkout exists (never copied);
(d) the synthetic CYP2D6 diploid sample (configs[3]).

  python tools/affine_divergence.py [--pairs 1500] [--alleles 200] [--targets 6] [--out profiles/r02_affine_divergence.json]
"""
from __future__ import annotations

import argparse
import gzip
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import affine_model as am  # noqa: E402
import flow_oracle as fo  # noqa: E402
import oracle_util as ou  # noqa: E402
from pb_starphase_b200 import synth  # noqa: E402

REAL_DB = Path("/root/reference/data/v0.14.1/pbstarphase_20240826.json.gz")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def neighbourhood(orc, target: bytes, seqs, n_top: int, n_rand: int, rng) -> list:
    """Indices of the alleles that matter for a call on `target`: the n_top closest by unit-cost distance plus random others."""
    D = orc.score_batch([target], seqs)[0]
    top = list(np.argsort(D, kind="stable")[:n_top])
    rest = [i for i in range(len(seqs)) if i not in set(top)]
    extra = list(rng.choice(rest, size=min(n_rand, len(rest)), replace=False)) if rest else []
    return sorted(int(i) for i in top + extra)


def hla_section(name, orc, rows_by_gene, args, rng, haplotype=None) -> dict:
    """rows_by_gene: {gene: [DbRow]} with DNA present.  Reads are simulated from haplotype(gene, allele index) -- the allele
    inside its genomic flanks; default: 250 random bases either side, as in the bench workload -- with 0.2 % HiFi-like errors,
    consensus-like targets with a tenth of that error rate."""
    out = {}
    t0 = time.time()

    def simulate(gene, dna, idx, n, err):
        """n reads drawn from the alleles listed in idx: (reads, source position in idx)."""
        if haplotype is None:
            return synth.hifi_reads(rng, [dna[i] for i in idx], n, err=err)
        return synth.hifi_reads(rng, [haplotype(gene, i) for i in idx], n, err=err, flank=0, lo=0, hi=1 << 20)

    # ---- pair level: random (read, allele) pairs of the whole set, both cost sets ----
    for gene, rows in rows_by_gene.items():
        dna = [r[3].encode() for r in rows]
        cdna = [r[4].encode() for r in rows]
        pick = [int(x) for x in rng.choice(len(dna), size=min(64, len(dna)), replace=False)]
        reads, src = simulate(gene, dna, pick, 64, 0.002)
        src = [pick[int(s)] for s in src]
        ctg = [synth.hifi_reads(rng, [cdna[int(s)]], 1, flank=50, lo=0, hi=1400)[0][0] for s in src]
        n = args.pairs // (2 * len(rows_by_gene))
        # half of the sampled pairs are near-identical (allele drawn from the read's neighbourhood), half random
        pairs = []
        for k in range(n):
            r = int(rng.integers(0, len(reads)))
            if k % 2 == 0:
                pairs.append((r, int(rng.integers(0, len(dna)))))
            else:
                nb = neighbourhood(orc, reads[r], dna, 20, 0, rng)
                pairs.append((r, int(rng.choice(nb))))
        out[f"{gene} pairs DNA a=5 (score_read site)"] = am.pair_divergence(orc, ou.COSTS_ALLELE_SCORING, reads, dna, pairs, args.threads)
        out[f"{gene} pairs DNA a=1 (realigner site)"] = am.pair_divergence(orc, ou.COSTS_MAP_HIFI, reads, dna, pairs, args.threads)
        out[f"{gene} pairs cDNA a=5 (score_read site)"] = am.pair_divergence(orc, ou.COSTS_ALLELE_SCORING, ctg, cdna, pairs, args.threads)
        log(f"[{name}] {gene} pair level done {time.time() - t0:.0f}s")
        # ---- call level ----
        aff5 = am.AffineFlowOracle(orc, ou.COSTS_ALLELE_SCORING, args.threads)
        aff1 = am.AffineFlowOracle(orc, ou.COSTS_MAP_HIFI, args.threads)
        targets, sr = [], dict(targets=0, best_allele_differs=0, best_stats_differ=0, alleles_with_different_stats=0, alleles=0, floor_disagrees=0)
        for k in range(args.targets):
            a = int(rng.integers(0, len(dna)))
            t_dna = simulate(gene, dna, [a], 1, 0.0002)[0][0]
            t_cdna = synth.hifi_reads(rng, [cdna[a]], 1, err=0.0002, flank=50, lo=0, hi=1400)[0][0]
            sub = [rows[i] for i in neighbourhood(orc, t_dna, dna, args.alleles * 3 // 4, args.alleles // 4, rng)]
            res = am.score_read_flips(orc, aff5, sub, gene, [(t_dna, t_cdna)])
            for key in sr:
                sr[key] += res[key]
            log(f"[{name}] {gene} score_read target {k}: {res} {time.time() - t0:.0f}s")
        out[f"{gene} score_read ({args.alleles} nearest + random alleles per target)"] = sr
        # north_star (2) + realigner: a het sample of `reads_per_sample` reads from two alleles
        dp = dict(samples=0, pair_differs=0, hom_het_differs=0, cells=0, cells_differ=0)
        rl = dict(reads=0, assignment_differs=0, stats_differ=0, model_best_not_in_product_candidates=0)
        for k in range(args.samples):
            a1, a2 = (int(x) for x in rng.choice(len(dna), size=2, replace=False))
            rd, s2 = simulate(gene, dna, [a1, a2], args.reads_per_sample, 0.002)
            ct = [synth.hifi_reads(rng, [cdna[(a1, a2)[int(s)]]], 1, flank=50, lo=0, hi=1400)[0][0] for s in s2]
            nb = sorted(set(neighbourhood(orc, rd[0], dna, args.alleles // 3, args.alleles // 6, rng)) |
                        set(neighbourhood(orc, dna[a2], dna, args.alleles // 3, 0, rng)) | {a1, a2})
            sub = [rows[i] for i in nb]
            sample = [(f"r{q}", rd[q], ct[q]) for q in range(len(rd))]
            res = am.diplotype_flips(orc, aff5, sub, gene, [sample])
            for key in dp:
                dp[key] += res[key]
            res2 = am.realign_flips(orc, aff1, sub, [gene], [(q, s) for q, s, _ in sample])
            for key in rl:
                rl[key] += res2[key]
            log(f"[{name}] {gene} sample {k}: K2 {res} realign {res2} {time.time() - t0:.0f}s")
        out[f"{gene} allele-pair diplotype (north_star 2)"] = dp
        out[f"{gene} realign_record"] = rl
    return out


def real_haplotypes(orc, db, rows_by_gene, flank: int = 250):
    """haplotype(gene, a): allele a of the real database placed into the hg38 sequence of its gene (test_data/refseq_faux: chr6
    N-masked except HLA-A / HLA-B +- 2 kb), in gene orientation, cut `flank` bases outside the gene coordinates -- so that a read
    carries real genomic flanks, and a longer allele's extra UTR bases meet the sequence they were taken from."""
    import re

    fa = gzip.open("/root/reference/test_data/refseq_faux/hg38_chr6_masked.fa.gz").read().decode()
    seq = fa.split("\n", 1)[1].replace("\n", "")
    region, cache = {}, {}
    for gene in rows_by_gene:
        c = db["hla_config"]["hla_coordinates"][gene]
        lo, hi = c["start"] - flank, c["end"] + flank
        g = seq[lo:hi].upper().encode()
        assert re.fullmatch(rb"[ACGT]+", g), "masked reference does not cover the gene + flanks"
        region[gene] = g if db["hla_config"]["hla_is_forward_strand"][gene] else fo.so.reverse_complement(g)

    def haplotype(gene, a):
        if (gene, a) not in cache:
            allele = rows_by_gene[gene][a][3].encode()
            _, S, E = orc.score_spans([region[gene]], [allele])
            cache[(gene, a)] = region[gene][:int(S[0, 0])] + allele + region[gene][int(E[0, 0]):]
        return cache[(gene, a)]

    return haplotype


def faux_section(orc) -> dict:
    g = json.loads((ROOT / "tests/golden/hla_faux.json").read_text())
    rows = [(k, v["gene_name"], v["star_allele"], v["dna_sequence"], v["cdna_sequence"]) for k, v in g["hla_sequences"].items()]
    rng = np.random.default_rng(7)
    out = {}
    for gene in ("HLA-A", "HLA-B"):
        row = next(r for r in rows if r[1] == gene)
        dna, cdna = row[3].encode(), row[4].encode()
        snp = bytearray(dna); snp[1500] = ord("A") if snp[1500] != ord("A") else ord("C")
        hp = dna[:2000] + dna[2000:2001] + dna[2000:]
        targets = [(dna, cdna), (bytes(snp), cdna), (hp, cdna), (b"ACGT", b"N")]
        aff5 = am.AffineFlowOracle(orc, ou.COSTS_ALLELE_SCORING)
        out[f"{gene} score_read: exact copy, 1 SNP, 1-bp homopolymer insertion, 4-bp junk"] = am.score_read_flips(orc, aff5, rows, gene, targets)
        pairs = [(t, 0) for t in range(3)]
        out[f"{gene} pairs"] = am.pair_divergence(orc, ou.COSTS_ALLELE_SCORING, [t[0] for t in targets[:3]], [dna], pairs)
    del rng
    return out


def cyp_section(orc, args) -> dict:
    c = synth.cyp2d6_diploid_sample(2001)
    labels = fo.labels_from_rows(c["regions"])
    qn = sorted(c["roi"])[:args.cyp_reads]
    segs = [r[2] for q in qn for r in c["roi"][q]]
    aff1 = am.AffineFlowOracle(orc, ou.COSTS_MAP_HIFI, args.threads)
    t0 = time.time()
    out = {"weight_sequence (segments of %d reads x %d consensuses)" % (len(qn), len(c["consensuses"])):
           am.weight_sequence_flips(orc, aff1, segs, c["consensuses"], labels)}
    log(f"[cyp2d6] weight_sequence done {time.time() - t0:.0f}s")
    # chain call on model numbers vs product numbers
    roi = {q: c["roi"][q] for q in qn}
    prod = fo.call_cyp2d6_chains(orc, c["consensuses"], c["regions"], roi, False, True)
    saved = fo.NO_MAPPING_PERMILLE
    fo.NO_MAPPING_PERMILLE = 1000
    try:
        model = fo.call_cyp2d6_chains(aff1, c["consensuses"], c["regions"], roi, False, True)
    finally:
        fo.NO_MAPPING_PERMILLE = saved
    out["call_cyp2d6_chains"] = dict(product=prod["gene_details"]["diplotypes"][0]["diplotype"], model=model["gene_details"]["diplotypes"][0]["diplotype"],
                                     same_best_chains=prod["best_chains"] == model["best_chains"],
                                     same_gene_details_json=fo.so.serde_pretty(prod["gene_details"]) == fo.so.serde_pretty(model["gene_details"]))
    # template search (a10): product flow vs the same flow on model numbers
    templates = [(t, s, q) for (t, s), q in zip(c["template_labels"], c["templates"])]
    reads = c["reads"][:args.cyp_template_reads]
    hp = fo.find_base_type_in_sequences(orc, templates, reads, 0.5)
    hm = fo.find_base_type_in_sequences(aff1, templates, reads, 0.5)
    out["find_base_type_in_sequence (%d reads x 39 templates)" % len(reads)] = dict(
        reads=len(reads), regions_product=sum(map(len, hp)), regions_model=sum(map(len, hm)),
        reads_with_different_labels=sum([h[0] for h in a] != [h[0] for h in b] for a, b in zip(hp, hm)),
        reads_with_different_spans=sum([h[:3] for h in a] != [h[:3] for h in b] for a, b in zip(hp, hm)),
        reads_with_different_stats=sum(a != b for a, b in zip(hp, hm)))
    log(f"[cyp2d6] template search done {time.time() - t0:.0f}s")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1200)
    ap.add_argument("--alleles", type=int, default=160)
    ap.add_argument("--targets", type=int, default=4)
    ap.add_argument("--samples", type=int, default=2)
    ap.add_argument("--reads-per-sample", type=int, default=12)
    ap.add_argument("--cyp-reads", type=int, default=24)
    ap.add_argument("--cyp-template-reads", type=int, default=6)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--skip", default="", help="comma list of sections to skip: faux,synthetic,imgt,cyp2d6")
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r02_affine_divergence.json"))
    args = ap.parse_args()
    skip = set(filter(None, args.skip.split(",")))
    orc = ou.Oracle()
    res = dict(note="product = unit-cost oracle (== the CUDA kernels, bit for bit); model = best local alignment under minimap2's "
                    "two-piece affine costs (oracle/sp_oracle_affine.c).  Not minimap2 itself: seeding / chaining / z-drop are not modelled.",
               args=vars(args))
    if "faux" not in skip:
        res["configs[0] HLA-faux"] = faux_section(orc)
    if "synthetic" not in skip:
        genes = synth.hla_wgs_workload(synth.DEFAULT_SEED, 64, 1.0)
        rows = {g: [(f"HLA:SYN{gi}{a:05d}", g, ["01", f"{a:05d}"], v["dna"][a].decode(), v["cdna"][a].decode()) for a in range(len(v["dna"]))]
                for gi, (g, v) in enumerate(genes.items())}
        res["configs[1] synthetic allele trees"] = hla_section("synthetic", orc, rows, args, np.random.default_rng(11))
    if "imgt" not in skip and REAL_DB.exists():
        db = json.load(gzip.open(REAL_DB))
        rows = {}
        for g in ("HLA-A", "HLA-B"):
            rows[g] = [(k, g, v["star_allele"], v["dna_sequence"], v["cdna_sequence"]) for k, v in sorted(db["hla_sequences"].items(), key=lambda kv: kv[0].encode())
                       if v["gene_name"] == g and v.get("dna_sequence")]
        res["IMGT/HLA 3.57.0 alleles (data/v0.14.1 of the reference checkout) inside their hg38 flanks, simulated reads"] = \
            hla_section("imgt", orc, rows, args, np.random.default_rng(13), real_haplotypes(orc, db, rows))
    if "cyp2d6" not in skip:
        res["configs[3] CYP2D6 diploid sample"] = cyp_section(orc, args)
    Path(args.out).write_text(json.dumps(res, indent=1, default=int) + "\n")
    print(json.dumps(res, indent=1, default=int))


if __name__ == "__main__":
    main()
