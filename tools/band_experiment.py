#!/usr/bin/env python3
"""Which diagonal band does the affine refinement (K9) need around a unit-cost placement (K4)?  CPU only, build container.

For pairs of the two HLA call sites (score_read: every nearby allele against a consensus-like target, a = 5; realigner: a read
against its candidate alleles, a = 1) on the real IMGT alleles inside their hg38 flanks, compares the unbanded affine optimum
with the banded one under two band rules:
  A  half width = half the start/end diagonal difference + nm + 24           (every edit may turn into drift)
  B  half width = half the diagonal hull of the unit-cost path + 24          (the path's own excursions + slack)
and prints, per rule, the band class histogram and how many records differ from the unbanded model.
  python tools/band_experiment.py [--targets 4] [--alleles 80]"""
import argparse
import gzip
import json
import sys
from collections import Counter
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))

import oracle_util as ou  # noqa: E402
import affine_divergence as ad  # noqa: E402
from pb_starphase_b200 import synth  # noqa: E402

CLASSES = (31, 63, 127, 255)


def hull(u):
    d = u["t_start"] - u["p_start"]
    lo = hi = d
    for ln, op in u["cigar"]:
        if op == 1:      # I: pattern bases without text
            d -= ln
        elif op == 2:    # D: text bases without pattern
            d += ln
        lo, hi = min(lo, d), max(hi, d)
    return lo, hi


def rule_a(u):
    d0, d1 = u["t_start"] - u["p_start"], u["t_end"] - u["p_end"]
    return (d0 + d1) // 2, (abs(d1 - d0) + 1) // 2 + u["nm"] + 24


def rule_b(u):
    lo, hi = hull(u)
    return (lo + hi) // 2, (hi - lo + 1) // 2 + 24


def key(r):
    return (r["score"], r["nm"], r["p_start"], r["p_end"], r["t_start"], r["t_end"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=4)
    ap.add_argument("--alleles", type=int, default=80)
    args = ap.parse_args()
    orc = ou.Oracle()
    db = json.load(gzip.open(ad.REAL_DB))
    rows = {}
    for g in ("HLA-A", "HLA-B"):
        rows[g] = [(k, g, v["star_allele"], v["dna_sequence"], v["cdna_sequence"]) for k, v in sorted(db["hla_sequences"].items(), key=lambda kv: kv[0].encode())
                   if v["gene_name"] == g and v.get("dna_sequence")]
    hap = ad.real_haplotypes(orc, db, rows)
    rng = np.random.default_rng(5)
    out = {}
    for site, costs in (("score_read a=5", ou.COSTS_ALLELE_SCORING), ("realign a=1", ou.COSTS_MAP_HIFI)):
        aff = ou.AffineOracle(costs)
        stats = {r: dict(classes=Counter(), differs=0, skipped=0) for r in "AB"}
        n = 0
        for gene in rows:
            dna = [r[3].encode() for r in rows[gene]]
            for t in range(args.targets):
                src = int(rng.integers(len(dna)))
                if site.startswith("score_read"):
                    text = synth.hifi_reads(rng, [hap(gene, src)], 1, err=0.0002, flank=0, lo=0, hi=1 << 20)[0][0]
                else:
                    text = synth.hifi_reads(rng, [hap(gene, src)], 1, err=0.002, flank=0, lo=0, hi=1 << 20)[0][0]
                idx = ad.neighbourhood(orc, text, dna, args.alleles * 3 // 4, args.alleles // 4, rng)
                for a in idx:
                    u = orc.align(dna[a], text)
                    if not u["cigar"]:
                        continue
                    full = key(aff.align(dna[a], text))
                    n += 1
                    for r, rule in (("A", rule_a), ("B", rule_b)):
                        c, w = rule(u)
                        if w > CLASSES[-1]:
                            stats[r]["skipped"] += 1
                            continue
                        band = next(b for b in CLASSES if w <= b)
                        stats[r]["classes"][band] += 1
                        if key(aff.align(dna[a], text, centre=c, band=band)) != full:
                            stats[r]["differs"] += 1
                print(site, gene, "target", t, {r: dict(differs=s["differs"], skipped=s["skipped"], classes=dict(s["classes"])) for r, s in stats.items()}, "pairs", n, flush=True)
        out[site] = dict(pairs=n, **{r: dict(differs=s["differs"], skipped=s["skipped"], classes=dict(s["classes"])) for r, s in stats.items()})
    print(json.dumps(out, indent=1))
    Path(ROOT / "profiles" / "r02_band_experiment.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
