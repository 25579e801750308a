"""Seeded synthetic inputs of the shapes SURVEY.md §8(d) names (there is no network / no BAM
fixture, so every bench and most tests run on these).  Pure numpy; deterministic per seed."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
DEFAULT_SEED = 20251106

# (n_dna_alleles, root_len, min_len, max_len, n_cdna_alleles, cdna_len) -- v2.0.0 DB statistics,
# /root/reference/data/v2.0.0/pbstarphase_20251106.db_stat.txt:44-47 and SURVEY.md §6
GENE_SHAPES = {
    "HLA-A": dict(n_dna=5695, root=3100, lo=1301, hi=3537, n_cdna=8949, cdna=1098),
    "HLA-B": dict(n_dna=6756, root=2990, lo=2657, hi=4103, n_cdna=10680, cdna=1089),
}


def random_seq(rng: np.random.Generator, n: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=n)]


def mutate(rng: np.random.Generator, seq: np.ndarray, n_snp: int, n_indel: int) -> np.ndarray:
    s = seq.copy()
    if len(s) == 0:
        return s
    for pos in rng.integers(0, len(s), size=n_snp):
        s[pos] = ACGT[(np.searchsorted(ACGT, s[pos]) + rng.integers(1, 4)) % 4]
    for _ in range(n_indel):
        pos = int(rng.integers(0, len(s)))
        ln = int(rng.integers(1, 4))
        if rng.random() < 0.5:
            s = np.concatenate([s[:pos], random_seq(rng, ln), s[pos:]])
        else:
            s = np.concatenate([s[:pos], s[pos + ln:]])
    return s


def allele_tree(rng: np.random.Generator, n: int, root_len: int, lo: int, hi: int, partial_frac: float = 0.12) -> List[bytes]:
    """Mutation tree from one random root: each child copies a random earlier allele and gets
    Poisson(6) SNPs and (10 %) a 1-3 bp indel, so neighbours differ by a few edits like IMGT/HLA.
    A fraction is truncated (partial entries) or extended so lengths span [lo, hi]."""
    seqs = [random_seq(rng, root_len)]
    for _ in range(1, n):
        parent = seqs[int(rng.integers(0, len(seqs)))]
        child = mutate(rng, parent, int(rng.poisson(6)), 1 if rng.random() < 0.10 else 0)
        seqs.append(child)
    out = []
    for s in seqs:
        r = rng.random()
        if r < partial_frac and len(s) > lo:
            new_len = int(rng.integers(lo, len(s)))
            start = int(rng.integers(0, len(s) - new_len + 1))
            s = s[start:start + new_len]
        elif r > 1.0 - partial_frac / 2 and len(s) < hi:
            ext = int(rng.integers(1, hi - len(s) + 1))
            s = np.concatenate([random_seq(rng, ext // 2), s, random_seq(rng, ext - ext // 2)])
        s = s[:hi]
        out.append(s.tobytes())
    return out


def hifi_reads(rng: np.random.Generator, alleles: List[bytes], n: int, err: float = 0.002, flank: int = 250,
               lo: int = 2800, hi: int = 4400) -> Tuple[List[bytes], np.ndarray]:
    """Reads = random allele + `err` errors (70 % homopolymer +-1, rest substitutions) + random flanks."""
    reads, src = [], np.zeros(n, dtype=np.int64)
    for i in range(n):
        a = int(rng.integers(0, len(alleles)))
        src[i] = a
        s = np.frombuffer(alleles[a], dtype=np.uint8)
        n_err = int(rng.poisson(err * len(s)))
        s = s.copy()
        for _ in range(n_err):
            pos = int(rng.integers(0, len(s)))
            if rng.random() < 0.7:  # homopolymer length error
                if rng.random() < 0.5:
                    s = np.concatenate([s[:pos], s[pos:pos + 1], s[pos:]])
                else:
                    s = np.concatenate([s[:pos], s[pos + 1:]])
            else:
                s[pos] = ACGT[(np.searchsorted(ACGT, s[pos]) + rng.integers(1, 4)) % 4]
        s = np.concatenate([random_seq(rng, flank), s, random_seq(rng, flank)])
        if len(s) > hi:
            s = s[:hi]
        reads.append(s.tobytes())
    return reads, src


def hla_gene(seed: int, gene: str, n_alleles: int | None = None, n_reads: int = 64, with_cdna: bool = False):
    """One gene's synthetic workload: (dna_alleles, reads, source_allele_index[, cdna_alleles])."""
    shape = GENE_SHAPES[gene]
    rng = np.random.default_rng([seed, sum(gene.encode())])
    n = n_alleles if n_alleles is not None else shape["n_dna"]
    alleles = allele_tree(rng, n, shape["root"], shape["lo"], shape["hi"])
    reads, src = hifi_reads(rng, alleles, n_reads)
    if not with_cdna:
        return alleles, reads, src
    nc = shape["n_cdna"] if n_alleles is None else n_alleles
    cdna = allele_tree(rng, nc, shape["cdna"], 534, 1208, partial_frac=0.05)
    return alleles, reads, src, cdna


def hla_wgs_workload(seed: int = DEFAULT_SEED, n_reads: int = 2048, scale: float = 1.0):
    """BASELINE.json configs[1] ("HLA-A/HLA-B WGS 30x: ~2k synthetic HiFi reads x full IMGT/HLA allele set"),
    SURVEY.md §8(d).2.  Per gene: DNA alleles, cDNA alleles (the first n_dna of them belong to the DNA
    alleles in order, the rest are cDNA-only entries), reads drawn from that gene's DNA alleles and a
    cDNA target per read (its source allele's cDNA + errors + 50 bp flanks, the stand-in for splice_read,
    src/hla/caller.rs:1518-1576).  `scale` shrinks the allele counts for tests / CPU samples."""
    genes = {}
    per_gene_reads = n_reads // 2
    for gi, gene in enumerate(("HLA-A", "HLA-B")):
        shape = GENE_SHAPES[gene]
        rng = np.random.default_rng([seed, gi, 17])
        n_dna = max(2, int(round(shape["n_dna"] * scale)))
        n_cdna = max(n_dna, int(round(shape["n_cdna"] * scale)))
        dna = allele_tree(rng, n_dna, shape["root"], shape["lo"], shape["hi"])
        cdna = allele_tree(rng, n_cdna, shape["cdna"], 534, 1208, partial_frac=0.05)
        reads, src = hifi_reads(rng, dna, per_gene_reads if gi == 0 else n_reads - per_gene_reads)
        ctargets = []  # cDNA target k belongs to read k
        for s in src:
            one, _ = hifi_reads(rng, [cdna[int(s)]], 1, flank=50, lo=0, hi=1400)
            ctargets.append(one[0])
        genes[gene] = dict(dna=dna, cdna=cdna, reads=reads, ctargets=ctargets, src=src)
    return genes


# ---------------------------------------------------------------------------------------------
# CYP2D6-shaped sample: region lengths of src/cyp2d6/definitions.rs:13-14, :137-172
# ---------------------------------------------------------------------------------------------
CYP_REGION_LENS = dict(d6=6165, d7=5938, star5=3500, rep6=2772, rep7=2772, spacer=1564, link=2919)


def cyp2d6_templates(rng: np.random.Generator):
    """The 39 search templates of generate_cyp_hybrids (src/cyp2d6/definitions.rs:346-464): D6, D7 (97 % identical),
    *5 signature, 16 + 16 D6/D7 hybrids, REP6, REP7, spacer, link."""
    d6 = random_seq(rng, CYP_REGION_LENS["d6"])
    d7 = mutate(rng, d6, int(0.03 * len(d6)), 12)[: CYP_REGION_LENS["d7"]]
    regions = dict(d6=d6, d7=d7, star5=random_seq(rng, CYP_REGION_LENS["star5"]),
                   rep6=random_seq(rng, CYP_REGION_LENS["rep6"]), spacer=random_seq(rng, CYP_REGION_LENS["spacer"]),
                   link=random_seq(rng, CYP_REGION_LENS["link"]))
    regions["rep7"] = mutate(rng, regions["rep6"], 60, 3)
    hybrids = []
    for k in range(16):
        cut = 350 + 340 * k
        hybrids.append(np.concatenate([d6[:cut], d7[cut:]]))
        hybrids.append(np.concatenate([d7[:cut], d6[cut:]]))
    templates = [d6, d7, regions["star5"]] + hybrids + [regions["rep6"], regions["rep7"], regions["spacer"], regions["link"]]
    assert len(templates) == 39
    return regions, [t.tobytes() for t in templates]


def cyp2d6_sample(seed: int, n_reads: int = 256, n_consensus: int = 24, n_chains: int = 200):
    """One sample's CYP2D6 inputs: templates (39), consensuses (K haplotype-region sequences), reads (several regions
    in a row + flanks), per-read segments (the region slices weight_sequence scores, src/cyp2d6/caller.rs:435-517),
    and candidate chains over the consensus indices."""
    rng = np.random.default_rng([seed, 2, 6])
    regions, templates = cyp2d6_templates(rng)
    order = ["rep6", "d6", "link", "rep7", "spacer", "d7"]  # the reference haplotype layout
    cons, cons_kind = [], []
    for k in range(n_consensus):
        kind = order[k % len(order)]
        cons.append(mutate(rng, regions[kind], int(rng.integers(0, 12)), int(rng.integers(0, 2))).tobytes())
        cons_kind.append(kind)
    by_kind = {kind: [i for i, kk in enumerate(cons_kind) if kk == kind] for kind in order}
    reads, segments, seg_read = [], [], []
    for r in range(n_reads):
        start = int(rng.integers(0, len(order)))
        w = int(rng.integers(1, 5))
        parts = []
        for t in range(w):
            kind = order[(start + t) % len(order)]
            c = cons[int(rng.choice(by_kind[kind]))]
            seg, _ = hifi_reads(rng, [c], 1, flank=0, lo=0, hi=1 << 30)
            parts.append(seg[0])
            segments.append(seg[0])
            seg_read.append(r)
        reads.append(random_seq(rng, 300).tobytes() + b"".join(parts) + random_seq(rng, 300).tobytes())
    chains = []
    for _ in range(n_chains):
        start = int(rng.integers(0, len(order)))
        ln = int(rng.integers(1, 10))
        chains.append([int(rng.choice(by_kind[order[(start + t) % len(order)]])) for t in range(ln)])
    return dict(templates=templates, consensuses=cons, reads=reads, segments=segments,
                seg_read=np.asarray(seg_read, dtype=np.int32), chains=chains)


# labels of the 39 templates in the order cyp2d6_templates returns them (src/cyp2d6/definitions.rs:346-464)
_HYBRID_PARTS = ["intron1", "exon2", "intron2", "exon3", "intron3", "exon4", "intron4", "exon5", "intron5", "exon6", "intron6",
                 "exon7", "intron7", "exon8", "intron8", "exon9"]


def cyp2d6_template_labels():
    labels = [("CYP2D6", None), ("CYP2D7", None), ("CYP2D6*5", None)]
    for part in _HYBRID_PARTS:
        labels.append(("Hybrid", f"CYP2D6::CYP2D7::{part}"))
        labels.append(("Hybrid", f"CYP2D7::CYP2D6::{part}"))
    return labels + [("REP6", None), ("REP7", None), ("spacer", None), ("link_region", None)]


def cyp2d6_diploid_sample(seed: int, n_reads: int = 96, star_alleles=("1.001", "4.001")):
    """One diploid CYP2D6 sample at the reference's region lengths: two haplotypes REP6 - CYP2D6*x - link - REP7 -
    spacer - CYP2D7 whose region copies differ by ~0.5 % (so reads assign uniquely), consensus sequences + labels for
    both, and HiFi-like reads spanning 2-4 consecutive regions of one haplotype with the region coordinates inside
    each read (what Cyp2d6Extractor + the consensus step hand to the chaining code, src/cyp2d6/caller.rs:430-537).
    Returns dict(templates, template_labels, consensuses, regions [(type, subtype, unique_id)], reads, roi)."""
    rng = np.random.default_rng([seed, 2, 66])
    regions, templates = cyp2d6_templates(rng)
    order = ["rep6", "d6", "link", "rep7", "spacer", "d7"]
    kinds = dict(rep6="REP6", d6="CYP2D6", link="link_region", rep7="REP7", spacer="spacer", d7="CYP2D7")
    cons, rows = [], []
    for hap, sub in enumerate(star_alleles):
        for kind in order:
            base = regions[kind]
            cons.append(mutate(rng, base, max(6, len(base) // 200), 1).tobytes())
            rows.append((kinds[kind], sub if kind == "d6" else None, len(rows)))
    reads, roi = [], {}
    for r in range(n_reads):
        hap, start, w = int(rng.integers(0, 2)), int(rng.integers(0, 5)), int(rng.integers(2, 5))
        lead = random_seq(rng, int(rng.integers(100, 400))).tobytes()
        pos, parts, regs = len(lead), [lead], []
        for t in range(start, min(start + w, 6)):
            seg, _ = hifi_reads(rng, [cons[hap * 6 + t]], 1, flank=0, lo=0, hi=1 << 30)
            regs.append((pos, pos + len(seg[0]), seg[0]))
            parts.append(seg[0])
            pos += len(seg[0])
        parts.append(random_seq(rng, int(rng.integers(100, 400))).tobytes())
        reads.append(b"".join(parts))
        roi[f"m84/{seed}/{r:04d}/ccs"] = regs
    return dict(templates=templates, template_labels=cyp2d6_template_labels(), consensuses=cons, regions=rows, reads=reads, roi=roi)
