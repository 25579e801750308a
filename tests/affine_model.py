"""Affine-gap reference MODEL (TEST INFRASTRUCTURE ONLY): the flows of tests/flow_oracle.py re-run with every alignment number
taken from oracle/sp_oracle_affine.c -- the best local alignment under minimap2's two-piece affine costs (map-hifi: a=1 b=4
q=6 e=2 q2=26 e2=1, src/util/mapping.rs:8-14; a=5 in score_read, src/hla/caller.rs:1370-1381) -- instead of the unit-cost
quantities the product computes (K1's infix distance, K4's canonical path).

It answers "how far are the product's numbers from the reference's cost model": per pair (`pair_divergence`), per HLA call
(`score_read_flips`, `diplotype_flips`, `realign_flips`) and per CYP2D6 decision (`weight_sequence_flips`), plus how the three
product-side heuristics that stand in for "minimap2 reported nothing / these five hits" compare with the model (the 35 % cut of
weight_sequences, the -s 200 floor evaluated on K4's CIGAR, the bases-explained ranking of the realigner's candidates).
It says nothing about minimap2's seeding, chaining and z-drop: parity with the real reference stays unpinned."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))

import flow_oracle as fo  # noqa: E402
import oracle_util  # noqa: E402
import starphase_oracle as so  # noqa: E402

MIN_DP_SCORE = 200  # minimap2 -s for map-hifi


class AffineFlowOracle:
    """Looks like oracle_util.Oracle to the flows, answers from the affine model.  A pair whose best local score is below
    minimap2's -s 200 floor has no mapping: its distance is |P| (nothing of the pattern explained) and its CIGAR is empty."""

    def __init__(self, base: oracle_util.Oracle, costs, nthreads: int = 0):
        self.base, self.aff, self.nthreads = base, oracle_util.AffineOracle(costs), nthreads
        self.memo = {}  # (pattern, text) -> normalised record of matrix() calls (CIGAR only when it was asked for)

    def _norm(self, r: dict, m: int) -> dict:
        if r["score"] < MIN_DP_SCORE:
            return {"dist": m, "nm": 0, "p_start": 0, "p_end": 0, "t_start": 0, "t_end": 0, "cigar": [], "score": r["score"]}
        return r

    def align(self, pattern: bytes, text: bytes) -> dict:
        return self._norm(self.aff.align(pattern, text), len(pattern))

    def infix(self, pattern: bytes, text: bytes, prefix: bool = False, impl: str = "dp"):
        r = self.align(pattern, text)
        return r["dist"], r["t_end"]

    def matrix(self, targets: Sequence[bytes], patterns: Sequence[bytes], want_cigar: bool = False):
        pairs = [(t, p) for t in range(len(targets)) for p in range(len(patterns))]
        recs = self.aff.align_batch(list(targets), list(patterns), pairs, self.nthreads, want_cigar=want_cigar) if pairs else []
        recs = [self._norm(r, len(patterns[p])) for r, (_, p) in zip(recs, pairs)]
        for r, (t, p) in zip(recs, pairs):
            key = (bytes(patterns[p]), bytes(targets[t]))
            self.memo[key] = r
            if want_cigar:
                self.aff.cache[key] = r
        return recs

    def has_mapping(self, pattern: bytes, text: bytes) -> bool:
        r = self.memo.get((bytes(pattern), bytes(text)))
        return (r if r is not None else self.align(pattern, text))["score"] >= MIN_DP_SCORE

    def score_batch(self, targets, patterns, **_kw):
        recs = self.matrix(targets, patterns)
        return np.asarray([r["dist"] for r in recs], dtype=np.int32).reshape(len(targets), len(patterns))

    def score_spans(self, targets, patterns, nthreads: int = 0):
        recs = self.matrix(targets, patterns)
        shape = (len(targets), len(patterns))
        return tuple(np.asarray([r[k] for r in recs], dtype=np.int32).reshape(shape) for k in ("dist", "t_start", "t_end"))

    def pair_minsum_topk(self, *a, **kw):
        return self.base.pair_minsum_topk(*a, **kw)


def pair_divergence(orc: oracle_util.Oracle, costs, targets: Sequence[bytes], patterns: Sequence[bytes],
                    pairs: Sequence[Tuple[int, int]], nthreads: int = 0) -> dict:
    """Per (target, pattern) pair: the product's (dist, nm, unmapped, span) -- K1 / K4 == the unit-cost oracle, bit for bit
    (tests/test_k1_gpu.py, test_k4_gpu.py) -- against the affine model's."""
    aff = oracle_util.AffineOracle(costs).align_batch(list(targets), list(patterns), list(pairs), nthreads, want_cigar=False)
    out = dict(pairs=len(pairs), dist_differs=0, dist_model_below_product=0, split_differs=0, span_differs=0, no_mapping_in_model=0,
               max_dist_gap=0, sum_dist_gap=0, hist={})
    for (t, p), a in zip(pairs, aff):
        u = orc.align(patterns[p], targets[t])
        m = len(patterns[p])
        if a["score"] < MIN_DP_SCORE:
            out["no_mapping_in_model"] += 1
            continue
        gap = a["dist"] - u["dist"]
        out["dist_differs"] += gap != 0
        out["dist_model_below_product"] += gap < 0
        out["max_dist_gap"] = max(out["max_dist_gap"], gap)
        out["sum_dist_gap"] += gap
        out["hist"][gap] = out["hist"].get(gap, 0) + 1
        out["split_differs"] += (a["nm"], m - (a["p_end"] - a["p_start"])) != (u["nm"], m - (u["p_end"] - u["p_start"]))
        out["span_differs"] += (a["t_start"], a["t_end"], a["p_start"], a["p_end"]) != (u["t_start"], u["t_end"], u["p_start"], u["p_end"])
    out["hist"] = {str(k): v for k, v in sorted(out["hist"].items())}
    return out


def score_read_flips(orc, aff_orc: AffineFlowOracle, db, gene: str, targets: Sequence[Tuple[bytes, bytes]]) -> dict:
    """score_read (src/hla/caller.rs:1332-1511) per consensus-like target on product numbers and on model numbers: how often the
    best allele, or the (len, nm, unmapped) that would be written to the result JSON for it, changes."""
    out = dict(targets=len(targets), best_allele_differs=0, best_stats_differ=0, alleles_with_different_stats=0, alleles=0,
               floor_disagrees=0)
    al = fo.allowed_alleles(db, gene)
    for dna_t, cdna_t in targets:
        aff_orc.aff.prefetch([r[3].encode() for r in al if r[3]], dna_t, aff_orc.nthreads)
        aff_orc.aff.prefetch([r[4].encode() for r in al], cdna_t, aff_orc.nthreads)
        su, bu, _ = fo.score_read(orc, dna_t, cdna_t, db, gene)
        sa, ba, _ = fo.score_read(aff_orc, dna_t, cdna_t, db, gene)
        out["best_allele_differs"] += bu != ba
        out["best_stats_differ"] += bu == ba and bu != "" and su[bu] != sa[ba]
        out["alleles"] += len(su)
        out["alleles_with_different_stats"] += sum(su[k] != sa[k] for k in su)
        # the -s 200 floor, evaluated by the product on K4's CIGAR and by the model on its own optimum
        out["floor_disagrees"] += sum((su[k][i] is None) != (sa[k][i] is None) for k in su for i in (0, 1))
    return out


def diplotype_flips(orc, aff_orc: AffineFlowOracle, db, gene: str, samples: Sequence[Sequence[Tuple[str, bytes, bytes]]]) -> dict:
    """north_star (2): best allele pair by sum_r min(D[r,i], D[r,j]) with the (cDNA, DNA) key, product vs model distances."""
    out = dict(samples=len(samples), pair_differs=0, hom_het_differs=0, cells=0, cells_differ=0)
    al = fo.allowed_alleles(db, gene)
    for reads in samples:
        res = []
        for o in (orc, aff_orc):
            Dd = o.score_batch([r[1] for r in reads], [a[3].encode() for a in al])
            Dc = o.score_batch([r[2] for r in reads], [a[4].encode() for a in al])
            _, _, i, j, c1 = orc.pair_minsum_topk(Dc, 10, D2=Dd)[0]
            ids = (al[i][0], al[j][0]) if i == j else so.choose_diplotype(al[i][0], al[j][0], c1, len(reads) - c1)
            res.append((i, j, ids, Dd, Dc))
        out["pair_differs"] += res[0][:2] != res[1][:2]
        out["hom_het_differs"] += res[0][2] != res[1][2]
        out["cells"] += res[0][3].size + res[0][4].size
        out["cells_differ"] += int((res[0][3] != res[1][3]).sum() + (res[0][4] != res[1][4]).sum())
    return out


def realign_flips(orc, aff_orc: AffineFlowOracle, db, genes, reads: Sequence[Tuple[str, bytes]], n_candidates: int = 5) -> dict:
    """HlaRealigner::realign_record (src/hla/realigner.rs:98-211).  Product: candidates = the five alleles with the most bases
    explained by K1, K4 on those, the reference's thresholds.  Model: the five hits with the highest affine DP score
    (minimap2 ranks its hits by score), the same thresholds."""
    alleles = [r for r in sorted(db, key=lambda r: r[0].encode()) if r[1] in genes and r[3] is not None]
    seqs = [r[3].encode() for r in alleles]
    prod = fo.realign_records(orc, genes, db, reads, n_candidates)
    recs = aff_orc.matrix([r[1] for r in reads], seqs)
    out = dict(reads=len(reads), assignment_differs=0, stats_differ=0, model_best_not_in_product_candidates=0)
    D = orc.score_batch([r[1] for r in reads], seqs)
    for r, (qname, seq) in enumerate(reads):
        row = recs[r * len(seqs):(r + 1) * len(seqs)]
        order = sorted(range(len(seqs)), key=lambda a: (-row[a]["score"], a))[:n_candidates]
        best, best_a = so.MappingStats(len(seq), len(seq), 0), None
        for a in order:
            al = row[a]
            if al["score"] < MIN_DP_SCORE:
                continue
            tl = len(seqs[a])
            st = so.MappingStats(tl, al["nm"], tl - (al["p_end"] - al["p_start"]))
            if st.mapping_score() <= 0.5 and st.custom_score(False) <= 0.03 and st.custom_score(False) < best.custom_score(False):
                best, best_a = st, a
        model_id = alleles[best_a][0] if best_a is not None else "REFERENCE"
        out["assignment_differs"] += prod[r]["best_hla_id"] != model_id
        ps = prod[r]["best_mapping_stats"]["dna_stats"]
        out["stats_differ"] += prod[r]["best_hla_id"] == model_id and (ps["seq_len"], ps["nm"], ps["unmapped"]) != (best.seq_len, best.nm, best.unmapped)
        cand = sorted(range(len(seqs)), key=lambda a: (fo.CANDIDATE_EDIT_WEIGHT * int(D[r, a]) - len(seqs[a]), a))[:n_candidates]
        out["model_best_not_in_product_candidates"] += best_a is not None and best_a not in cand
    return out


def weight_sequence_flips(orc, aff_orc: AffineFlowOracle, segments: Sequence[bytes], consensuses: Sequence[bytes], labels) -> dict:
    """weight_sequence (src/cyp2d6/chaining.rs:28-103): product = K3 spans + the 35 % "no mapping" cut; model = a hit exists iff
    the affine local score reaches minimap2's -s 200."""
    prod = fo.weight_sequences(orc, segments, consensuses, labels)
    saved = fo.NO_MAPPING_PERMILLE
    fo.NO_MAPPING_PERMILLE = 1000
    try:
        model = fo.weight_sequences(aff_orc, segments, consensuses, labels)
    finally:
        fo.NO_MAPPING_PERMILLE = saved
    out = dict(segments=len(segments), emptiness_differs=0, argmin_set_differs=0, ed_differs=0, overlap_differs=0,
               cut_35pct_hides_model_hit=0, cut_35pct_keeps_model_miss=0, pairs=len(segments) * len(consensuses))
    D, _, _ = orc.score_spans(consensuses, segments)
    for s, (wp, wm) in enumerate(zip(prod, model)):
        if (not wp) != (not wm):
            out["emptiness_differs"] += 1
            continue
        if not wp:
            continue
        mp, mm = min(w[0] for w in wp), min(w[0] for w in wm)
        out["argmin_set_differs"] += [k for k, w in enumerate(wp) if w[0] == mp] != [k for k, w in enumerate(wm) if w[0] == mm]
        out["ed_differs"] += sum(a[0] != b[0] for a, b in zip(wp, wm))
        out["overlap_differs"] += sum(a[0] == b[0] and a[1] != b[1] for a, b in zip(wp, wm))
    for s, seg in enumerate(segments):
        for k, con in enumerate(consensuses):
            cut = int(D[k, s]) * 1000 > len(seg) * saved or D[k, s] >= len(seg)
            has = aff_orc.has_mapping(seg, con)
            out["cut_35pct_hides_model_hit"] += cut and has
            out["cut_35pct_keeps_model_miss"] += (not cut) and not has
    return out
