"""Oracle-side restatement of the host flows (TEST INFRASTRUCTURE ONLY): the same call sequences as
pb_starphase_b200/host/*.cpp, but every alignment number comes from the CPU oracle (oracle/sp_oracle.c through
oracle_util.Oracle) and every piece of host logic from oracle/starphase_oracle.py.  The GPU tests compare the C++
host (GPU-backed) with these flows: identical integers, identical calls, byte-identical JSON."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

import starphase_oracle as so  # noqa: E402

DbRow = Tuple[str, str, List[str], Optional[str], str]  # hla_id, gene, star fields, dna, cdna


def dp_score(cigar) -> int:
    """minimap2 DP score under a=5 b=4 q=6 e=2 q2=26 e2=1 (src/hla/caller.rs:1370-1381)."""
    s = 0
    for ln, op in cigar:
        if op == 7:
            s += 5 * ln
        elif op == 8:
            s -= 4 * ln
        else:
            s -= min(6 + 2 * ln, 26 + ln)
    return s


# ---- K4 then K9, as GpuAligner::align_pairs(.., match_score) runs them ---------------------------------------
AFFINE_REFINE = True   # AlignerStandIns::affine_refine of the C++ host
REFINE_SLACK, REFINE_MAX_BAND = 24, 255
_AFFINE = {}


def dp_score_generic(cigar, match_score: int) -> int:
    s = 0
    for ln, op in cigar:
        s += match_score * ln if op == 7 else -4 * ln if op == 8 else -min(6 + 2 * ln, 26 + ln)
    return s


def align_scored(orc, pattern: bytes, text: bytes, match_score: int) -> dict:
    """The alignment a host flow works with: the unit-cost placement (K4), re-aligned under the reference's affine costs
    (a = match_score, b=4 q=6 e=2 q2=26 e2=1) inside a diagonal band around that placement (K9) when the band fits; `score` is the
    DP score compared with minimap2's -s floor.  An `orc` that already answers from the affine model is used as it is."""
    u = dict(orc.align(pattern, text))
    if "score" in u:
        return u
    u["score"] = dp_score_generic(u["cigar"], match_score)
    if not AFFINE_REFINE or not u["cigar"]:
        return u
    # the band: the diagonal hull of the unit-cost path (text column - pattern row along the CIGAR) widened by the slack
    d = lo = hi = u["t_start"] - u["p_start"]
    for ln, op in u["cigar"]:
        d += -ln if op == 1 else ln if op == 2 else 0
        lo, hi = min(lo, d), max(hi, d)
    w = (hi - lo + 1) // 2 + REFINE_SLACK
    if w > REFINE_MAX_BAND:
        return u
    import oracle_util

    aff = _AFFINE.setdefault(match_score, oracle_util.AffineOracle((match_score, 4, 6, 2, 26, 1)))
    band = 31 if w <= 31 else 63 if w <= 63 else 127 if w <= 127 else 255   # the band classes of the host
    a = dict(aff.align(pattern, text, centre=(lo + hi) // 2, band=band))  # floor: independent of the coordinate origin
    if a["score"] == 0:
        a["dist"] = len(pattern)
    return a


def mapping_from_alignment(a: dict, pattern_len: int, text_len: int, min_dp_score: int = 200) -> Optional[so.Mapping]:
    if not a["cigar"] or a["score"] < min_dp_score:
        return None
    return so.Mapping(a["p_start"], a["p_end"], pattern_len, a["t_start"], a["t_end"], text_len, a["nm"], True, a["cigar"])


def allowed_alleles(db: Sequence[DbRow], gene: str, require_dna: bool = True) -> List[DbRow]:
    return [r for r in sorted(db, key=lambda r: r[0].encode()) if r[1] == gene and (r[3] is not None or not require_dna)]


def score_read(orc, dna_target: bytes, cdna_target: bytes, db: Sequence[DbRow], gene: str, disable_cdna: bool = False):
    """src/hla/caller.rs:1332-1511 with the oracle's traceback alignment in place of minimap2."""
    stats, best = {}, so.HlaProcessedMatch.worst_match(2)
    for hla_id, _, star, dna, cdna in allowed_alleles(db, gene):
        cur = so.HlaProcessedMatch(hla_id)
        for target, seq in ((cdna_target, None if disable_cdna else cdna), (dna_target, dna)):
            m = None
            if seq is not None:
                cand = mapping_from_alignment(align_scored(orc, seq.encode(), target, 5), len(seq), len(target))
                idx, _ = so.select_best_mapping([cand] if cand else [], False, True)
                m = cand if idx is not None else None
            cur.add_mapping(m)
        st = [None if s is None else (s.seq_len, s.nm, s.unmapped) for s in cur.full_mapping_stats]
        if cur.is_better_match(best):
            best = cur
        stats[hla_id] = tuple(st)
    star = ""
    if best.haplotype:
        star = ":".join(next(r[2] for r in db if r[0] == best.haplotype))
    return stats, best.haplotype, star


def score_read_debug(orc, dna_target: bytes, cdna_target: bytes, db: Sequence[DbRow], gene: str, disable_cdna: bool = False) -> dict:
    """The ReadMappingStats score_read returns next to the scores (src/hla/caller.rs:1397, :1464-1477, :1502-1509, src/hla/debug.rs):
    per allele, keyed by star allele, the DetailedMappingStats of the best cDNA / DNA mapping."""
    per_allele, best = {}, so.HlaProcessedMatch.worst_match(2)
    for hla_id, _, star, dna, cdna in allowed_alleles(db, gene):
        cur, detail = so.HlaProcessedMatch(hla_id), []
        for target, seq in ((cdna_target, None if disable_cdna else cdna), (dna_target, dna)):
            m = None
            if seq is not None:
                cand = mapping_from_alignment(align_scored(orc, seq.encode(), target, 5), len(seq), len(target))
                idx, _ = so.select_best_mapping([cand] if cand else [], False, True)
                m = cand if idx is not None else None
            cur.add_mapping(m)
            detail.append(None if m is None else so.detailed_mapping_stats(m, target, seq.encode()))
        key = ":".join(star)
        if key in per_allele:
            raise ValueError(f"Entry {key} is already occupied!")
        per_allele[key] = tuple(detail)
        if cur.is_better_match(best):
            best = cur
    best_id = best.haplotype or None
    best_star = ":".join(next(r[2] for r in db if r[0] == best_id)) if best_id else None
    return so.read_mapping_stats_json(best_id, best_star, per_allele)


def hla_debug_json(orc, db: Sequence[DbRow], gene: str, consensuses: Sequence[Tuple[str, bytes, bytes]], is_dual: bool, counts1: int,
                   counts2: int) -> str:
    """hla_debug.json of one gene as diplotype_hla_batch fills it (src/hla/caller.rs:805, :877, :914)."""
    reads = {gene: {q: score_read_debug(orc, d, c, db, gene) for q, d, c in consensuses}}
    return so.hla_debug_json(reads, {gene: so.dual_passing_stats(is_dual, counts1, counts2)})


def score_consensus(orc, reference_sequence: bytes, ref_start: int, consensus: bytes, db: Sequence[DbRow], gene: str,
                    exons: Sequence[Tuple[int, int]], is_forward_strand: bool):
    """src/hla/caller.rs:1259-1319 on the oracle's alignment: consensus -> reference mapping -> soft-clipped CIGAR -> splice -> score_read.
    Returns (stats, best id, best star, ReadMappingStats JSON dict), all empty for an empty / unalignable consensus."""
    empty = ({}, "", "", so.read_mapping_stats_json(None, None, {}))
    if not consensus:
        return empty
    al = align_scored(orc, consensus, reference_sequence, 1)
    cand = []
    if al["cigar"] and al["score"] >= 200:
        cand.append(so.Mapping(al["p_start"], al["p_end"], len(consensus), al["t_start"], al["t_end"], len(reference_sequence), al["nm"], True,
                               al["cigar"]))
    if not cand:
        return empty
    idx, _ = so.select_best_mapping(cand, True, True)
    assert idx is not None
    m = cand[idx]
    cigar = ([(m.query_start, 4)] if m.query_start > 0 else []) + list(m.cigar) + ([(m.query_len - m.query_end, 4)] if m.query_len > m.query_end else [])
    dna_t, cdna_t = so.prepare_score_read_targets(consensus, ref_start + m.target_start, cigar, exons, is_forward_strand)
    stats, best_id, best_star = score_read(orc, dna_t, cdna_t, db, gene)
    return stats, best_id, best_star, score_read_debug(orc, dna_t, cdna_t, db, gene)


CANDIDATE_EDIT_WEIGHT = 5  # kCandidateEditWeight of pb_starphase_b200/host/sp_host_hla.cpp


def realign_records(orc, genes: Sequence[str], db: Sequence[DbRow], reads: Sequence[Tuple[str, bytes]], n_candidates: int = 5,
                    D: Optional[np.ndarray] = None) -> List[dict]:
    """src/hla/realigner.rs:98-211: candidates = the n alleles with the smallest 5 * (nm + unmapped) - |allele| (minimap2 ranks hits by
    alignment score: about aligned bases - 5 * edits under map-hifi), ties by database order."""
    alleles = [r for r in sorted(db, key=lambda r: r[0].encode()) if r[1] in genes and r[3] is not None]
    seqs = [r[3].encode() for r in alleles]
    if D is None:
        D = orc.score_batch([r[1] for r in reads], seqs) if alleles and reads else np.zeros((len(reads), 0), np.int32)
    out = []
    for r, (qname, seq) in enumerate(reads):
        best, best_a = so.MappingStats(len(seq), len(seq), 0), None
        order = sorted(range(len(alleles)), key=lambda a: (CANDIDATE_EDIT_WEIGHT * int(D[r, a]) - len(seqs[a]), a))[:max(n_candidates, 1)] if len(seq) else []
        for a in order:
            al = align_scored(orc, seqs[a], seq, 1)
            if not al["cigar"] or al["score"] < 200:  # db_aligner is the plain map-hifi preset (a = 1)
                continue
            tl = len(seqs[a])
            st = so.MappingStats(tl, al["nm"], tl - (al["p_end"] - al["p_start"]))
            if st.mapping_score() <= 0.5 and st.custom_score(False) <= 0.03 and st.custom_score(False) < best.custom_score(False):
                best, best_a = st, a
        hs = so.HlaMappingStats(None, best)
        if best_a is None:
            out.append(so.mapping_details_json(qname, "REFERENCE", "REFERENCE", hs, True))
        else:
            row = alleles[best_a]
            out.append(so.mapping_details_json(qname, row[0], f"{row[1]}*{':'.join(row[2])}", hs, False))
    return out


def realign_records_full(orc, genes: Sequence[str], db: Sequence[DbRow], gene_defs: Dict[str, Tuple[bool, bytes]],
                         reads: Sequence[Tuple[str, bytes]], n_candidates: int = 5) -> List[dict]:
    """HlaRealigner::new + realign_record in full (src/hla/realigner.rs:42-91, :98-350) on the oracle's alignments."""
    alleles = [r for r in sorted(db, key=lambda r: r[0].encode()) if r[1] in genes and r[3] is not None]
    seqs = [r[3].encode() if gene_defs[r[1]][0] else so.reverse_complement(r[3].encode()) for r in alleles]  # create_hla_fasta :511-515
    D = orc.score_batch([r[1] for r in reads], seqs) if alleles and reads else np.zeros((len(reads), 0), np.int32)
    out = []
    for r, (qname, seq) in enumerate(reads):
        best, best_a, best_al = so.MappingStats(len(seq), len(seq), 0), None, None
        order = sorted(range(len(alleles)), key=lambda a: (CANDIDATE_EDIT_WEIGHT * int(D[r, a]) - len(seqs[a]), a))[:max(n_candidates, 1)] if len(seq) else []
        for a in order:
            al = align_scored(orc, seqs[a], seq, 1)
            if not al["cigar"] or al["score"] < 200:
                continue
            tl = len(seqs[a])
            st = so.MappingStats(tl, al["nm"], tl - (al["p_end"] - al["p_start"]))
            if st.mapping_score() <= 0.5 and st.custom_score(False) <= 0.03 and st.custom_score(False) < best.custom_score(False):
                best, best_a, best_al = st, a, al
        hs = so.HlaMappingStats(None, best)
        if best_a is None:
            out.append(dict(gene_name="", read_mapping_stats=so.read_mapping_stats_json(None, None, {}),
                            mapping_details=so.mapping_details_json(qname, "REFERENCE", "REFERENCE", hs, True), realigned_record=None))
            continue
        row = alleles[best_a]
        gene, star = row[1], ":".join(row[2])
        fwd, ref = gene_defs[gene]
        # minimap2's view: query = read, target = allele -> roles and I / D letters swap relative to the oracle's (pattern, text)
        swap = {1: 2, 2: 1}
        bm = so.Mapping(best_al["t_start"], best_al["t_end"], len(seq), best_al["p_start"], best_al["p_end"], len(seqs[best_a]), best_al["nm"],
                        True, [(ln, swap.get(op, op)) for ln, op in best_al["cigar"]])
        rms = so.read_mapping_stats_json(row[0], star, {row[0]: (None, so.detailed_mapping_stats(bm, seqs[best_a], seq))})
        res = dict(gene_name=gene, read_mapping_stats=rms,
                   mapping_details=so.mapping_details_json(qname, row[0], f"{gene}*{star}", hs, False), realigned_record=None)
        out.append(res)
        db_start, db_end = bm.query_start, bm.query_end
        bs, be = max(db_start - 1000, 0), min(db_end + 1000, len(seq))
        sal = align_scored(orc, seq[bs:be], ref, 1)
        cand = []
        if sal["cigar"] and sal["score"] >= 200:
            cand.append(so.Mapping(sal["p_start"], sal["p_end"], be - bs, sal["t_start"], sal["t_end"], len(ref), sal["nm"], True, sal["cigar"]))
        idx, _ = so.select_best_mapping(cand, True, True)
        if idx is None:
            continue
        rm = cand[idx]
        hg_start, hg_end = bs + rm.query_start, bs + rm.query_end
        seg = (min(db_start, hg_start), max(db_end, hg_end))
        d = rm.target_start
        h = so.hpc_pos(ref, d)
        if not hg_start < db_start:
            aal = align_scored(orc, seqs[best_a], ref, 1)
            acand = []
            if aal["cigar"] and aal["score"] >= 200:
                acand.append(so.Mapping(aal["p_start"], aal["p_end"], len(seqs[best_a]), aal["t_start"], aal["t_end"], len(ref), aal["nm"], True,
                                        aal["cigar"]))
            aidx, _ = so.select_best_mapping(acand, False, True)
            if aidx is not None:
                am = acand[aidx]
                added = max(am.target_start - am.query_start, 0)
                d = added + bm.target_start
                h = so.hpc_pos(ref, added) + so.hpc_pos(row[3].encode(), bm.target_start)  # the stored (gene-strand) allele, as :312 does
        dna = seq[seg[0]:seg[1]]
        res["realigned_record"] = (seg[0], seg[1], d, h, dna, so.hpc(dna))
    return out


def diplotype_hla_gene(orc, db: Sequence[DbRow], gene: str, reads: Sequence[Tuple[str, bytes, bytes]]) -> dict:
    """north_star (2) + src/hla/caller.rs:889-901, :1046-1065."""
    al = allowed_alleles(db, gene)
    Dd = orc.score_batch([r[1] for r in reads], [a[3].encode() for a in al])
    Dc = orc.score_batch([r[2] for r in reads], [a[4].encode() for a in al])
    score, score2, i, j, c1 = orc.pair_minsum_topk(Dc, 10, D2=Dd)[0]
    c2 = len(reads) - c1
    id1, id2 = al[i][0], al[j][0]
    if i == j:
        ids = (id1, id1)
    else:
        ids = so.choose_diplotype(id1, id2, c1, c2)
    star = {a[0]: "*" + ":".join(a[2]) for a in al}
    details = realign_records(orc, [gene], al, [(r[0], r[1]) for r in reads], D=Dd)
    gd = so.gene_details_from_mappings([so.diplotype_json(star[ids[0]], star[ids[1]])], details)
    return dict(hla_id1=ids[0], hla_id2=ids[1], counts1=c1, counts2=c2, pair_score_cdna=score, pair_score_dna=score2, gene_details=gd)


# ---- CYP2D6 -----------------------------------------------------------------------------------------------
def labels_from_rows(rows) -> List[so.RegionLabel]:
    return [so.RegionLabel(t, s) for t, s, _ in rows]


def index_label(row) -> str:
    lab = so.RegionLabel(row[0], row[1])
    return (f"{row[2]}" if row[2] is not None else "X") + "_" + lab.full_allele()


NO_MAPPING_PERMILLE = 350  # unit-cost distance above which a (segment, consensus) pair counts as "no mapping"


def weight_sequences(orc, segments: Sequence[bytes], consensuses: Sequence[bytes], labels) -> List[list]:
    """src/cyp2d6/chaining.rs:28-103 per segment; one hit per consensus unless nothing of the segment aligns."""
    if not segments:
        return []
    D, S, E = orc.score_spans(consensuses, segments)  # [consensus (text)][segment (pattern)]
    out = []
    for s, seg in enumerate(segments):
        hits = []
        for k, con in enumerate(consensuses):
            if len(con) == 0 or len(seg) == 0 or D[k, s] >= len(seg) or int(D[k, s]) * 1000 > len(seg) * NO_MAPPING_PERMILLE:
                hits.append([])  # nothing aligns, or so far apart that an aligner reports no mapping
            else:
                hits.append([(int(D[k, s]), int(S[k, s]), len(con) - int(E[k, s]), len(con))])
        out.append(so.weight_sequence_from_hits(len(seg), labels, hits))
    return out


def call_cyp2d6_chains(orc, consensuses: Sequence[bytes], rows, roi: Dict[str, List[Tuple[int, int, bytes]]], infer: bool,
                       normalize_all: bool) -> dict:
    """The chaining half of diplotype_cyp2d6, src/cyp2d6/caller.rs:430-739."""
    labels = labels_from_rows(rows)
    cfg = so.Cyp2d6Config.default()
    qnames = sorted(roi)
    segs = [r[2] for q in qnames for r in roi[q]]
    ws = weight_sequences(orc, segs, consensuses, labels)
    rw, k = {}, 0
    for q in qnames:
        rw[q] = ws[k:k + len(roi[q])]
        k += len(roi[q])
    chains, scores, counts = so.build_chains(rw, len(labels))
    mm = []
    for q in sorted(chains):
        if len(chains[q]) == 1:
            for con, reg in zip(chains[q][0], roi[q]):
                mm.append(dict(read_qname=q, read_position=dict(start=reg[0], end=reg[1]), consensus_id=con,
                               consensus_star_allele=index_label(rows[con])))
    rows = list(rows)
    for c, cnt in enumerate(counts):
        if cnt == 0 and labels[c].region_type not in (so.UNKNOWN, so.FALSE_ALLELE):
            labels[c] = so.RegionLabel(so.FALSE_ALLELE, labels[c].subtype_label)
            rows[c] = (so.FALSE_ALLELE, rows[c][1], rows[c][2])
    best, danglers, dbg = so.find_best_chain_pair(cfg, chains, scores, labels, infer, normalize_all, so.ChainPenalties(), False,
                                                  return_debug=True)
    uids = [r[2] for r in rows]

    def hap(k, level):
        return so.convert_chain_to_hap(best[k], labels, level, cfg.cyp_translate, uids)

    deep = dict(basic_diplotype=so.diplotype_json(hap(0, so.DEEP), hap(1, so.DEEP)), haplotype_1=None, haplotype_2=None)
    gd = so.gene_details_from_multi_mappings([so.diplotype_json(hap(0, so.SUB), hap(1, so.SUB))],
                                             [so.diplotype_json(hap(0, so.CORE), hap(1, so.CORE))], [deep], mm)
    return dict(best_chains=best, score=dbg["best"]["score"], dangling=danglers, n_possible_chains=len(dbg["possible_chains"]),
                gene_details=gd)


# ---- CYP2D6 template search: src/cyp2d6/haplotyper.rs:142-315 -------------------------------------------------
def dp_score_a1(cigar) -> int:
    """minimap2 DP score under the map-hifi defaults of standard_hifi_aligner (a=1 b=4 q=6 e=2 q2=26 e2=1)."""
    s = 0
    for ln, op in cigar:
        if op == 7:
            s += ln
        elif op == 8:
            s -= 4 * ln
        else:
            s -= min(6 + 2 * ln, 26 + ln)
    return s


def hla_consensus_step(records, min_consensus_count: int = 3, min_consensus_fraction: float = 0.10, min_cdf: float = 0.001,
                       expected_maf: float = 0.45):
    """run_dual_consensus_with_offsets (src/hla/caller.rs:1151-1219) + the per-group re-consensus (:706-760) on the oracle's search
    (oracle/consensus_oracle.py).  records: (qname, dna_sequence, hpc_sequence, dna_offset, hpc_offset); taken in qname order.
    Returns (dual consensus dict, is_passing, (consensus 1, consensus 2 or None))."""
    import consensus_oracle as co

    recs = sorted(records, key=lambda r: r[0].encode())
    cfg = co.Config(min_count=min_consensus_count, min_af=min_consensus_fraction, allow_early_termination=True, max_queue_size=20,
                    max_capacity_per_size=10, offset_window=400)
    half = cfg.offset_window // 2

    def offsets(vals):
        lo = min(vals)
        return [None if v == lo else v - lo + half for v in vals]

    def passing(d):
        if d["consensus2"] is None:
            return False
        c1 = sum(d["is_consensus1"])
        return so.is_passing_dual(c1, len(d["is_consensus1"]) - c1, min_consensus_fraction, min_cdf, expected_maf)

    d = co.dual_consensus([r[2] for r in recs], offsets([r[4] for r in recs]), cfg)[0]  # HPC first
    if not passing(d):
        d = co.dual_consensus([r[1] for r in recs], offsets([r[3] for r in recs]), cfg)[0]
    groups = []
    for want in (True, False):
        sel = [r for r, first in zip(recs, d["is_consensus1"]) if first == want]
        if not want and d["consensus2"] is None:
            groups.append(None)
        elif not sel:
            groups.append(b"")
        else:
            groups.append(co.consensus([r[1] for r in sel], offsets([r[3] for r in sel]), cfg)[0][0])
    return d, passing(d), tuple(groups)


def find_full_type_in_sequences(orc, templates, seqs: Sequence[bytes], max_missing_frac: float, force_assignment: bool, db: dict,
                                graph_band: int = 128):
    """find_full_type_in_sequence + assign_haplotype (src/cyp2d6/haplotyper.rs:326-361, :371-601) as
    Cyp2d6Extractor::find_full_type_in_sequences runs them.  db: backbone (bytes), backbone_start, variants [(pos, ref, alt)],
    metadata [(label, is_vi)], haplotype_lookup {star: 0/1 list}, mapped_hybrids [(type, subtype)].
    Per sequence: None ("no matches found") or ((region type, subtype), region variants or None)."""
    import graph_oracle as go

    by_name = {so.RegionLabel(t[0], t[1]).full_allele(): (t[0], t[1]) for t in templates}
    out = []
    for seq, hits in zip(seqs, find_base_type_in_sequences(orc, templates, seqs, max_missing_frac)):
        if not hits:
            out.append(None)
            continue
        best = min(hits, key=lambda h: so.MappingStats(*h[3]).custom_score(True))  # the first of equal minima, like min_by
        label = by_name[best[0]]
        if label not in [tuple(x) for x in db["mapped_hybrids"]]:
            out.append((label, None))
            continue
        a = align_scored(orc, seq, db["backbone"], 1)  # the consensus (query) on the backbone (target)
        assert a["cigar"] and a["score"] >= 200
        ts, te = a["t_start"], a["t_end"]
        g = go.build_graph(db["backbone"][ts:te], db["backbone_start"] + ts, db["variants"])
        score, nodes = go.align(g, seq[a["p_start"]:a["p_end"]], band=graph_band)
        assert score < go.INF
        vec = so.alleles_from_traversal(len(db["variants"]), nodes, g.node_to_alleles)
        star, rv, _ = so.assign_haplotype_from_alleles(vec, db["haplotype_lookup"], [m[0] for m in db["metadata"]],
                                                       [m[1] for m in db["metadata"]], force_assignment)
        out.append(((so.UNKNOWN, None) if star is None else (so.CYP2D6, star), rv))
    return out


def find_base_type_in_sequences(orc, templates, seqs: Sequence[bytes], max_missing_frac: float):
    """templates: [(region_type, subtype, sequence bytes)].  Same search as Cyp2d6Extractor::find_base_type_in_sequences
    in pb_starphase_b200/host/sp_host_cyp2d6.cpp, but every traceback runs over the whole (sub)segment with the oracle's
    full-matrix DP instead of the GPU's placement window."""
    PEN = (so.DELETION, so.REP6, so.REP7)
    tl = sorted(templates, key=lambda t: so.RegionLabel(t[0], t[1]).full_allele())  # stable, like the host
    tseqs = [t[2] for t in tl]
    labels = [so.RegionLabel(t[0], t[1]) for t in tl]
    out = []
    for s, seq in enumerate(seqs):
        hits, n_hits = [], [0] * len(tl)
        items = [(0, len(seq), t) for t in range(len(tl))] if len(seq) else []
        for _round in range(5):
            nxt = []
            for lo, hi, t in items:
                m = len(tseqs[t])
                d, _ = orc.infix(tseqs[t], seq[lo:hi])  # the K1 prefilter of every round
                if m == 0 or 2 * d > m:
                    continue
                a = align_scored(orc, tseqs[t], seq[lo:hi], 1)
                if not a["cigar"] or a["score"] < 200:
                    continue
                st = so.MappingStats(m, a["nm"], m - (a["p_end"] - a["p_start"]), a["p_start"], m - a["p_end"])
                if st.custom_score(labels[t].region_type in PEN) > 0.05:
                    continue
                hs, he = lo + a["t_start"], lo + a["t_end"]
                hits.append((hs, he, st, t))
                n_hits[t] += 1
                if n_hits[t] >= 5:
                    continue
                min_len = max(int(0.9 * (1.0 - min(max_missing_frac, 1.0)) * float(m)), 200)
                if hs > lo and hs - lo >= min_len:
                    nxt.append((lo, hs, t))
                if hi > he and hi - he >= min_len:
                    nxt.append((he, hi, t))
            items = nxt
            if not items:
                break
        hits.sort(key=lambda h: (h[3], h[2].custom_score(True)))
        hits.sort(key=lambda h: (h[0], h[1]))
        regions, cur = [], None
        for h in hits:
            if cur is None:
                cur = h
                continue
            mn, mx = min(h[1], cur[1]), max(h[0], cur[0])
            ov = 0.0 if mx >= mn else float(mn - mx) / min(float(h[1] - h[0]), float(cur[1] - cur[0]))
            if ov > 0.9:
                pen = labels[h[3]].region_type in PEN or labels[cur[3]].region_type in PEN
                hp = 1 if labels[h[3]].region_type == so.DELETION else 0
                cp = 1 if labels[cur[3]].region_type == so.DELETION else 0
                if (h[2].custom_score(pen) < cur[2].custom_score(pen) and hp >= cp) or hp > cp:
                    cur = h
            else:
                regions.append(cur)
                cur = h
        if cur is not None:
            regions.append(cur)
        out.append([(labels[t].full_allele(), hs, he, (st.seq_len, st.nm, st.unmapped, st.clipped_start, st.clipped_end))
                    for hs, he, st, t in regions if not st.custom_score(True) > max_missing_frac])
    return out


# ---- the consensus stage of the CYP2D6 caller (src/cyp2d6/caller.rs:145-310, :750-893) ------------------------------------
def hpc_with_guide(sequence: bytes, guide_sequence: bytes, guide_offset: int):
    """src/util/homopolymers.rs:53-64."""
    return so.hpc(sequence), so.hpc_pos(guide_sequence, guide_offset)


_SEEDS = {so.DELETION: 0, so.REP6: 1, so.REP7: 2, so.SPACER: 3, so.LINK: 4}  # src/cyp2d6/caller.rs:224-231


def cyp2d6_consensus_inputs(read_sequences: Dict[str, bytes], roi: Dict[str, list], templates, max_missing_consensus_frac: float,
                            offset_window: int = 50) -> dict:
    """The loop at src/cyp2d6/caller.rs:168-212 (+ the seeds of :224-231).  roi: {read id: [(type, subtype, start, end, stats 5-tuple)]}."""
    by_label = {(t[0], t[1]): t[2] for t in templates}
    out = dict(raw_sequences=[], hpc_sequences=[], base_offsets=[], hpc_offsets=[], sequence_ids=[], seeds=[])
    for read_id in sorted(roi):
        for t, sub, start, end, st in roi[read_id]:
            stats = so.MappingStats(*st)
            if stats.custom_score(True) > max_missing_consensus_frac:
                continue
            prefix = st[3] or 0
            seq = read_sequences[read_id][start:end]
            hp, hp_off = hpc_with_guide(seq, by_label[(t, sub)], prefix)
            out["raw_sequences"].append(seq)
            out["base_offsets"].append(0 if prefix == 0 else prefix + offset_window)
            out["hpc_sequences"].append(hp)
            out["hpc_offsets"].append(0 if hp_off == 0 else hp_off + offset_window)
            out["sequence_ids"].append(f"{read_id}_{start}_{end}_{so.RegionLabel(t, sub).full_allele()}")
            out["seeds"].append(_SEEDS.get(t))
    return out


def merge_consensus_results(orc, sequences: Sequence[bytes], offsets: Sequence[int], cfg, raw_consensuses, raw_indices: Sequence[int],
                            templates, db: dict, max_missing_consensus_frac: float):
    """merge_consensus_results (src/cyp2d6/caller.rs:750-893) on oracle numbers: typing through find_full_type_in_sequences above,
    merged groups through consensus_oracle.consensus.  raw_consensuses[group] = [(hpc sequence, scores), (full sequence, scores)].
    Returns (consensuses [(sequence, scores)], sequence_indices)."""
    import consensus_oracle as co

    cyp = so.Cyp2d6Config.default()
    to_type = [levels[1][0].strip(b"*") for levels in raw_consensuses]
    typed = find_full_type_in_sequences(orc, templates, to_type, max_missing_consensus_frac, False, db)
    consensus_set: Dict[Tuple[bytes, str], List[int]] = {}
    unknown_set: Dict[bytes, List[int]] = {}
    unknown = so.RegionLabel(so.UNKNOWN).full_allele()
    for i, levels in enumerate(raw_consensuses):
        label = so.RegionLabel(*typed[i][0]) if typed[i] is not None else so.RegionLabel(so.UNKNOWN)
        if not label.is_allowed_label():
            unknown_set.setdefault(levels[0][0], []).append(i)
        else:
            consensus_set.setdefault((levels[0][0], label.simplify_allele(True, cyp.cyp_translate)), []).append(i)
    ignore = set()
    for hp in sorted(unknown_set):
        others = [k for k in sorted(consensus_set) if k[0] == hp]
        if len(others) == 1:
            consensus_set[others[0]].extend(unknown_set[hp])
        else:
            if len(others) > 1:
                ignore.add((hp, unknown))
            assert (hp, unknown) not in consensus_set
            consensus_set[(hp, unknown)] = unknown_set[hp]
    cons, idx = [], [None] * len(raw_indices)
    for key in sorted(consensus_set, key=lambda k: (k[0], k[1].encode())):
        members, ci = consensus_set[key], len(cons)
        if key in ignore:
            n = 0
            for i, si in enumerate(raw_indices):
                if si in members:
                    idx[i] = ci
                    n += 1
            cons.append((b"", [0] * n))
        elif len(members) == 1:
            for i, si in enumerate(raw_indices):
                if si == members[0]:
                    idx[i] = ci
            cons.append((raw_consensuses[members[0]][1][0], list(raw_consensuses[members[0]][1][1])))
        else:
            reads, offs = [], []
            for si, (seq, off) in enumerate(zip(sequences, offsets)):
                if raw_indices[si] in members:
                    reads.append(seq)
                    offs.append(None if off == 0 else off)
                    idx[si] = ci
            first = co.consensus(reads, offs, cfg)[0]
            cons.append((first[0], list(first[1])))
    assert all(v is not None for v in idx)
    return cons, idx
