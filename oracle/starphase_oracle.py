"""oracle/starphase_oracle.py -- CPU ORACLE (Python half).  TEST INFRASTRUCTURE ONLY.

Restates the reference's own (in-tree, pure Rust) integer / float logic that sits around the
alignment arithmetic on the hot path, each function citing the reference file:line it follows
(paths relative to /root/reference, pb-StarPhase v2.0.1).  These pieces ARE pinned: the
known-answer vectors of the reference's unit tests are re-expressed in tests/test_host_logic_cpu.py.

The alignment numbers themselves (minimap2 `nm`, spans) are "parity unpinned" -- see the header of
oracle/sp_oracle.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; nothing under pb_starphase_b200/ does.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

# ==========================================================================================
# Scores -- src/data_types/mapping.rs, src/hla/mapping.rs
# ==========================================================================================
WORST_SCORE = 1.0  # MappingScore::worst_value, src/data_types/mapping.rs:140-142


def score_value(mapping_len: int, nm: int, unmapped: int) -> float:
    """MappingScore::score_value, src/data_types/mapping.rs:191-195."""
    return max(float(nm + unmapped), 0.1) / float(mapping_len)


@dataclass
class MappingStats:
    """src/data_types/mapping.rs:7-22."""
    seq_len: int
    nm: int
    unmapped: int
    clipped_start: Optional[int] = None
    clipped_end: Optional[int] = None

    def custom_score(self, penalize_unmapped: bool) -> float:
        """src/data_types/mapping.rs:69-85."""
        if penalize_unmapped:
            return score_value(self.seq_len, self.nm, self.unmapped)
        return score_value(self.seq_len - self.unmapped, self.nm, 0)

    def mapping_score(self) -> float:
        return self.custom_score(True)

    def to_json(self) -> dict:
        return dict(seq_len=self.seq_len, nm=self.nm, unmapped=self.unmapped,
                    clipped_start=self.clipped_start, clipped_end=self.clipped_end)


@dataclass
class HlaMappingStats:
    """src/hla/mapping.rs:9-14; score = (cDNA, DNA) compared lexicographically (:111-117)."""
    cdna_stats: Optional[MappingStats] = None
    dna_stats: Optional[MappingStats] = None

    def mapping_score(self) -> Tuple[float, float]:
        c = self.cdna_stats.mapping_score() if self.cdna_stats is not None else WORST_SCORE
        d = self.dna_stats.mapping_score() if self.dna_stats is not None else WORST_SCORE
        return (c, d)

    def to_json(self) -> dict:
        return dict(cdna_stats=self.cdna_stats.to_json() if self.cdna_stats else None,
                    dna_stats=self.dna_stats.to_json() if self.dna_stats else None)


def harmonic_mean(scores: Sequence[float]) -> float:
    """MappingScore::harmonic_mean, src/data_types/mapping.rs:166-178."""
    total = 0.0
    for s in scores:
        assert s > 0.0
        total += 1.0 / s
    return len(scores) / total if total > 0.0 else 0.0


@dataclass
class Mapping:
    """The fields of minimap2::Mapping the reference reads."""
    query_start: int
    query_end: int
    query_len: int
    target_start: int
    target_end: int
    target_len: int
    nm: int
    forward: bool = True
    cigar: Optional[List[Tuple[int, int]]] = None  # (length, op) with minimap2 op codes


def select_best_mapping(mappings: Sequence[Mapping], unmapped_from_target: bool, penalize_unmapped: bool,
                        base_length_override: Optional[int] = None) -> Tuple[Optional[int], MappingStats]:
    """src/util/mapping.rs:22-57.  Returns (index of best mapping or None, its stats)."""
    o = base_length_override if base_length_override is not None else 1
    best_stats = MappingStats(o, o, 0)
    best_idx = None
    for idx, m in enumerate(mappings):
        if unmapped_from_target:
            bl = base_length_override if base_length_override is not None else m.target_len
            um = bl - (m.target_end - m.target_start)
        else:
            bl = base_length_override if base_length_override is not None else m.query_len
            um = bl - (m.query_end - m.query_start)
        stats = MappingStats(bl, m.nm, um)
        if stats.custom_score(penalize_unmapped) < best_stats.custom_score(penalize_unmapped):
            best_stats, best_idx = stats, idx
    return best_idx, best_stats


# ==========================================================================================
# HLA processed matches -- src/hla/processed_match.rs
# ==========================================================================================
def process_mm_cigar(cigar: Sequence[Tuple[int, int]], target_offset: int, target_len: int, clip_start: int,
                     clip_end: int) -> List[int]:
    """src/hla/processed_match.rs:210-263: number of edits before each target position."""
    zero_padding = max(target_offset - clip_start, 0)
    nm_padding = target_offset - zero_padding
    ret = [0] * (zero_padding + 1)
    cur = 0
    for _ in range(nm_padding):
        cur += 1
        ret.append(cur)
    for length, op in cigar:
        if op == 1:  # I
            cur += length
        elif op in (2, 8):  # D | X
            for _ in range(length):
                cur += 1
                ret.append(cur)
        elif op == 7:  # =
            ret.extend([cur] * length)
        else:
            raise ValueError(f"Unexpected cigar type: {op}")
    missing = target_len + 1 - len(ret)
    ext = min(clip_end, missing)
    for _ in range(ext):
        cur += 1
        ret.append(cur)
    assert len(ret) <= target_len + 1
    ret.extend([cur] * (missing - ext))
    return ret


@dataclass
class HlaProcessedMatch:
    """src/hla/processed_match.rs:10-100."""
    haplotype: str = ""
    full_mapping_stats: List[Optional[MappingStats]] = field(default_factory=list)
    processed_cigars: List[Optional[List[int]]] = field(default_factory=list)
    processed_ranges: List[Tuple[int, int]] = field(default_factory=list)

    @staticmethod
    def worst_match(num_sequences: int) -> "HlaProcessedMatch":
        return HlaProcessedMatch("", [None] * num_sequences, [None] * num_sequences, [(0, 0)] * num_sequences)

    def add_mapping(self, m: Optional[Mapping]):
        """:53-100"""
        if m is None:
            self.full_mapping_stats.append(None)
            self.processed_cigars.append(None)
            self.processed_ranges.append((0, 0))
            return
        if not m.forward:
            raise ValueError("Reverse strand mappings are not supported by HlaProcessedMatch")
        clip_start = m.query_start
        clip_end = m.query_len - m.query_end
        pc = process_mm_cigar(m.cigar, m.target_start, m.target_len, clip_start, clip_end)
        pc_start = max(m.target_start - clip_start, 0)
        pc_end = m.target_end + min(clip_end, m.target_len - m.target_end)
        unmapped = m.query_len - (m.query_end - m.query_start)
        assert len(pc) == m.target_len + 1
        self.full_mapping_stats.append(MappingStats(m.query_len, m.nm, unmapped))
        self.processed_cigars.append(pc)
        self.processed_ranges.append((pc_start, pc_end))

    def is_better_match(self, rhs: "HlaProcessedMatch") -> bool:
        """:103-184"""
        assert len(self.processed_cigars) == len(rhs.processed_cigars)
        for i, (l, r) in enumerate(zip(self.processed_cigars, rhs.processed_cigars)):
            if l is not None and r is not None:
                os_ = max(self.processed_ranges[i][0], rhs.processed_ranges[i][0])
                oe = min(self.processed_ranges[i][1], rhs.processed_ranges[i][1])
                if os_ < oe:
                    lnm, rnm = l[oe] - l[os_], r[oe] - r[os_]
                else:
                    lnm, rnm = 0, 0
                if lnm < rnm:
                    return True
                if lnm > rnm:
                    return False
            elif l is None and r is None:
                continue
            elif l is not None:
                return True
            else:
                return False
        assert len(self.full_mapping_stats) == 2 and len(rhs.full_mapping_stats) == 2
        ls = HlaMappingStats(self.full_mapping_stats[0], self.full_mapping_stats[1]).mapping_score()
        rs = HlaMappingStats(rhs.full_mapping_stats[0], rhs.full_mapping_stats[1]).mapping_score()
        return ls < rs


def realign_select(read_len: int, hits: Sequence[Mapping]) -> Tuple[Optional[int], MappingStats]:
    """HlaRealigner::realign_record acceptance loop, src/hla/realigner.rs:124-146."""
    best = MappingStats(read_len, read_len, 0)
    best_idx = None
    for idx, m in enumerate(hits):
        um = m.target_len - (m.target_end - m.target_start)
        st = MappingStats(m.target_len, m.nm, um)
        if st.mapping_score() <= 0.5 and st.custom_score(False) <= 0.03 and st.custom_score(False) < best.custom_score(False):
            best, best_idx = st, idx
    return best_idx, best


# ==========================================================================================
# het / hom decision -- src/hla/caller.rs:1225-1247, :889-901
# ==========================================================================================
def binomial_cdf_exact(n: int, p: float, k: int) -> float:
    """P[X <= k], X ~ Binomial(n, p) from the exact rational pmf (independent check of binomial_cdf below)."""
    if k >= n:
        return 1.0
    pf = Fraction(p)
    q = 1 - pf
    acc = Fraction(0)
    for x in range(0, k + 1):
        acc += math.comb(n, x) * pf ** x * q ** (n - x)
    return float(acc)


def beta_reg(a: float, b: float, x: float) -> float:
    """statrs 0.16 function::beta::beta_reg (Math.NET's BetaRegularized: modified Lentz continued fraction, at most 140
    iterations, symmetry transform above (a + 1) / (a + b + 2)).  Pinned by the reference's own documentation: the example
    hla_debug.json of docs/debug_outputs.md:128-135 (counts 27 / 10 -> cdf 0.019406414321609413) is reproduced digit for digit
    (tests/test_host_logic_cpu.py), which the exact sum (...60988) is not."""
    bt = 0.0 if x == 0.0 or x == 1.0 else math.exp(ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b) + a * math.log(x) + b * math.log(1.0 - x))
    symm = x >= (a + 1.0) / (a + b + 2.0)
    eps = 0.00000000000000011102230246251565
    fpmin = 2.2250738585072014e-308 / eps
    if symm:
        a, b, x = b, a, 1.0 - x
    qab, qap, qam = a + b, a + 1.0, a - 1.0
    c = 1.0
    d = 1.0 - qab * x / qap
    if abs(d) < fpmin:
        d = fpmin
    d = 1.0 / d
    h = d
    for mi in range(1, 141):
        m = float(mi)
        m2 = m * 2.0
        aa = m * (b - m) * x / ((qam + m2) * (a + m2))
        d = 1.0 + aa * d
        if abs(d) < fpmin:
            d = fpmin
        c = 1.0 + aa / c
        if abs(c) < fpmin:
            c = fpmin
        d = 1.0 / d
        h = h * d * c
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2))
        d = 1.0 + aa * d
        if abs(d) < fpmin:
            d = fpmin
        c = 1.0 + aa / c
        if abs(c) < fpmin:
            c = fpmin
        d = 1.0 / d
        dl = d * c
        h *= dl
        if abs(dl - 1.0) <= eps:
            break
    return 1.0 - bt * h / a if symm else bt * h / a


def binomial_cdf(n: int, p: float, k: int) -> float:
    """statrs 0.16 Binomial::cdf: 1.0 for k >= n, else beta_reg(n - k, k + 1, 1 - p)."""
    if k >= n:
        return 1.0
    return beta_reg(float(n) - float(k), float(k) + 1.0, 1.0 - p)


def is_passing_dual(counts1: int, counts2: int, min_consensus_fraction: float = 0.10, min_cdf: float = 0.001,
                    expected_maf: float = 0.45) -> bool:
    """src/hla/caller.rs:1225-1247 (CLI defaults from src/cli/diplotype.rs)."""
    total = counts1 + counts2
    minor = min(counts1, counts2)
    maf = float(minor) / float(total)
    cdf = binomial_cdf(total, expected_maf, minor)
    return maf >= min_consensus_fraction and cdf >= min_cdf


def choose_diplotype(best_id1: str, best_id2: str, counts1: int, counts2: int, **kw) -> Tuple[str, str]:
    """src/hla/caller.rs:889-901."""
    if is_passing_dual(counts1, counts2, **kw):
        return (best_id1, best_id2)
    return (best_id1, best_id1) if counts1 > counts2 else (best_id2, best_id2)


# ==========================================================================================
# statistics -- src/util/stats.rs (+ statrs 0.16 ln_factorial / ln_gamma, restated)
# ==========================================================================================
_FCACHE = [1.0]
for _i in range(1, 171):
    _FCACHE.append(_FCACHE[-1] * float(_i))

_GAMMA_R = 10.900511
_GAMMA_DK = [
    2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
    -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
    -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9,
]
_LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978
_LN_PI = 1.1447298858494001741434273513530587116472948129153


def ln_gamma(x: float) -> float:
    """statrs::function::gamma::ln_gamma (Lanczos, Math.NET coefficients)."""
    if x < 0.5:
        s = _GAMMA_DK[0]
        for i in range(1, len(_GAMMA_DK)):
            s += _GAMMA_DK[i] / (float(i) - x)
        return (_LN_PI - math.log(math.sin(math.pi * x)) - math.log(s) - _LN_2_SQRT_E_OVER_PI
                - (0.5 - x) * math.log((0.5 - x + _GAMMA_R) / math.e))
    s = _GAMMA_DK[0]
    for i in range(1, len(_GAMMA_DK)):
        s += _GAMMA_DK[i] / (x + float(i) - 1.0)
    return math.log(s) + _LN_2_SQRT_E_OVER_PI + (x - 0.5) * math.log((x - 0.5 + _GAMMA_R) / math.e)


def ln_factorial(x: int) -> float:
    """statrs::function::factorial::ln_factorial: table of 171 factorials, ln_gamma beyond."""
    if x < len(_FCACHE):
        return math.log(_FCACHE[x])
    return ln_gamma(float(x) + 1.0)


def multinomial_ln_pmf(probs: Sequence[float], obs: Sequence[int]) -> float:
    """src/util/stats.rs:11-36."""
    assert len(probs) == len(obs)
    total = sum(obs)
    assert total > 0
    coeff = ln_factorial(total)
    for o in obs:
        coeff -= ln_factorial(o)
    acc = 0.0
    for p, x in zip(probs, obs):
        # Rust: xi as f64 * pi.ln(); ln(0) = -inf and 0 * -inf = NaN exactly as in the reference
        lp = math.log(p) if p > 0.0 else float("-inf")
        acc = acc + float(x) * lp
    return coeff + acc


# ==========================================================================================
# CYP2D6 labels -- src/cyp2d6/region_label.rs, src/cyp2d6/definitions.rs
# ==========================================================================================
UNKNOWN, REP6, CYP2D6, LINK, REP7, SPACER, CYP2D7, DELETION, HYBRID, FALSE_ALLELE = (
    "UNKNOWN", "REP6", "CYP2D6", "link_region", "REP7", "spacer", "CYP2D7", "CYP2D6*5", "Hybrid", "FalseAllele")
_CYP2D = {CYP2D6, CYP2D7, DELETION, HYBRID}
_FLOAT_RE = re.compile(r"^[+-]?((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?|inf|infinity|nan)$", re.I)


@dataclass(frozen=True)
class RegionLabel:
    """Cyp2d6RegionLabel, src/cyp2d6/region_label.rs:72-79."""
    region_type: str
    subtype_label: Optional[str] = None

    def is_cyp2d(self) -> bool:  # :39-55
        return self.region_type in _CYP2D

    def is_rep(self) -> bool:  # :58-60
        return self.region_type in (REP6, REP7)

    def is_reported_allele(self) -> bool:  # :63-69
        return self.region_type in (CYP2D6, DELETION, HYBRID)

    def full_allele(self) -> str:  # :139-168
        t, s = self.region_type, self.subtype_label
        if t == CYP2D6:
            return f"{t}*{s}" if s is not None else t
        if t == HYBRID:
            return s if s is not None else t
        if t == FALSE_ALLELE:
            return f"{t}_{s}" if s is not None else t
        return t

    def simplify_allele(self, detailed: bool, cyp_translate: Dict[str, str]) -> str:  # :101-136
        if self.region_type in (CYP2D6, HYBRID):
            s = self.subtype_label
            if s is None:
                return self.full_allele()
            if s in cyp_translate:
                return f"*{cyp_translate[s]}"
            if detailed:
                return f"*{s}"
            if _FLOAT_RE.match(s):
                v = float(s)
                iv = 0 if math.isnan(v) else (2 ** 63 - 1 if v == math.inf else (-2 ** 63 if v == -math.inf else math.floor(v)))
                return f"*{iv}"
            return f"*{s}"
        if self.region_type == DELETION:
            return "*5"
        return self.full_allele()

    def is_allowed_label(self) -> bool:  # :171-173
        return self.region_type not in (UNKNOWN, FALSE_ALLELE)

    def is_allowed_label_pair(self, nxt: "RegionLabel") -> bool:  # :178-222
        t1, t2 = self.region_type, nxt.region_type
        c1, c2 = self.is_cyp2d(), nxt.is_cyp2d()
        double_star5 = t1 == DELETION and t2 == DELETION
        unexpected = (
            t2 == REP6
            or (c1 and t1 != DELETION and t2 != LINK)
            or (t2 == LINK and not c1)
            or (t1 == LINK and not nxt.is_rep())
            or (nxt.is_rep() and t1 != LINK)
            or (self.is_rep() and not (t2 == SPACER or c2))
            or (t2 == SPACER and not (self.is_rep() or t1 == DELETION))
            or (t1 == SPACER and not c2)
            or (t2 == CYP2D7 and t1 != SPACER)
            or t1 == CYP2D7
        )
        return (not double_star5) and (not unexpected)

    def is_normalizing_allele(self, normalize_all: bool) -> bool:  # :254-262
        return self.is_cyp2d() if normalize_all else self.region_type == CYP2D6

    def is_candidate_chain_head(self, normalize_all: bool) -> bool:  # :227-246
        if self.region_type in (REP6, DELETION):
            return True
        if self.region_type in (CYP2D6, HYBRID):
            return self.is_normalizing_allele(normalize_all)
        return False


@dataclass
class Cyp2d6Config:
    """The three members of Cyp2d6Config the chaining code reads (src/cyp2d6/definitions.rs:128-336)."""
    cyp_translate: Dict[str, str]
    inferred_connections: set
    unexpected_singletons: set

    @staticmethod
    def default() -> "Cyp2d6Config":
        tr = {}
        for part in ("intron1", "exon2", "intron2", "exon3", "intron3", "exon4", "intron4", "exon5", "intron5",
                     "exon6", "intron6", "exon7", "intron7", "exon8", "intron8", "exon9"):
            tr[f"CYP2D7::CYP2D6::{part}"] = "13"  # definitions.rs:238-254
        tr.update({"CYP2D6::CYP2D7::intron1": "68", "CYP2D6::CYP2D7::exon2": "68",
                   "CYP2D6::CYP2D7::exon8": "61", "CYP2D6::CYP2D7::intron8": "63"})  # :256-259
        dups = ["1", "2", "3", "4", "6", "9", "10", "17", "28", "29", "35", "41", "43", "45", "146"]  # :268-283
        conns = {(f"*{d}", f"*{d}") for d in dups} | {("*4", "*68"), ("*10", "*36")}  # :285-286
        return Cyp2d6Config(tr, conns, {"*36", "*68"})  # :292-296


CORE, SUB, DEEP = "CoreAlleles", "SubAlleles", "DeepAlleles"


def deep_label(label: "RegionLabel", unique_id: Optional[int], variants: Optional[Sequence[dict]]) -> str:
    """Cyp2d6Region::deep_label, src/cyp2d6/region.rs:60-95: index label + the variants that differ from the assigned allele."""
    parts = [(f"{unique_id}_" if unique_id is not None else "X_") + label.full_allele()]
    for v in variants or []:
        st = v["variant_state"]
        if st in ("Match", "UnknownUnexpected"):
            continue
        parts.append({"Unexpected": "+", "Missing": "-"}.get(st, "?") + v["label"])
    return " ".join(parts)


def convert_chain_to_hap(chain: Sequence[int], labels: Sequence[RegionLabel], detail_level: str,
                         cyp_translate: Dict[str, str], unique_ids: Optional[Sequence[Optional[int]]] = None,
                         variants: Optional[Sequence[Optional[Sequence[dict]]]] = None) -> str:
    """src/cyp2d6/caller.rs:907-957.  DeepAlleles uses Cyp2d6Region::deep_label (src/cyp2d6/region.rs:60-95)."""
    num_non_deletion = 0
    reportable = []
    for c in reversed(chain):
        lab = labels[c]
        keep = lab.is_cyp2d() and lab.region_type != CYP2D7
        if keep and lab.region_type != DELETION:
            num_non_deletion += 1
        if keep:
            reportable.append(c)
    names = []
    for c in reportable:
        lab = labels[c]
        if lab.region_type == DELETION and num_non_deletion > 0:
            continue
        if detail_level == CORE:
            names.append(lab.simplify_allele(False, cyp_translate))
        elif detail_level == SUB:
            names.append(lab.simplify_allele(True, cyp_translate))
        else:
            uid = unique_ids[c] if unique_ids is not None else None
            names.append("(" + deep_label(lab, uid, variants[c] if variants is not None else None) + ")")
    out = []
    i = 0
    while i < len(names):
        j = i
        while j < len(names) and names[j] == names[i]:
            j += 1
        out.append(f"{names[i]}x{j - i}" if j - i > 1 else names[i])
        i = j
    return " + ".join(out)


# ==========================================================================================
# CYP2D6 weights and chains -- src/cyp2d6/chaining.rs, src/cyp2d6/caller.rs:430-537
# ==========================================================================================
SequenceWeights = List[Tuple[int, float]]


def weight_sequence_from_hits(seq_len: int, labels: Sequence[RegionLabel],
                              hits: Sequence[Sequence[Tuple[int, int, int, int]]]) -> SequenceWeights:
    """weight_sequence, src/cyp2d6/chaining.rs:28-103, with the aligner factored out.
    hits[k] = list of (match_score = nm + unmapped_of_segment, clipped_start, clipped_end, con_len)
    for consensus k, in the order the aligner returned them."""
    ret: SequenceWeights = [(seq_len, 0.0)] * len(labels)
    min_ed_frac = 1.0
    for k, (lab, hs) in enumerate(zip(labels, hits)):
        if not lab.is_allowed_label():
            continue
        for (ms, cs, ce, con_len) in hs:
            overlap = 1.0 - float(cs + ce) / float(con_len)
            if ms < ret[k][0] or (ms == ret[k][0] and overlap > ret[k][1]):
                ret[k] = (ms, overlap)
                min_ed_frac = min(min_ed_frac, max(float(ms), 0.1) / float(seq_len))
    return ret if min_ed_frac <= 0.05 else []


def build_chains(read_weights: Dict[str, List[SequenceWeights]], n_haps: int):
    """Chain building, src/cyp2d6/caller.rs:430-537.  read_weights[qname] = weight_sequence result per
    region of that read in read order (empty list = region dropped).  Returns
    (qname_chains, qname_chain_scores, best_allele_mapping_counts)."""
    qname_chains: Dict[str, List[List[int]]] = {}
    qname_scores: Dict[str, List[SequenceWeights]] = {}
    unique_counts = [0] * n_haps
    for qname in sorted(read_weights):
        regions = read_weights[qname]
        if not regions:
            continue
        putative: List[List[int]] = [[]]
        weighted: List[SequenceWeights] = []
        for ws in regions:
            if not ws:
                continue
            min_ed = min(w[0] for w in ws)
            n_min = sum(1 for w in ws if w[0] == min_ed)
            new_pc = []
            for pc in putative:
                for ci, w in enumerate(ws):
                    if w[0] == min_ed:
                        new_pc.append(pc + [ci])
                        if n_min == 1:
                            unique_counts[ci] += 1
            putative = new_pc
            weighted.append(ws)
        if not putative or (len(putative) == 1 and not putative[0]):
            continue
        qname_chains[qname] = [list(c) for c in putative]
        qname_scores[qname] = weighted
    for qname, cs in qname_chains.items():
        kept = [c for c in cs if all(unique_counts[x] > 0 for x in c)]
        if not kept:
            raise RuntimeError(f"chain collapse: {cs} => []")
        qname_chains[qname] = kept
    return qname_chains, qname_scores, unique_counts


@dataclass
class ChainPenalties:
    """src/cyp2d6/chaining.rs:107-139."""
    lasso_penalty: float = 4.0
    ln_ed_penalty: float = 2.0
    unexpected_chain_penalty: float = 10.0
    inferred_edge_penalty: float = 2.0


class NoChainingHead(Exception):
    pass


class NoChainsFound(Exception):
    pass


class NoScorePairs(Exception):
    pass


def is_sub(haystack: Sequence[int], needle: Sequence[int]) -> bool:
    """src/cyp2d6/chaining.rs:782-784 (windows(0) panics in Rust; chains are never empty here)."""
    n = len(needle)
    return any(list(haystack[s:s + n]) == list(needle) for s in range(0, len(haystack) - n + 1))


def containment_score(c1: Sequence[int], c2: Sequence[int], weights: Sequence[SequenceWeights]):
    """src/cyp2d6/chaining.rs:683-731."""
    optimum = sum(min(w for w, _ in sc) for sc in weights)
    worst = sum(max(w for w, _ in sc) for sc in weights)
    best = 2 * worst
    best_chains: List[List[int]] = []
    wl = len(weights)
    for other in (c1, c2):
        if len(other) < wl:
            continue
        for s in range(0, len(other) - wl + 1):
            total = sum(weights[t][other[s + t]][0] for t in range(wl))
            if total < best:
                best = total
                best_chains = []
            if total == best:
                best_chains.append(list(other[s:s + wl]))
    assert best >= optimum
    return best - optimum, best_chains


def chain_best_window(chain: Sequence[int], weights: Sequence[SequenceWeights]) -> int:
    """Per-chain half of containment_score: min over windows, or the 2*worst sentinel when the chain is
    shorter than the read's segment count.  containment(c1,c2) = min(B(c1), B(c2)) - optimum."""
    worst = sum(max(w for w, _ in sc) for sc in weights)
    wl = len(weights)
    best = 2 * worst
    for s in range(0, len(chain) - wl + 1):
        best = min(best, sum(weights[t][chain[s + t]][0] for t in range(wl)))
    return best


def unexpected_count(chain: Sequence[int], labels: Sequence[RegionLabel], cfg: Cyp2d6Config) -> int:
    """src/cyp2d6/chaining.rs:739-775."""
    reduced = [labels[c].simplify_allele(False, cfg.cyp_translate) for c in chain
               if labels[c].is_cyp2d() and labels[c].region_type != CYP2D7]
    errors = 0
    if not reduced or not reduced[0].startswith("*"):
        errors += 1
    if len(reduced) == 1 and reduced[0] in cfg.unexpected_singletons:
        errors += 1
    for a, b in zip(reduced, reduced[1:]):
        if (a, b) not in cfg.inferred_connections:
            errors += 1
    return errors


def count_unexpected_alleles(labels, hap_counts, ignore_limits: bool, normalize_all: bool) -> int:
    """src/cyp2d6/chaining.rs:794-819."""
    total = 0
    for lab, hc in zip(labels, hap_counts):
        if lab.is_allowed_label() and (ignore_limits or lab.is_normalizing_allele(normalize_all) or lab.is_reported_allele()):
            if hc > 0:
                total += hc - 1
    return total


def count_inferred_edges(ci, cj, inferred) -> int:
    """src/cyp2d6/chaining.rs:828-840."""
    n = 0
    for chain in (ci, cj):
        for a, b in zip(chain, chain[1:]):
            if inferred[a][b]:
                n += 1
    return n


def check_chain_inferrences(cfg: Cyp2d6Config, chain, labels, inferred) -> Tuple[bool, bool]:
    """src/cyp2d6/chaining.rs:603-674."""
    last = chain[-1]
    last_is_cyp2d = labels[last].is_cyp2d()
    opt_index = None
    for ci in range(len(chain) - 2, -1, -1):
        if labels[chain[ci]].is_cyp2d():
            opt_index = ci
            break
    start = opt_index if opt_index is not None else 0
    detected = any(inferred[a][b] for a, b in zip(chain[start:], chain[start + 1:]))
    if not detected:
        return True, True
    if not last_is_cyp2d:
        return True, False
    if opt_index is None:
        return True, True
    prev = chain[opt_index]
    h1, h2 = labels[prev], labels[last]
    h1m, h2m = h1.simplify_allele(False, cfg.cyp_translate), h2.simplify_allele(False, cfg.cyp_translate)
    connected = prev != last and (h1m, h2m) in cfg.inferred_connections
    d7_tail = h2.region_type == CYP2D7 and h1.region_type != CYP2D7 and h1.is_cyp2d()
    allowed = connected or d7_tail
    return allowed, allowed


def get_multinomial_score(labels, hap_counts, hap_weights, ignore_limits, normalize_all, ci, cj):
    """src/cyp2d6/chaining.rs:854-903.  Returns None for an invalid pair, else (penalty, alleles, probs, coverage)."""
    alleles, counts, coverage = [], [], []
    for h, lab in enumerate(labels):
        if hap_counts[h] > 0 and (ignore_limits or lab.is_normalizing_allele(normalize_all)):
            alleles.append(h)
            counts.append(hap_counts[h])
            coverage.append(int(rust_round(hap_weights[h])))
    total = sum(counts)
    probs = [float(c) / float(total) for c in counts]
    if not probs or sum(coverage) == 0:
        if (not normalize_all and any(labels[h].region_type == DELETION for h in ci)
                and any(labels[h].region_type == DELETION for h in cj)):
            return 0.0, alleles, probs, coverage
        return None
    return abs(multinomial_ln_pmf(probs, coverage)), alleles, probs, coverage


def rust_round(x: float) -> float:
    """f64::round: half away from zero (Python's round() is half-to-even)."""
    if math.isnan(x) or math.isinf(x):
        return x
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def enumerate_chains(cfg: Cyp2d6Config, obs_chains: Dict[str, List[List[int]]], labels: Sequence[RegionLabel],
                     infer: bool, normalize_all: bool, ignore_limits: bool):
    """Edges, heads and the DFS of src/cyp2d6/chaining.rs:243-396.  Returns (possible_chains, inferred)."""
    n = len(labels)
    down = [[False] * n for _ in range(n)]
    for qname in sorted(obs_chains):
        for chain in obs_chains[qname]:
            for a, b in zip(chain, chain[1:]):
                if labels[a].is_allowed_label() and labels[b].is_allowed_label():
                    if ignore_limits or labels[a].is_allowed_label_pair(labels[b]):
                        down[a][b] = True
    inferred = [[False] * n for _ in range(n)]
    if infer:
        for i, h1 in enumerate(labels):
            down_no_link = not any(down[i])
            for j, h2 in enumerate(labels):
                up_no_link = not any(down[r][j] for r in range(n))
                if ((down_no_link or up_no_link) and not down[i][j] and h1.is_allowed_label() and h2.is_allowed_label()
                        and h1.is_allowed_label_pair(h2)):
                    inferred[i][j] = True
    heads = [i for i, lab in enumerate(labels) if ignore_limits or lab.is_candidate_chain_head(normalize_all)]
    if not heads:
        raise NoChainingHead()
    remaining = [[h] for h in heads]
    possible: List[List[int]] = []
    while remaining:
        cur = remaining.pop()
        ok_inf, ok_cand = check_chain_inferrences(cfg, cur, labels, inferred)
        if not ok_inf:
            continue
        simplified = convert_chain_to_hap(cur, labels, SUB, cfg.cyp_translate)
        if ignore_limits or (simplified != "" and ok_cand):
            possible.append(list(cur))
        last = cur[-1]
        for ext in range(n):
            if down[last][ext] and cur.count(ext) < 3:
                remaining.append(cur + [ext])
        if infer:
            for ext in range(n):
                if inferred[last][ext] and cur.count(ext) < 3:
                    remaining.append(cur + [ext])
    if not possible:
        raise NoChainsFound()
    return possible, inferred


def score_chain_pair(cfg, labels, possible, inferred, obs_chains, chain_scores, i, j, infer, normalize_all, penalties,
                     ignore_limits):
    """One iteration of the pair loop, src/cyp2d6/chaining.rs:414-513, without the heap early-out.
    Returns None for an invalid pair, else a dict of all ChainScore members."""
    n = len(labels)
    ci, cj = possible[i], possible[j]
    hap_counts = [0] * n
    for c in list(ci) + list(cj):
        hap_counts[c] += 1
    lasso = penalties.lasso_penalty * float(count_unexpected_alleles(labels, hap_counts, ignore_limits, normalize_all))
    unmet = 0
    for qname in sorted(obs_chains):
        if not any(is_sub(ci, ch) or is_sub(cj, ch) for ch in obs_chains[qname]):
            unmet += 1
    mismatch = 0 if ignore_limits else unexpected_count(ci, labels, cfg) + unexpected_count(cj, labels, cfg)
    unexpected_pen = float(mismatch) * penalties.unexpected_chain_penalty
    n_inf = count_inferred_edges(ci, cj, inferred) if infer else 0
    inferred_pen = float(n_inf) * penalties.inferred_edge_penalty
    ed = 0
    hap_weights = [0.0] * n
    for qname in sorted(chain_scores):
        cw = chain_scores[qname]
        score, matches = containment_score(ci, cj, cw)
        ed = min(ed + score, 2 ** 64 - 1)
        if matches:
            split = 1.0 / float(len(matches))
            for ch in matches:
                for off, con in enumerate(ch):
                    hap_weights[con] += split * cw[off][con][1]
    ln_ed = float(ed) * penalties.ln_ed_penalty
    mn = get_multinomial_score(labels, hap_counts, hap_weights, ignore_limits, normalize_all, ci, cj)
    if mn is None:
        return None
    mn_pen, alleles, probs, coverage = mn
    primary = ln_ed + mn_pen + lasso + unexpected_pen + inferred_pen  # ChainScore::primary_score, :172-174
    return dict(score=primary, i=i, j=j, edit_distance=ed, unmet_observations=unmet, ln_ed_penalty=ln_ed,
                mn_llh_penalty=mn_pen, allele_expected_penalty=lasso, unexpected_chain_penalty=unexpected_pen,
                inferred_chain_penalty=inferred_pen, reduced_alleles=alleles, reduced_probs=probs,
                reduced_coverage=coverage, partial=lasso + unexpected_pen + inferred_pen)


def find_best_chain_pair(cfg: Cyp2d6Config, obs_chains: Dict[str, List[List[int]]],
                         chain_scores: Dict[str, List[SequenceWeights]], labels: Sequence[RegionLabel], infer: bool,
                         normalize_all: bool, penalties: ChainPenalties, ignore_limits: bool, return_debug: bool = False):
    """src/cyp2d6/chaining.rs:223-592.  Returns (best_chains, dangling allele names[, debug])."""
    if penalties.lasso_penalty < 0.0:
        raise ValueError("Lasso penalty must be >= 0.0")
    possible, inferred = enumerate_chains(cfg, obs_chains, labels, infer, normalize_all, ignore_limits)
    heap: List[dict] = []  # at most 10 entries; "top" = max by (score, i, j)
    max_heap = 10

    def key(e):
        return (e["score"], e["i"], e["j"])

    for i in range(len(possible)):
        for j in range(i, len(possible)):
            if len(heap) >= max_heap:
                # early-out on the cheap terms, :457-464 (needs the partial cost before the expensive part)
                n = len(labels)
                hc = [0] * n
                for c in possible[i] + possible[j]:
                    hc[c] += 1
                partial = (penalties.lasso_penalty * float(count_unexpected_alleles(labels, hc, ignore_limits, normalize_all))
                           + float(0 if ignore_limits else unexpected_count(possible[i], labels, cfg) + unexpected_count(possible[j], labels, cfg)) * penalties.unexpected_chain_penalty
                           + float(count_inferred_edges(possible[i], possible[j], inferred) if infer else 0) * penalties.inferred_edge_penalty)
                if partial >= max(heap, key=key)["score"]:
                    continue
            e = score_chain_pair(cfg, labels, possible, inferred, obs_chains, chain_scores, i, j, infer, normalize_all,
                                 penalties, ignore_limits)
            if e is None:
                continue
            if len(heap) < max_heap or e["score"] < max(heap, key=key)["score"]:
                heap.append(e)
            if len(heap) > max_heap:
                heap.remove(max(heap, key=key))
    if not heap:
        raise NoScorePairs()
    best = min(heap, key=key)
    best_chains = sorted([list(possible[best["i"]]), list(possible[best["j"]])])
    used = set(c for ch in best_chains for c in ch)
    danglers = [f"{i}_{labels[i].full_allele()}" for i in range(len(labels)) if i not in used]
    if return_debug:
        return best_chains, danglers, dict(possible_chains=possible, inferred=inferred, top=sorted(heap, key=key), best=best)
    return best_chains, danglers


# ==========================================================================================
# Result JSON -- serde_json::to_writer_pretty of StarphaseJson (src/data_types/starphase_json.rs:13-21,
# src/util/file_io.rs:37-52): 2-space indent, struct field order, BTreeMap keys sorted, None -> null
# ==========================================================================================
def _json_str(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o == 8:
            out.append("\\b")
        elif o == 12:
            out.append("\\f")
        elif o < 0x20:
            out.append(f"\\u{o:04x}")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def format_f64(v: float) -> str:
    """serde_json's f64: ryu's shortest round-trip digits (Python's repr yields the same digit string) in the layout
    of ryu::pretty::format64 -- decimals for a decimal-point position in (-5, 16], scientific otherwise, ".0" on
    integral values; non-finite -> null."""
    if math.isnan(v) or math.isinf(v):
        return "null"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    from decimal import Decimal

    sign, digs, exp = Decimal(repr(abs(v))).as_tuple()
    digits = "".join(map(str, digs)).lstrip("0")
    stripped = digits.rstrip("0")
    exp += len(digits) - len(stripped)
    digits = stripped
    length, k = len(digits), exp
    kk = length + k
    out = "-" if v < 0 else ""
    if 0 <= k and kk <= 16:
        return out + digits + "0" * k + ".0"
    if 0 < kk <= 16:
        return out + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:
        return out + "0." + "0" * (-kk) + digits
    if length == 1:
        return out + digits + "e" + str(kk - 1)
    return out + digits[0] + "." + digits[1:] + "e" + str(kk - 1)


def serde_pretty(v, indent: int = 0) -> str:
    """serde_json PrettyFormatter: empty containers print as [] / {}, otherwise one item per line."""
    pad = "  " * (indent + 1)
    end = "  " * indent
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return format_f64(v)  # only hla_debug.json holds floats; the main result JSON has none
    if isinstance(v, str):
        return _json_str(v)
    if isinstance(v, (list, tuple)):
        if not v:
            return "[]"
        return "[\n" + ",\n".join(pad + serde_pretty(x, indent + 1) for x in v) + "\n" + end + "]"
    if isinstance(v, dict):
        if not v:
            return "{}"
        return "{\n" + ",\n".join(pad + _json_str(k) + ": " + serde_pretty(x, indent + 1) for k, x in v.items()) + "\n" + end + "}"
    raise TypeError(type(v))


def diplotype_json(hap1: str, hap2: str) -> dict:
    """Diplotype, src/data_types/pgx_diplotype.rs:9-26."""
    return dict(hap1=hap1, hap2=hap2, diplotype=f"{hap1}/{hap2}")


def gene_details_from_mappings(diplotypes: List[dict], mapping_details: List[dict]) -> dict:
    """PgxGeneDetails::new_from_mappings, src/data_types/starphase_json.rs:147-161 (HLA)."""
    return dict(diplotypes=diplotypes, simple_diplotypes=None, inexact_diplotypes=None, variant_details=None,
                mapping_details=mapping_details, multi_mapping_details=None)


def gene_details_from_multi_mappings(diplotypes, simple_diplotypes, inexact_diplotypes, multi_mapping_details) -> dict:
    """PgxGeneDetails::new_from_multi_mappings, src/data_types/starphase_json.rs:169-188 (CYP2D6)."""
    return dict(diplotypes=diplotypes, simple_diplotypes=simple_diplotypes, inexact_diplotypes=inexact_diplotypes,
                variant_details=None, mapping_details=None, multi_mapping_details=multi_mapping_details)


def mapping_details_json(read_qname: str, best_hla_id: str, best_star_allele: str, stats: HlaMappingStats,
                         is_ignored: bool) -> dict:
    """PgxMappingDetails, src/data_types/starphase_json.rs:271-283."""
    return dict(read_qname=read_qname, best_hla_id=best_hla_id, best_star_allele=best_star_allele,
                best_mapping_stats=stats.to_json(), is_ignored=is_ignored)


def starphase_json(pbstarphase_version: str, database_metadata: dict, gene_details: Dict[str, dict]) -> str:
    """StarphaseJson, src/data_types/starphase_json.rs:13-21; metadata field order from
    src/database/pgx_database.rs:359-371.  Returns the exact bytes save_json would write."""
    md = {k: database_metadata[k] for k in ("pbstarphase_version", "cpic_version", "hla_version", "pharmvar_version", "build_time")}
    doc = dict(pbstarphase_version=pbstarphase_version, database_metadata=md,
               gene_details={k: gene_details[k] for k in sorted(gene_details)})
    return serde_pretty(doc)


# ==========================================================================================
# hla_debug.json -- src/hla/debug.rs (HlaDebug, ReadMappingStats, DetailedMappingStats, DualPassingStats)
# ==========================================================================================
_CIGAR_OPS = "MIDNSHP=XB"


def cigar_string(cigar: Sequence[Tuple[int, int]]) -> str:
    """minimap2 crate `cigar_str`: length + op letter per entry."""
    return "".join(f"{ln}{_CIGAR_OPS[op]}" for ln, op in cigar)


def _nt4(c: int) -> int:
    return {65: 0, 97: 0, 67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}.get(c, 4)


def md_string(cigar: Sequence[Tuple[int, int]], target: bytes, t_off: int, query: bytes, q_off: int) -> str:
    """minimap2 2.28 format.c write_MD_core (unvendored dependency, restated from its published source): counts of equal
    bases, mismatches as the target base, deletions as ^bases, no trailing zero."""
    out, run = [], 0
    for ln, op in cigar:
        if op in (0, 7, 8):
            for j in range(ln):
                tq = _nt4(target[t_off + j])
                if _nt4(query[q_off + j]) != tq:
                    out.append(f"{run}{'ACGTN'[tq]}")
                    run = 0
                else:
                    run += 1
            t_off += ln
            q_off += ln
        elif op == 1:
            q_off += ln
        elif op == 2:
            out.append(f"{run}^" + "".join("ACGTN"[_nt4(target[t_off + j])] for j in range(ln)))
            run = 0
            t_off += ln
        elif op == 3:
            t_off += ln
        else:
            raise ValueError(f"Unexpected cigar type: {op}")
    if run > 0:
        out.append(str(run))
    return "".join(out)


def detailed_mapping_stats(m: Mapping, target: bytes, query: bytes) -> dict:
    """DetailedMappingStats::from_mapping, src/hla/debug.rs:161-181 (match_len = mm_reg1_t::mlen = bases under '=')."""
    return dict(query_len=m.query_len, target_len=m.target_len, match_len=sum(ln for ln, op in m.cigar if op == 7), nm=m.nm,
                query_unmapped=m.query_len - (m.query_end - m.query_start), target_unmapped=m.target_len - (m.target_end - m.target_start),
                cigar=cigar_string(m.cigar), md=md_string(m.cigar, target, m.target_start, query, m.query_start))


def read_mapping_stats_json(best_id: Optional[str], best_star: Optional[str], mapping_stats: Dict[str, Tuple[Optional[dict], Optional[dict]]]) -> dict:
    """ReadMappingStats, src/hla/debug.rs:64-73 (BTreeMap: keys sorted bytewise)."""
    return dict(best_match_id=best_id, best_match_star=best_star,
                mapping_stats={k: dict(cdna_mapping=mapping_stats[k][0], dna_mapping=mapping_stats[k][1])
                               for k in sorted(mapping_stats, key=lambda x: x.encode())})


def dual_passing_stats(is_dual: bool, counts1: int = 0, counts2: int = 0, min_consensus_fraction: float = 0.10, min_cdf: float = 0.001,
                       expected_maf: float = 0.45) -> dict:
    """is_passing_dual's DualPassingStats, src/hla/caller.rs:1225-1247 + src/hla/debug.rs:186-226."""
    if not is_dual:
        return dict(is_passing=False, is_dual=False, counts1=None, counts2=None, maf=None, cdf=None)
    total, minor = counts1 + counts2, min(counts1, counts2)
    maf = float(minor) / float(total)
    cdf = binomial_cdf(total, expected_maf, minor)
    return dict(is_passing=maf >= min_consensus_fraction and cdf >= min_cdf, is_dual=True, counts1=counts1, counts2=counts2, maf=maf, cdf=cdf)


def hla_debug_json(read_mapping_stats: Dict[str, Dict[str, dict]], dual_passing: Optional[Dict[str, dict]]) -> str:
    """HlaDebug through save_json, src/hla/debug.rs:6-12."""
    doc = dict(read_mapping_stats={g: {q: read_mapping_stats[g][q] for q in sorted(read_mapping_stats[g], key=lambda x: x.encode())}
                                   for g in sorted(read_mapping_stats, key=lambda x: x.encode())},
               dual_passing_stats=None if dual_passing is None else {g: dual_passing[g] for g in sorted(dual_passing, key=lambda x: x.encode())})
    return serde_pretty(doc)


# ==========================================================================================
# consensus preparation in front of score_read -- src/hla/caller.rs:1337-1368, :1518-1576, src/util/sequence.rs:9-23
# ==========================================================================================
def reverse_complement(seq: bytes) -> bytes:
    comp = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    out = bytearray()
    for c in reversed(seq):
        if c not in comp:
            raise ValueError(f"Unexpected character for reverse-complement: {c}")
        out.append(comp[c])
    return bytes(out)


def aligned_pairs(pos: int, cigar: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """rust-htslib bam::Record::aligned_pairs: [read index, reference position] for every M / = / X column."""
    out, q, r = [], 0, pos
    for ln, op in cigar:
        if op in (0, 7, 8):
            out.extend((q + k, r + k) for k in range(ln))
            q += ln
            r += ln
        elif op in (1, 4):
            q += ln
        elif op in (2, 3):
            r += ln
        elif op in (5, 6):
            pass
        else:
            raise ValueError(f"Unexpected cigar type: {op}")
    return out


def splice_read(sequence: bytes, pos: int, cigar: Sequence[Tuple[int, int]], exons: Sequence[Tuple[int, int]]) -> Tuple[bytes, int]:
    """src/hla/caller.rs:1518-1576."""
    lookup = {}
    for qi, ri in aligned_pairs(pos, cigar):
        lookup[ri] = qi
    offset, segments = 0, []
    for start, end in exons:
        first, last = start, end - 1
        while first not in lookup and first <= last:
            first += 1
        while last not in lookup and first <= last:
            last -= 1
        if not segments:
            offset += first - start
        if first <= last:
            segments.append((lookup[first], lookup[last] + 1))
    return b"".join(sequence[a:b] for a, b in segments), offset


def prepare_score_read_targets(read_sequence: bytes, pos: int, cigar, exons, is_forward_strand: bool, disable_cdna_scoring: bool = False):
    """src/hla/caller.rs:1337-1368: (DNA target, cDNA target)."""
    dna = read_sequence if is_forward_strand else reverse_complement(read_sequence)
    if disable_cdna_scoring:
        return dna, b"N"
    fw, _ = splice_read(read_sequence, pos, cigar, exons)
    if not fw:
        return dna, b"N"
    return dna, (fw if is_forward_strand else reverse_complement(fw))


# ==========================================================================================
# hemizygous test -- src/hla/caller.rs:1583-1653 (statrs 0.16 Normal::ln_pdf, Binomial::ln_pmf restated)
# ==========================================================================================
LN_SQRT_2PI = 0.91893853320467274178032973640561763986139747363778341281715


def normal_ln_pdf(mean: float, std_dev: float, x: float) -> float:
    d = (x - mean) / std_dev
    return (-0.5 * d * d) - LN_SQRT_2PI - math.log(std_dev)


def binomial_ln_pmf(n: int, p: float, x: int) -> float:
    if x > n:
        return -math.inf
    if p == 0.0:
        return 0.0 if x == 0 else -math.inf
    if p == 1.0:
        return 0.0 if x == n else -math.inf
    return (ln_factorial(n) - ln_factorial(x) - ln_factorial(n - x)) + float(x) * math.log(p) + float(n - x) * math.log(1.0 - p)


def is_hemizygous_better(scores1: Sequence[Optional[int]], scores2: Sequence[Optional[int]], is_consensus1: Sequence[bool], is_dual: bool,
                         dual_max_ed_delta: int, normalized_coverage: Optional[float]) -> bool:
    read_count = len(is_consensus1)
    min_ed = 0
    if is_dual:
        c1 = c2 = 0
        for o1, o2 in zip(scores1, scores2):
            assert o1 is not None or o2 is not None
            s1 = o1 if o1 is not None else (o2 or 0) + dual_max_ed_delta
            s2 = o2 if o2 is not None else (o1 or 0) + dual_max_ed_delta
            mn = min(s1, s2)
            c1 += s1 - mn
            c2 += s2 - mn
        min_ed = min(c1, c2)
    haploid_ed_cost = 2.0 * float(min_ed)
    nc_hap = normalized_coverage if normalized_coverage is not None else float(read_count)
    nc_dev = nc_hap * 0.1
    if not nc_dev > 0.0:
        raise ValueError("Bad distribution parameters")
    haploid_cost = haploid_ed_cost + abs(normal_ln_pdf(nc_hap, nc_dev, float(read_count)))
    obs1 = sum(1 for b in is_consensus1 if b)
    balance = 2.0 * abs(binomial_ln_pmf(read_count, 0.5, obs1)) if is_dual else 0.0
    nc_dip = 2.0 * nc_hap
    diploid_cost = balance + abs(normal_ln_pdf(nc_dip, nc_dev, float(read_count)))
    return haploid_cost < diploid_cost


# ==========================================================================================
# CYP2D6 allele-vector typing -- src/cyp2d6/haplotyper.rs:452-601 (the in-tree half of assign_haplotype; the WFA graph
# traversal that produces `traversed nodes` is hiphase / waffle_con code and not restated)
# ==========================================================================================
def alleles_from_traversal(num_variants: int, traversed_nodes: Sequence[int], node_to_alleles: Dict[int, List[Tuple[int, int]]]) -> List[int]:
    """:454-468: 3 = unset, first assignment wins, a conflicting second one turns the site into 2 (ambiguous)."""
    alleles = [3] * num_variants
    for node in traversed_nodes:
        for var_index, assignment in node_to_alleles.get(node, []):
            if alleles[var_index] == 3:
                alleles[var_index] = assignment
            elif alleles[var_index] != assignment:
                alleles[var_index] = 2
    return alleles


def variant_match(alleles: Sequence[int], hap: Sequence[int], is_vi: Sequence[bool]) -> Tuple[int, int]:
    """:475-506: (vi_match, all_match) of one observed vector against one haplotype definition."""
    assert len(alleles) == len(hap)
    vi_match = all_match = 0
    for i, (sv, hv) in enumerate(zip(alleles, hap)):
        assert hv in (0, 1)
        if sv in (0, 1):
            ok = hv == sv
        elif sv == 2:
            ok = True
        elif sv == 3:
            ok = False
        else:
            raise ValueError(f"Unexpected seq_value={sv}")
        if ok:
            all_match += 1
            if is_vi[i]:
                vi_match += 1
    return vi_match, all_match


_VARIANT_STATE = {(0, 0): "Match", (0, 1): "Unexpected", (0, 2): "AmbiguousUnexpected", (0, 3): "UnknownUnexpected",
                  (1, 0): "Missing", (1, 1): "Match", (1, 2): "AmbiguousMissing", (1, 3): "UnknownMissing"}


def assign_haplotype_from_alleles(alleles: Sequence[int], haplotype_lookup: Dict[str, Sequence[int]], labels: Sequence[str],
                                  is_vi: Sequence[bool], force_assignment: bool):
    """:470-601 for star alleles `haplotype_lookup` (key = the subtype label of a Cyp2d6 region label; BTreeMap order =
    bytewise key order).  Returns (best star allele or None for Unknown, region variants or None, best (vi, all) score)."""
    best_set, best_score = {None}, (0, 0)
    for star in sorted(haplotype_lookup, key=lambda x: x.encode()):
        score = variant_match(alleles, haplotype_lookup[star], is_vi)
        if score > best_score:
            best_set, best_score = {star}, score
        elif score == best_score:
            best_set.add(star)
    if len(best_set) == 1:
        best = next(iter(best_set))
    else:
        # full_allele(): "Unknown" for the initial entry, "CYP2D6*{star}" for star alleles (src/cyp2d6/region_label.rs:140-168)
        ordered = sorted(best_set, key=lambda x: (RegionLabel(UNKNOWN) if x is None else RegionLabel(CYP2D6, x)).full_allele().encode())
        best = ordered[0] if force_assignment else None
    if best is None:
        return None, None, best_score
    rv = []
    for i, (sv, hv) in enumerate(zip(alleles, haplotype_lookup[best])):
        state = _VARIANT_STATE[(hv, sv)]
        if state == "Match" and hv == 0:
            continue
        rv.append(dict(label=labels[i], is_vi=bool(is_vi[i]), variant_state=state))
    return best, rv, best_score


# ==========================================================================================
# homopolymer compression -- src/util/homopolymers.rs:18-42
# ==========================================================================================
def hpc(seq: bytes) -> bytes:
    out = bytearray()
    for c in seq:
        if not out or out[-1] != c:
            out.append(c)
    return bytes(out)


def hpc_pos(seq: bytes, position: int) -> int:
    total = offset = 0
    i = 0
    while i < len(seq):
        j = i
        while j < len(seq) and seq[j] == seq[i]:
            j += 1
        total += j - i
        if position < total:
            break
        offset += 1
        i = j
    return offset


def cyp2d6_alleles_json(best: Sequence[Sequence[int]], labels: Sequence[RegionLabel], unique_ids: Sequence[Optional[int]],
                        variants: Sequence[Optional[Sequence[dict]]], cyp_translate: Dict[str, str]) -> str:
    """DeeplotypeDebug through save_json, src/cyp2d6/debug.rs:8-71 (cyp2d6_alleles.json)."""
    assert len(best) == 2

    def hap(chain):
        return dict(deep_form=convert_chain_to_hap(chain, labels, DEEP, cyp_translate, unique_ids, variants),
                    suballele_form=convert_chain_to_hap(chain, labels, SUB, cyp_translate, unique_ids, variants),
                    core_form=convert_chain_to_hap(chain, labels, CORE, cyp_translate, unique_ids, variants))

    alleles = {}
    for lab, uid, var in zip(labels, unique_ids, variants):
        if var is None:
            continue
        key = (f"{uid}_" if uid is not None else "X_") + lab.full_allele()
        assert key not in alleles
        alleles[key] = [dict(label=v["label"], is_vi=bool(v["is_vi"]), variant_state=v["variant_state"]) for v in var]
    return serde_pretty(dict(hap1=hap(best[0]), hap2=hap(best[1]), alleles={k: alleles[k] for k in sorted(alleles, key=lambda x: x.encode())}))
