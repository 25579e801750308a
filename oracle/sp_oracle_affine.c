/*
 * oracle/sp_oracle_affine.c -- CPU ORACLE, second opinion.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, tools/affine_divergence.py (a measurement script that feeds DESIGN.md §3) and
 * __graft_entry__.build() may build, load or call this file; nothing under pb_starphase_b200/ does.
 *
 * PARITY STATUS: "parity unpinned" -- this is a MODEL of the reference's arithmetic, not the
 * reference.  pb-StarPhase obtains nm / clips / CIGAR from minimap2 2.28 (minimap2 crate 0.1.23 /
 * minimap2-sys 0.1.21, Cargo.lock:1137-1149; not vendored, not buildable here).  minimap2 seeds
 * (k = 19, w = 19 minimisers), chains, aligns between anchors globally and extends both ends with
 * ksw2's two-piece affine gap DP, stopping each extension at its best-scoring cell.  For one
 * co-linear chain that is the best-scoring LOCAL alignment of the query inside the target under
 *
 *     match +a, mismatch -b, ambiguous (non-ACGT) -1,
 *     gap of k bases  -min(q + k*e, q2 + k*e2)
 *
 * with the map-hifi preset a=1 b=4 q=6 e=2 q2=26 e2=1 (src/util/mapping.rs:8-14; the constants are
 * spelled out at src/hla/caller.rs:1381) and a=5 at the allele-scoring site
 * (src/hla/caller.rs:1370-1379).  This file computes exactly that local alignment (full O(mn)
 * Gotoh DP with five states, no band, no z-drop, no seeding) and reports what the reference reads
 * from a minimap2::Mapping (src/hla/processed_match.rs:53-100, src/util/mapping.rs:22-57):
 * query_start / query_end / target_start / target_end / nm / EQX CIGAR, plus the DP score that
 * minimap2 compares with its -s 200 floor.
 *
 * What it is for: measuring how often the product's unit-cost quantities (K1's infix distance D,
 * K4's canonical path) differ from this cost model's `nm + unmapped`, and how often a downstream
 * call changes (tests/test_affine_divergence_cpu.py, tools/affine_divergence.py).  What it is not:
 * evidence about minimap2's seeding / chaining / z-drop heuristics.
 *
 * Conventions: P = minimap2's query (allele / template / consensus), T = its target.  Rows i =
 * pattern, columns j = text.  CIGAR ops as in BAM: 1 = I (pattern base without a text base),
 * 2 = D (text base without a pattern base), 7 = '=', 8 = X.  Ties: the diagonal wins, then the
 * short-gap deletion, short-gap insertion, long-gap deletion, long-gap insertion (ksw2 takes the
 * first maximum in that order); the end cell is the first maximum in anti-diagonal order (ksw2's
 * extension keeps a maximum only when it is strictly exceeded), the start is the first zero met
 * walking back -- ties prefer the shorter alignment at both ends.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int aff_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

#define AFF_NEG (-0x3fffffff)
/* trace byte: bits 0-2 = source of H (0 diag, 1 E1, 2 F1, 3 E2, 4 F2, 5 start), bit 3..6 = E1 / F1 / E2 / F2 extended */
enum { SRC_DIAG = 0, SRC_E1 = 1, SRC_F1 = 2, SRC_E2 = 3, SRC_F2 = 4, SRC_START = 5 };

/*
 * rec[8] = { score, nm, p_start, p_end, t_start, t_end, n_cigar, dist = nm + (m - (p_end - p_start)) }.
 * costs[6] = { a, b, q, e, q2, e2 }.  Returns the number of CIGAR entries, or -1 when `cap` is too
 * small / memory ran out.  An alignment of score 0 (nothing matches) is the empty alignment:
 * p_start = p_end = t_start = t_end = 0, dist = m.
 */
int64_t sp_oracle_affine_local_banded(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n, const int32_t *costs, int64_t centre,
                                      int64_t band, int32_t *rec, uint32_t *cigar, int64_t cap);

int64_t sp_oracle_affine_local(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n, const int32_t *costs,
                               int32_t *rec, uint32_t *cigar, int64_t cap) {
    return sp_oracle_affine_local_banded(P, m, T, n, costs, 0, -1, rec, cigar, cap);
}

/*
 * The same with the DP restricted to the diagonal band |(j - i) - centre| <= band (1-based cells), the way the product's K9
 * kernel computes it: a cell outside the band counts as H = 0 with no open gap (so alignments may start at the band edge and a
 * gap may open from the cell left of it).  band < 0: no band.
 */
int64_t sp_oracle_affine_local_banded(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n, const int32_t *costs, int64_t centre,
                                      int64_t band, int32_t *rec, uint32_t *cigar, int64_t cap) {
    const int a = costs[0], b = costs[1], q = costs[2], e = costs[3], q2 = costs[4], e2 = costs[5];
    memset(rec, 0, 8 * sizeof(int32_t));
    rec[7] = (int32_t)m;
    if (m == 0 || n == 0) return 0;
    uint8_t *tr = (uint8_t *)malloc((size_t)m * (size_t)n);
    int32_t *H = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));   /* previous row, then current */
    int32_t *F1 = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));  /* vertical gap states per column */
    int32_t *F2 = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
    uint8_t *tc = (uint8_t *)malloc((size_t)n);
    if (!tr || !H || !F1 || !F2 || !tc) { free(tr); free(H); free(F1); free(F2); free(tc); return -1; }
    for (int64_t j = 0; j <= n; ++j) { H[j] = 0; F1[j] = AFF_NEG; F2[j] = AFF_NEG; }
    for (int64_t j = 0; j < n; ++j) tc[j] = (uint8_t)aff_code(T[j]);
    int32_t best = 0;
    int64_t bi = 0, bj = 0;
    for (int64_t i = 1; i <= m; ++i) {
        const int pc = aff_code(P[i - 1]);
        int32_t hdiag = H[0];  /* H[i-1][0] = 0 */
        int32_t hleft = 0;     /* H[i][0] */
        int32_t e1 = AFF_NEG, e2s = AFF_NEG;
        uint8_t *trow = tr + (size_t)(i - 1) * (size_t)n;
        for (int64_t j = 1; j <= n; ++j) {
            uint8_t t = 0;
            if (band >= 0 && ((j - i) - centre > band || (j - i) - centre < -band)) {  /* outside the band: H = 0, no open gap */
                hdiag = H[j];
                H[j] = 0; F1[j] = AFF_NEG; F2[j] = AFF_NEG;
                hleft = 0; e1 = AFF_NEG; e2s = AFF_NEG;
                trow[j - 1] = (uint8_t)SRC_START;
                continue;
            }
            /* horizontal (deletion) states from H[i][j-1] */
            int32_t o = hleft - q - e, x = e1 - e;
            if (x > o) { e1 = x; t |= 1u << 3; } else e1 = o;
            o = hleft - q2 - e2; x = e2s - e2;
            if (x > o) { e2s = x; t |= 1u << 5; } else e2s = o;
            /* vertical (insertion) states from H[i-1][j] */
            const int32_t hup = H[j];
            o = hup - q - e; x = F1[j] - e;
            int32_t f1, f2;
            if (x > o) { f1 = x; t |= 1u << 4; } else f1 = o;
            o = hup - q2 - e2; x = F2[j] - e2;
            if (x > o) { f2 = x; t |= 1u << 6; } else f2 = o;
            F1[j] = f1; F2[j] = f2;
            const int c = tc[j - 1];
            const int32_t s = (pc == 4 || c == 4) ? -1 : (pc == c ? a : -b);
            int32_t h = hdiag + s;
            int src = SRC_DIAG;
            if (e1 > h) { h = e1; src = SRC_E1; }
            if (f1 > h) { h = f1; src = SRC_F1; }
            if (e2s > h) { h = e2s; src = SRC_E2; }
            if (f2 > h) { h = f2; src = SRC_F2; }
            if (h <= 0) { h = 0; src = SRC_START; }
            trow[j - 1] = (uint8_t)(t | src);
            hdiag = hup;
            H[j] = h;
            hleft = h;
            if (h > best || (h == best && h > 0 && i + j < bi + bj)) { best = h; bi = i; bj = j; }
        }
    }
    free(H); free(F1); free(F2); free(tc);
    if (best == 0) { free(tr); return 0; }
    /* walk back */
    uint32_t *rev = (uint32_t *)malloc((size_t)(m + n + 2) * sizeof(uint32_t));
    if (!rev) { free(tr); return -1; }
    int64_t nrev = 0, i = bi, j = bj;
    uint32_t cur_op = 0, cur_len = 0;
    int32_t nm = 0;
    int state = 0; /* 0 = H, 1 E1, 2 F1, 3 E2, 4 F2 */
#define AFF_EMIT(op)                                                        \
    do {                                                                    \
        if ((op) == cur_op) ++cur_len;                                      \
        else { if (cur_len) rev[nrev++] = (cur_len << 4) | cur_op; cur_op = (op); cur_len = 1; } \
    } while (0)
    while (i > 0 && j > 0) {
        if (band >= 0 && ((j - i) - centre > band || (j - i) - centre < -band)) break;  /* left the band: the alignment starts here */
        const uint8_t t = tr[(size_t)(i - 1) * (size_t)n + (size_t)(j - 1)];
        if (state == 0) {
            const int src = t & 7;
            if (src == SRC_START) break;
            if (src == SRC_DIAG) {
                const int pc = aff_code(P[i - 1]), c = aff_code(T[j - 1]);
                const int eq = pc < 4 && pc == c;
                AFF_EMIT(eq ? 7u : 8u);
                nm += !eq;
                --i; --j;
            } else state = src;
        } else if (state == SRC_E1 || state == SRC_E2) {  /* deletion: consumes a text base */
            const int ext = state == SRC_E1 ? (t >> 3) & 1 : (t >> 5) & 1;
            AFF_EMIT(2u); ++nm; --j;
            if (!ext) state = 0;
        } else {  /* insertion: consumes a pattern base */
            const int ext = state == SRC_F1 ? (t >> 4) & 1 : (t >> 6) & 1;
            AFF_EMIT(1u); ++nm; --i;
            if (!ext) state = 0;
        }
    }
#undef AFF_EMIT
    if (cur_len) rev[nrev++] = (cur_len << 4) | cur_op;
    free(tr);
    rec[0] = best; rec[1] = nm; rec[2] = (int32_t)i; rec[3] = (int32_t)bi; rec[4] = (int32_t)j; rec[5] = (int32_t)bj;
    rec[6] = (int32_t)nrev; rec[7] = nm + (int32_t)(m - (bi - i));
    if (nrev > cap) { free(rev); return -1; }
    for (int64_t k = 0; k < nrev; ++k) cigar[k] = rev[nrev - 1 - k];
    free(rev);
    return nrev;
}

/*
 * Batch form, OpenMP over pairs: pair k = (patterns[pair_p[k]], targets[pair_t[k]]).
 * recs[k][8] as above; CIGARs are appended to `cigar` (capacity `cap` entries) in completion order,
 * cig_off[k] = first entry of pair k (-1 if the pool overflowed; rec[6] still holds the count).
 * Returns the number of pool entries used.
 */
int64_t sp_oracle_affine_batch(const uint8_t *tbases, const int64_t *toffs, const uint8_t *pbases, const int64_t *poffs,
                               int64_t n_pairs, const int32_t *pair_t, const int32_t *pair_p, const int32_t *costs,
                               int nthreads, int32_t *recs, int64_t *cig_off, uint32_t *cigar, int64_t cap) {
    int64_t used = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        const int64_t t = pair_t[k], p = pair_p[k];
        const int64_t m = poffs[p + 1] - poffs[p], n = toffs[t + 1] - toffs[t];
        uint32_t *tmp = (uint32_t *)malloc((size_t)(m + n + 2) * sizeof(uint32_t));
        int64_t nc = tmp ? sp_oracle_affine_local(pbases + poffs[p], m, tbases + toffs[t], n, costs, recs + 8 * k, tmp, m + n + 2) : -1;
        int64_t off = -1;
        if (nc >= 0 && cigar) {
#pragma omp critical(sp_affine_pool)
            {
                if (used + nc <= cap) { off = used; used += nc; }
            }
            if (off >= 0) memcpy(cigar + off, tmp, (size_t)nc * sizeof(uint32_t));
        }
        if (cig_off) cig_off[k] = off;
        free(tmp);
    }
    return used;
}
