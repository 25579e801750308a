/*
 * oracle/sp_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA scoring path, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call it.  Nothing under
 * pb_starphase_b200/ links or imports it.
 *
 * PARITY STATUS: "parity unpinned" for the alignment arithmetic itself.
 * The reference (pb-StarPhase v2.0.1) obtains every `nm`/`unmapped` number
 * from minimap2 2.28 through the `minimap2` crate 0.1.23 / `minimap2-sys`
 * 0.1.21+minimap2.2.28 (Cargo.lock:1137-1149), which is not vendored under
 * /root/reference and cannot be built here (no Rust toolchain, no network).
 * This oracle restates the published definition the reference consumes:
 *
 *     D(P, T) = min over placements of P inside T of (edits + unaligned P bases)
 *             = unit-cost infix ("HW" / semi-global) Levenshtein distance,
 *               P consumed completely, both ends of T free,
 *
 * which equals the reference's `nm + unmapped` (src/util/mapping.rs:22-57,
 * src/cyp2d6/chaining.rs:65-80, src/hla/processed_match.rs:78-80) whenever
 * minimap2's alignment is unit-cost optimal and is a lower bound otherwise.
 * It is pinned only in the regimes the reference's own tests pin:
 * exact copy => 0 (src/hla/caller.rs:1709-1773), single-base difference
 * ordering and `N` matching nothing (src/cyp2d6/chaining.rs:1050-1080).
 * Non-ACGT bytes never match anything, mirroring minimap2's NM = blen - mlen +
 * n_ambi accounting.
 *
 * Two independent implementations are kept so that they pin each other:
 *   sp_oracle_infix_dp     textbook O(mn) two-row DP (ground truth)
 *   sp_oracle_infix_myers  Myers 1999 / Hyyro 2003 blocked 64-bit bit-vectors
 *                          (the timed CPU baseline, OpenMP over pairs)
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* A,C,G,T (either case) -> 0..3 ; everything else (N, *, ...) -> 4 = matches nothing */
static inline int sp_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

/*
 * Ground truth.  Rows = pattern (0..m), columns = text (0..n).
 * D[i][0] = i; D[0][j] = 0 (free text prefix) or j when `anchored`
 * (alignment must start at text position 0; used for span recovery on
 * reversed sequences).  Result = min_j D[m][j]; *end_col = smallest such j
 * (number of text bases before the end of the best placement).
 */
int64_t sp_oracle_infix_dp(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n,
                           int anchored, int64_t *end_col) {
    int64_t *col = (int64_t *)malloc((size_t)(m + 1) * sizeof(int64_t));
    for (int64_t i = 0; i <= m; ++i) col[i] = i;
    int64_t best = col[m], best_j = 0;
    for (int64_t j = 1; j <= n; ++j) {
        int ct = sp_code(T[j - 1]);
        int64_t diag = col[0];
        col[0] = anchored ? j : 0;
        for (int64_t i = 1; i <= m; ++i) {
            int cp = sp_code(P[i - 1]);
            int64_t sub = diag + ((cp == ct && cp < 4) ? 0 : 1);
            int64_t del = col[i] + 1;     /* text base unpaired */
            int64_t ins = col[i - 1] + 1; /* pattern base unpaired */
            diag = col[i];
            int64_t v = sub < del ? sub : del;
            col[i] = v < ins ? v : ins;
        }
        if (col[m] < best) { best = col[m]; best_j = j; }
    }
    free(col);
    if (end_col) *end_col = best_j;
    return best;
}

/*
 * Myers/Hyyro blocked bit-parallel version of the same recurrence.
 * Block b holds rows 64b+1 .. 64b+64; bit r of Ph/Mh is the horizontal delta
 * of row 64b+r+1.  The score of row m is carried explicitly.
 */
int64_t sp_oracle_infix_myers(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n,
                              int anchored, int64_t *end_col) {
    if (m == 0) { if (end_col) *end_col = 0; return 0; }
    int64_t nb = (m + 63) / 64;
    uint64_t *peq = (uint64_t *)calloc((size_t)(5 * nb), sizeof(uint64_t));
    uint64_t *pv = (uint64_t *)malloc((size_t)nb * sizeof(uint64_t));
    uint64_t *mv = (uint64_t *)calloc((size_t)nb, sizeof(uint64_t));
    for (int64_t i = 0; i < m; ++i) {
        int c = sp_code(P[i]);
        if (c < 4) peq[c * nb + (i >> 6)] |= 1ull << (i & 63);
    }
    for (int64_t b = 0; b < nb; ++b) pv[b] = ~0ull;
    const int r = (int)((m - 1) & 63);
    int64_t score = m, best = m, best_j = 0;
    for (int64_t j = 1; j <= n; ++j) {
        const uint64_t *eqrow = peq + (int64_t)sp_code(T[j - 1]) * nb; /* row 4 is all zero */
        int hin = anchored ? 1 : 0;
        uint64_t ph = 0, mh = 0;
        for (int64_t b = 0; b < nb; ++b) {
            uint64_t eq = eqrow[b], p = pv[b], q = mv[b];
            uint64_t xv = eq | q;
            if (hin < 0) eq |= 1ull;
            uint64_t xh = (((eq & p) + p) ^ p) | eq;
            ph = q | ~(xh | p);
            mh = p & xh;
            int hout = (int)(ph >> 63) - (int)(mh >> 63);
            if (b == nb - 1) score += (int64_t)((ph >> r) & 1) - (int64_t)((mh >> r) & 1);
            ph <<= 1; mh <<= 1;
            if (hin < 0) mh |= 1ull; else if (hin > 0) ph |= 1ull;
            pv[b] = mh | ~(xv | ph);
            mv[b] = ph & xv;
            hin = hout;
        }
        if (score < best) { best = score; best_j = j; }
    }
    free(peq); free(pv); free(mv);
    if (end_col) *end_col = best_j;
    return best;
}

/*
 * Batched form: D[t * n_patterns + p] for every (target t, pattern p).
 * Sequences are concatenated ASCII with n+1 offsets (the layout of the C-ABI's
 * sp_seqset, include/starphase_gpu.h).  impl: 0 = DP, 1 = Myers.
 * Returns the number of DP cells (sum |P|*|T|) so callers can quote GCUPS.
 */
int64_t sp_oracle_score_batch(const uint8_t *tbases, const int64_t *toffs, int64_t nt,
                              const uint8_t *pbases, const int64_t *poffs, int64_t np,
                              int anchored, int impl, int nthreads,
                              int32_t *D, int32_t *end_col) {
    int64_t cells = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int64_t total = nt * np;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : cells)
    for (int64_t k = 0; k < total; ++k) {
        int64_t t = k / np, p = k % np;
        const uint8_t *T = tbases + toffs[t];
        int64_t n = toffs[t + 1] - toffs[t];
        const uint8_t *P = pbases + poffs[p];
        int64_t m = poffs[p + 1] - poffs[p];
        int64_t e = 0;
        int64_t d = impl ? sp_oracle_infix_myers(P, m, T, n, anchored, &e)
                         : sp_oracle_infix_dp(P, m, T, n, anchored, &e);
        D[k] = (int32_t)d;
        if (end_col) end_col[k] = (int32_t)e;
        cells += m * n;
    }
    return cells;
}

/*
 * Span of the optimal placement (the fields the reference reads from minimap2::Mapping at
 * src/cyp2d6/chaining.rs:66-81 and src/cyp2d6/haplotyper.rs:203-249): end = smallest end column of a best
 * placement; start = the rightmost start among best placements ending there, found by the anchored DP over the
 * reversed pattern and the reversed text prefix T[0..end).  Returns the distance.
 */
int64_t sp_oracle_span(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n, int64_t *start, int64_t *end) {
    int64_t e = 0;
    int64_t d = sp_oracle_infix_dp(P, m, T, n, 0, &e);
    uint8_t *rp = (uint8_t *)malloc((size_t)(m + 1));
    uint8_t *rt = (uint8_t *)malloc((size_t)(e + 1));
    for (int64_t i = 0; i < m; ++i) rp[i] = P[m - 1 - i];
    for (int64_t j = 0; j < e; ++j) rt[j] = T[e - 1 - j];
    int64_t c = 0;
    int64_t d2 = sp_oracle_infix_dp(rp, m, rt, e, 1, &c);
    free(rp); free(rt);
    if (d2 != d) { *start = -1; *end = e; return -1; } /* cannot happen: both are the optimum ending at e */
    *start = e - c;
    *end = e;
    return d;
}

void sp_oracle_span_batch(const uint8_t *tbases, const int64_t *toffs, int64_t nt,
                          const uint8_t *pbases, const int64_t *poffs, int64_t np, int nthreads,
                          int32_t *D, int32_t *S, int32_t *E) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int64_t total = nt * np;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = 0; k < total; ++k) {
        int64_t t = k / np, p = k % np, s = 0, e = 0;
        D[k] = (int32_t)sp_oracle_span(pbases + poffs[p], poffs[p + 1] - poffs[p], tbases + toffs[t],
                                       toffs[t + 1] - toffs[t], &s, &e);
        S[k] = (int32_t)s;
        E[k] = (int32_t)e;
    }
}

/*
 * Traceback alignment (checker of the C-ABI's sp_align_pairs): the fields the reference reads from a
 * minimap2::Mapping in HlaProcessedMatch::add_mapping (src/hla/processed_match.rs:53-100) for pattern P
 * (minimap2's query, the allele) inside text T (its target, the consensus): aligned spans, nm, and the EQX
 * CIGAR that process_mm_cigar (processed_match.rs:210-263) walks (1 = I, 2 = D, 7 = '=', 8 = X; clips are the
 * unaligned pattern ends and are not CIGAR entries).  Full (m+1) x (n+1) int32 matrix, then the canonical walk
 * back from (m, e), e = smallest end column of a best placement: diagonal if D[i-1][j-1] + cost == D[i][j],
 * else up if D[i-1][j] + 1 == D[i][j] ('I'), else left ('D'); column 0 only goes up.  Before the first diagonal or
 * 'D' step, up is preferred whenever it is optimal: a pattern end hanging over the text end (a read that stops inside
 * the allele) then comes out as one trailing 'I' run = clip, the way a local aligner reports it, instead of being
 * interleaved with chance matches.  Leading / trailing 'I' runs become p_start / |P| - p_end.  rec = {dist, nm, p_start, p_end, t_start, t_end, n_cigar}; returns the
 * number of run-length entries written to cigar (forward order), or -1 if cap is too small.
 */
int64_t sp_oracle_align(const uint8_t *P, int64_t m, const uint8_t *T, int64_t n, int32_t *rec, uint32_t *cigar,
                        int64_t cap) {
    const int64_t W = n + 1;
    int32_t *D = (int32_t *)calloc((size_t)((m + 1) * W), sizeof(int32_t));
    for (int64_t j = 0; j <= n; ++j) D[j] = 0;
    for (int64_t i = 1; i <= m; ++i) {
        int cp = sp_code(P[i - 1]);
        int32_t *row = D + i * W, *up = row - W;
        row[0] = (int32_t)i;
        for (int64_t j = 1; j <= n; ++j) {
            int ct = sp_code(T[j - 1]);
            int32_t v = up[j - 1] + ((cp == ct && cp < 4) ? 0 : 1);
            if (up[j] + 1 < v) v = up[j] + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            row[j] = v;
        }
    }
    int64_t e = 0;
    for (int64_t j = 1; j <= n; ++j) if (D[m * W + j] < D[m * W + e]) e = j;
    const int32_t d = D[m * W + e];
    uint32_t *ops = (uint32_t *)malloc((size_t)(m + n + 2) * sizeof(uint32_t)); /* backward run-length list */
    int64_t nops = 0, i = m, j = e;
    uint32_t cur_op = 0, cur_len = 0;
    int started = 0; /* a diagonal or 'D' step has been taken */
    while (i > 0) {
        uint32_t op;
        if (j == 0) { op = 1; --i; }
        else if (!started && D[(i - 1) * W + j] + 1 == D[i * W + j]) { op = 1; --i; }
        else {
            int match = sp_code(P[i - 1]) == sp_code(T[j - 1]) && sp_code(P[i - 1]) < 4;
            int32_t v = D[i * W + j];
            started = 1;
            if (D[(i - 1) * W + j - 1] + (match ? 0 : 1) == v) { op = match ? 7 : 8; --i; --j; }
            else if (D[(i - 1) * W + j] + 1 == v) { op = 1; --i; }
            else { op = 2; --j; }
        }
        if (op == cur_op) ++cur_len;
        else { if (cur_len) ops[nops++] = (cur_len << 4) | cur_op; cur_op = op; cur_len = 1; }
    }
    if (cur_len) ops[nops++] = (cur_len << 4) | cur_op;
    free(D);
    int64_t lo = 0, hi = nops; /* ops[hi-1] is the first entry in forward order */
    int32_t clip_s = 0, clip_e = 0;
    if (hi > lo && (ops[hi - 1] & 15u) == 1) { clip_s = (int32_t)(ops[hi - 1] >> 4); --hi; }
    if (hi > lo && (ops[lo] & 15u) == 1) { clip_e = (int32_t)(ops[lo] >> 4); ++lo; }
    rec[0] = d; rec[1] = d - clip_s - clip_e; rec[2] = clip_s; rec[3] = (int32_t)m - clip_e;
    rec[4] = (int32_t)j; rec[5] = (int32_t)e; rec[6] = (int32_t)(hi - lo);
    if (hi - lo > cap) { free(ops); return -1; }
    for (int64_t k = 0; k < hi - lo; ++k) cigar[k] = ops[hi - 1 - k];
    free(ops);
    return hi - lo;
}

/*
 * Diplotype pair scoring in the north_star ("pre-v0.13") form: for every
 * unordered allele pair i <= j, S[i,j] = sum_r min(D[r,i], D[r,j]); keep the k
 * smallest by the lexicographic key (S, [S2,] i, j) -- the same (score, index1,
 * index2) order the reference uses for chain pairs (src/cyp2d6/chaining.rs:188-197)
 * with allele index = BTreeMap order of hla_id (src/hla/caller.rs:1413).
 * c1 = #{r : D[r,i] <= D[r,j]} feeds the unchanged het/hom test
 * (src/hla/caller.rs:1225-1247).  D is [R][A] row-major int32.
 * Returns the number of records written (min(k, A(A+1)/2)).
 */
typedef struct { uint64_t score, score2; uint32_t i, j, c1, pad; } sp_oracle_pair_rec;

static int rec_less(const sp_oracle_pair_rec *a, const sp_oracle_pair_rec *b) {
    if (a->score != b->score) return a->score < b->score;
    if (a->score2 != b->score2) return a->score2 < b->score2;
    if (a->i != b->i) return a->i < b->i;
    return a->j < b->j;
}

static void topk_push(sp_oracle_pair_rec *heap, int *n, int k, const sp_oracle_pair_rec *r) {
    /* tiny k: keep a sorted array */
    if (*n == k && !rec_less(r, &heap[k - 1])) return;
    int pos = (*n < k) ? (*n)++ : k - 1;
    while (pos > 0 && rec_less(r, &heap[pos - 1])) { heap[pos] = heap[pos - 1]; --pos; }
    heap[pos] = *r;
}

/* D2 may be NULL.  With D2 the key is (S1, S2, i, j): the (cDNA, DNA) lexicographic order of
 * HlaMappingScore (src/hla/mapping.rs:111-117) carried over to pair sums, and a read counts for
 * the first allele when (D[r,i], D2[r,i]) <= (D[r,j], D2[r,j]). */
int sp_oracle_pair_minsum_topk(const int32_t *D, const int32_t *D2, int64_t R, int64_t A, int k, int nthreads,
                               sp_oracle_pair_rec *out) {
    if (k <= 0 || A <= 0) return 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    int maxt = omp_get_max_threads();
#else
    (void)nthreads;
    int maxt = 1;
#endif
    /* column-major copies so the inner loop over reads is contiguous */
    int32_t *Dt = (int32_t *)malloc((size_t)(R * A + 1) * sizeof(int32_t));
    int32_t *Dt2 = D2 ? (int32_t *)malloc((size_t)(R * A + 1) * sizeof(int32_t)) : NULL;
    for (int64_t r = 0; r < R; ++r)
        for (int64_t a = 0; a < A; ++a) {
            Dt[a * R + r] = D[r * A + a];
            if (D2) Dt2[a * R + r] = D2[r * A + a];
        }
    sp_oracle_pair_rec *heaps = (sp_oracle_pair_rec *)malloc((size_t)maxt * k * sizeof(*heaps));
    int *counts = (int *)calloc((size_t)maxt, sizeof(int));
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        sp_oracle_pair_rec *heap = heaps + (size_t)tid * k;
        int *cnt = &counts[tid];
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < A; ++i) {
            for (int64_t j = i; j < A; ++j) {
                uint64_t s = 0, s2 = 0; uint32_t c1 = 0;
                for (int64_t r = 0; r < R; ++r) {
                    int32_t x = Dt[i * R + r], y = Dt[j * R + r];
                    s += (uint64_t)(x < y ? x : y);
                    int le = x <= y;
                    if (D2) {
                        int32_t x2 = Dt2[i * R + r], y2 = Dt2[j * R + r];
                        s2 += (uint64_t)(x2 < y2 ? x2 : y2);
                        if (x == y) le = x2 <= y2;
                    }
                    c1 += (uint32_t)le;
                }
                sp_oracle_pair_rec rec = { s, s2, (uint32_t)i, (uint32_t)j, c1, 0 };
                topk_push(heap, cnt, k, &rec);
            }
        }
    }
    int n = 0;
    for (int t = 0; t < maxt; ++t)
        for (int q = 0; q < counts[t]; ++q) topk_push(out, &n, k, &heaps[(size_t)t * k + q]);
    free(heaps); free(counts); free(Dt); free(Dt2);
    return n;
}

/* Full symmetric matrix form used for the CYP2D6 chain-pair path: S[i*A+j], j>=i (lower part = 0). */
void sp_oracle_pair_minsum_full(const int32_t *D, int64_t R, int64_t A, int nthreads, uint64_t *S) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    memset(S, 0, (size_t)(A * A) * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < A; ++i)
        for (int64_t j = i; j < A; ++j) {
            uint64_t s = 0;
            for (int64_t r = 0; r < R; ++r) {
                int32_t x = D[r * A + i], y = D[r * A + j];
                s += (uint64_t)(x < y ? x : y);
            }
            S[i * A + j] = s;
        }
}

/*
 * Window scan of containment_score (src/cyp2d6/chaining.rs:683-731) for one (read, chain): best window sum, or
 * 2 * worst_weight when the chain is shorter than the read's segment count.  B is [n_reads][n_chains] row-major.
 */
void sp_oracle_chain_windows(int64_t n_chains, const int32_t *chain_off, const int32_t *chain_items, int64_t n_reads,
                             const int32_t *seg_off, const uint32_t *W, int64_t n_haps, int32_t *B) {
    for (int64_t r = 0; r < n_reads; ++r) {
        int64_t s0 = seg_off[r], w = seg_off[r + 1] - s0;
        uint64_t worst = 0;
        for (int64_t t = 0; t < w; ++t) {
            uint32_t mx = 0;
            for (int64_t k = 0; k < n_haps; ++k) if (W[(s0 + t) * n_haps + k] > mx) mx = W[(s0 + t) * n_haps + k];
            worst += mx;
        }
        for (int64_t c = 0; c < n_chains; ++c) {
            int64_t c0 = chain_off[c], len = chain_off[c + 1] - c0;
            uint64_t best = 2 * worst;
            for (int64_t s = 0; s + w <= len; ++s) {
                uint64_t tot = 0;
                for (int64_t t = 0; t < w; ++t) tot += W[(s0 + t) * n_haps + chain_items[c0 + s + t]];
                if (tot < best) best = tot;
            }
            B[r * n_chains + c] = (int32_t)(best > 0x7FFFFFFF ? 0x7FFFFFFF : best);
        }
    }
}

int sp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
