// sp_align.cuh -- K4: traceback alignment of selected (text, pattern) pairs (sm_100a only).
//
// Gives the host everything it reads from a minimap2::Mapping at the HLA call sites
// (src/hla/processed_match.rs:53-100 of the reference: query_start / query_end / target_start /
// target_end / nm / EQX cigar), so HlaProcessedMatch::add_mapping, process_mm_cigar (:210-263) and
// is_better_match (:103-184) run unchanged on GPU output, and MappingStats' (seq_len, nm, unmapped)
// split (src/data_types/mapping.rs:7-22) is available for the result JSON.  Row N2 of SURVEY.md §8f.
//
// One warp per pair, pattern-stationary like K1 (one pattern per bin, lane width ALN_U):
//   pass 1  infix distance d and the smallest end column e of a best placement (K1 with end columns)
//   pass 2  the same recurrence over the text window [e - (m + d), e) -- every optimal placement ending at
//           e lies inside it -- keeping per column the vertical deltas (~Pv, Mv are the state anyway; only
//           ~Pv is needed) and Hyyro's diagonal-zero vector D0 in HBM scratch
//   pass 3  the warp walks back from (m, e): diagonal when it explains the cell ('=' or 'X'), else up ('I',
//           a pattern base without a text base), else left ('D'); before the first diagonal / 'D' step up wins ties,
//           so a pattern end hanging over the text end is one trailing 'I' run (a clip), not 'I's between chance matches.  Taking the diagonal first while walking
//           backwards left-aligns gaps, the convention of minimap2's ksw2.  Lane k inspects the k-th cell down
//           the diagonal, so a run of up to 32 diagonal steps costs one round of (L2-latency-bound) loads.
//           Leading / trailing 'I' runs are the clipped pattern ends (query_start, query_len - query_end).
// The CIGAR is run-length encoded BAM style, (len << 4) | op with op 1 = I, 2 = D, 7 = '=', 8 = X.
#pragma once
#include "sp_kernels.cuh"

namespace sp {

constexpr int ALN_U = 16;
constexpr uint32_t CIG_I = 1, CIG_D = 2, CIG_EQ = 7, CIG_X = 8;

struct AlignRecDev {  // layout of sp_align_rec (include/starphase_gpu.h)
    int32_t dist, nm, p_start, p_end, t_start, t_end, n_cigar, pad_;
    long long cigar_off;
};

struct AlignParams {
    const uint32_t *blobs;     // [n distinct patterns] bins: forward rows, infix (wildcard) pad rows
    const uint8_t *tbases;     // ASCII texts
    const long long *toffs;
    const int32_t *pair_t;     // text index of each pair
    const int32_t *win_begin;  // optional [n_pairs]: the pair is aligned inside T[win_begin, win_end) only (coordinates in the
    const int32_t *win_end;    //   records are relative to win_begin); nullptr = the whole text
    const int32_t *pair_p;     // blob index of each pair
    const long long *cig_off;  // [n_pairs + 1] region of each pair in `cigar`
    uint32_t *cigar;
    uint32_t *scratch;         // [n_slots][slot_words]
    long long slot_words;
    AlignRecDev *recs;
    int n_pairs;
    uint32_t one, m1, seed_a, seed_b;
};

// the K1 recurrence over `ncols` text columns T[0 .. ncols) for this warp's bin; STORE keeps (D0, ~Pv) per column
template <bool STORE>
// band_lo / band_hi (STORE only): a column j keeps the words of this lane only if some row i of the lane has
// band_lo <= i - j <= band_hi (cell coordinates, 1-based); the walk back never leaves that diagonal band
__device__ __forceinline__ void k4_forward(const uint32_t *blob, const uint8_t *T, int ncols, const AlignParams &p,
                                           bool first, bool owns, int m, int &best, int &best_col, uint32_t *scr,
                                           int Wp, int wf4, int row_lo = 0, int band_lo = -0x40000000, int band_hi = 0x40000000) {
    constexpr int U = ALN_U;
    const int lane = threadIdx.x & 31;
    const int row_hi = row_lo + 32 * U - 1;
    uint32_t npv[U], mv[U], d0[U];
    load_row<U>(blob + 5 * (32 * U), lane, npv);
#pragma unroll
    for (int u = 0; u < U; ++u) { npv[u] = ~npv[u]; mv[u] = 0; d0[u] = 0; }
    int score = m, col = 0;
    best = m; best_col = 0;
    uint32_t carry_out = 0;
    const int nch = (ncols + K1_CHUNK - 1) / K1_CHUNK;
    const int nsteps = nch + 31;
    for (int s = 0; s < nsteps; ++s) {
        uint32_t cin = __shfl_up_sync(0xffffffffu, carry_out, 1);
        if (first) cin = 0u;  // infix: the row above the pattern is free
        const int idx = s - lane;
        if (static_cast<unsigned>(idx) < static_cast<unsigned>(nch)) {
            uint32_t X = cin << 24, Y = cin << 16;
            uint32_t cph = 0, cmh = 0;
#pragma unroll 1
            for (int c = 0; c < K1_CHUNK; ++c) {
                const int j = idx * K1_CHUNK + c;
                const uint32_t code = j < ncols ? base_code(T[j]) : 4u;
                column_step<U, true, STORE>(blob, lane, code, p.one, p.m1, p.seed_a, p.seed_b, npv, mv, X, Y, cph, cmh, score, best, col,
                                            best_col, d0);
                if (STORE && owns && j < ncols && row_hi >= j + 1 + band_lo && row_lo <= j + 1 + band_hi) {
                    uint32_t *dst = scr + static_cast<size_t>(j) * 2 * Wp + (lane * U - wf4);
#pragma unroll
                    for (int q = 0; q < U / 4; ++q) {
                        if (lane * U + 4 * q >= wf4) {
                            *reinterpret_cast<uint4 *>(dst + 4 * q) = make_uint4(d0[4 * q], d0[4 * q + 1], d0[4 * q + 2], d0[4 * q + 3]);
                            *reinterpret_cast<uint4 *>(dst + Wp + 4 * q) =
                                make_uint4(npv[4 * q], npv[4 * q + 1], npv[4 * q + 2], npv[4 * q + 3]);
                        }
                    }
                }
            }
            carry_out = cph | (cmh << 8);
        }
    }
}

__global__ void __launch_bounds__(K1_THREADS) k4_align(const AlignParams p) {
    constexpr int U = ALN_U;
    constexpr int BW = blob_words(U);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *blob = reinterpret_cast<uint32_t *>(smem_raw) + warp * BW;
    const int warps_per_cta = blockDim.x >> 5;  // the host spreads the slots over all SMs: 1..K1_WARPS warps per CTA
    const int slot = blockIdx.x * warps_per_cta + warp, n_slots = gridDim.x * warps_per_cta;
    uint32_t *scr = p.scratch + static_cast<size_t>(slot) * p.slot_words;
    int cur_blob = -1;
    for (int q = slot; q < p.n_pairs; q += n_slots) {
        const int pb = p.pair_p[q], t = p.pair_t[q];
        if (pb != cur_blob) {
            __syncwarp();
            const uint32_t *src = p.blobs + static_cast<size_t>(pb) * BW;
            for (int i = lane; i < BW; i += 32) blob[i] = src[i];
            __syncwarp();
            cur_blob = pb;
        }
        const uint32_t pat = blob[PEQ_ROWS * 32 * U + lane];
        const uint32_t info1 = blob[PEQ_ROWS * 32 * U + 32 + lane];
        const bool owns = pat != NO_PATTERN;
        const bool first = (info1 & INFO_FIRST) != 0, last = owns && (info1 & INFO_LAST) != 0;
        const int m = static_cast<int>(info1 & INFO_LEN_MASK);
        AlignRecDev rec = {0, 0, 0, 0, 0, 0, 0, 0, p.cig_off[q + 1]};
        if (blob[PEQ_ROWS * 32 * U] == NO_PATTERN) {  // empty pattern: distance 0, empty placement at column 0
            if (lane == 0) p.recs[q] = rec;
            continue;
        }
        const int m_all = __shfl_sync(0xffffffffu, m, 0);
        const int nl = (m_all + 32 * U - 1) / (32 * U);
        const int pad = nl * 32 * U - m_all;
        const int wf4 = (pad >> 5) & ~3, Wp = nl * U - wf4;
        const uint8_t *T = p.tbases + p.toffs[t];
        int n = static_cast<int>(p.toffs[t + 1] - p.toffs[t]);
        if (p.win_begin) { T += p.win_begin[q]; n = p.win_end[q] - p.win_begin[q]; }
        int best, best_col;
        const int src_lane = __ffs(__ballot_sync(0xffffffffu, last)) - 1;
        int d, e, w0, ncols;
        const int row_lo = lane * 32 * U - pad + 1;  // first pattern row (cell coordinates) of this lane
        if (n <= 2 * m_all) {
            // the whole text fits the pair's scratch slot (the host sizes it for min(n, 2m) columns): one pass that both
            // finds (d, e) and keeps the columns -- the consensus-sized texts of score_read and the placement windows of
            // the template search take this path
            k4_forward<true>(blob, T, n, p, first, owns, m, best, best_col, scr, Wp, wf4);
            d = __shfl_sync(0xffffffffu, best, src_lane); e = __shfl_sync(0xffffffffu, best_col, src_lane);
            w0 = 0; ncols = e;
        } else {
            k4_forward<false>(blob, T, n, p, first, owns, m, best, best_col, nullptr, 0, 0);
            d = __shfl_sync(0xffffffffu, best, src_lane); e = __shfl_sync(0xffffffffu, best_col, src_lane);
            w0 = max(0, e - (m_all + d)); ncols = e - w0;
            // d is known here: the walk back cannot leave the diagonal band |i - j - (m - ncols)| <= d (each edit moves i - j by
            // at most one), so only the lanes touching the band keep their columns.  (Measured: the kernel is bound by the
            // forward recurrence, not by these stores -- a 6 kb template costs the same with one full-store pass as with a
            // compute-only pass plus a banded-store pass -- so short texts stay on the single pass above.)
            const int delta_end = m_all - ncols;
            k4_forward<true>(blob, T + w0, ncols, p, first, owns, m, best, best_col, scr, Wp, wf4, row_lo, delta_end - d - 1,
                             delta_end + d + 1);
        }
        __threadfence_block();
        __syncwarp();
        {
            // pass 3, warp-cooperative: lane k looks at the diagonal cell (i - k, j - k); the leading run of cells the
            // diagonal explains is taken in one round (HiFi-like pairs are long '=' runs), anything else is one step.
            // (i, j, cur_op, cur_len, pos) are computed from ballots only, so they stay uniform across the warp.
            const long long cig_end = p.cig_off[q + 1];
            long long pos = cig_end;
            uint32_t cur_op = 0, cur_len = 0;
            int i = m_all, j = ncols;
            bool started = false;  // a diagonal or 'D' step has been taken: until then 'I' wins ties (trailing clip in one run)
            auto emit = [&](uint32_t op, uint32_t n) {
                if (op == cur_op) { cur_len += n; return; }
                if (cur_len) { --pos; if (lane == 0) p.cigar[pos] = (cur_len << 4) | cur_op; }
                cur_op = op; cur_len = n;
            };
            while (i > 0) {
                if (j == 0) { emit(CIG_I, static_cast<uint32_t>(i)); i = 0; break; }
                const int ci = i - lane, cj = j - lane;
                bool diag_ok = false, is_eq = false, pv = false;
                if (ci >= 1 && cj >= 1) {
                    const int rr = ci - 1 + pad, w = rr >> 5, b = rr & 31;
                    const uint32_t *colp = scr + static_cast<size_t>(cj - 1) * 2 * Wp + (w - wf4);
                    const uint32_t d0w = __ldcg(colp), npvw = __ldcg(colp + Wp);
                    const uint32_t code = base_code(T[w0 + cj - 1]);
                    is_eq = code < 4 && ((blob[code * 32 * U + row_word(U, w / U, w % U)] >> b) & 1u);
                    diag_ok = is_eq || !((d0w >> b) & 1u);  // the diagonal explains the cell ('=' or a substitution)
                    pv = !((npvw >> b) & 1u);                // vertical delta +1: the cell above explains it ('I')
                }
                const uint32_t ok_mask = __ballot_sync(0xffffffffu, diag_ok);
                const uint32_t eq_mask = __ballot_sync(0xffffffffu, is_eq);
                const bool pv0 = __shfl_sync(0xffffffffu, pv ? 1 : 0, 0) != 0;
                if (!started && pv0) { emit(CIG_I, 1); --i; continue; }
                started = true;
                const int n_diag = ok_mask == 0xffffffffu ? 32 : __ffs(~ok_mask) - 1;
                if (n_diag == 0) {
                    if (pv0) { emit(CIG_I, 1); --i; } else { emit(CIG_D, 1); --j; }
                    continue;
                }
                int done = 0;
                while (done < n_diag) {  // runs of '=' / 'X' among the first n_diag lanes, in walk order
                    const bool eq = (eq_mask >> done) & 1u;
                    const uint32_t rest = (eq ? ~eq_mask : eq_mask) >> done;  // first lane of the other kind
                    int run = rest ? __ffs(rest) - 1 : 32;
                    run = min(run, n_diag - done);
                    emit(eq ? CIG_EQ : CIG_X, static_cast<uint32_t>(run));
                    done += run;
                }
                i -= n_diag; j -= n_diag;
            }
            if (cur_len) { --pos; if (lane == 0) p.cigar[pos] = (cur_len << 4) | cur_op; }
            if (lane == 0) {
                int ncig = static_cast<int>(cig_end - pos), clip_s = 0, clip_e = 0;
                if (ncig > 0 && (p.cigar[pos] & 15u) == CIG_I) { clip_s = static_cast<int>(p.cigar[pos] >> 4); ++pos; --ncig; }
                if (ncig > 0 && (p.cigar[cig_end - 1] & 15u) == CIG_I) { clip_e = static_cast<int>(p.cigar[cig_end - 1] >> 4); --ncig; }
                rec.dist = d; rec.nm = d - clip_s - clip_e;
                rec.p_start = clip_s; rec.p_end = m_all - clip_e;
                rec.t_start = w0 + j; rec.t_end = e;
                rec.n_cigar = ncig; rec.cigar_off = pos;
                p.recs[q] = rec;
            }
        }
        __syncwarp();
    }
}

// dense copy of the kept CIGAR entries: one CTA per pair
__global__ void k4_compact_cigar(const AlignRecDev *__restrict__ recs, const uint32_t *__restrict__ cigar,
                                 const long long *__restrict__ out_off, uint32_t *__restrict__ out) {
    const AlignRecDev r = recs[blockIdx.x];
    const long long o = out_off[blockIdx.x];
    for (int i = threadIdx.x; i < r.n_cigar; i += blockDim.x) out[o + i] = cigar[r.cigar_off + i];
}

}  // namespace sp
