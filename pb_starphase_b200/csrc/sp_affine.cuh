// sp_affine.cuh -- K9: banded two-piece affine local alignment of selected pairs (sm_100a only).
//
// K1 / K4 compute unit-cost optima; the reference's numbers come from minimap2's two-piece affine alignment (match +a, mismatch
// -b, ambiguous -1, gap of k bases -min(q + k e, q2 + k e2); map-hifi a=1 b=4 q=6 e=2 q2=26 e2=1, a=5 in score_read:
// src/util/mapping.rs:8-14, src/hla/caller.rs:1370-1381).  For one co-linear chain that alignment is the best-scoring LOCAL
// alignment under those costs.  K9 computes exactly that for the pairs the host selects (the candidates K4 already placed), in a
// diagonal band around the unit-cost placement: nm, clips, spans, EQX CIGAR and the DP score that minimap2 compares with -s.
// DESIGN.md 3.1 measures what this closes: `nm + unmapped` of the unit-cost path differs from the cost model's for 5-35 % of HLA
// pairs (gap consolidation, end clipping); inside the band K9 is the cost model.
//
// One warp per pair, rows (pattern bases) in sequence, the 2W + 1 band cells of a row spread over the lanes (CELLS consecutive
// cells per lane).  In band coordinates (k = j - (i + centre - W)) the diagonal neighbour of the previous row keeps its index, the
// vertical one is k + 1 (one shuffle), and the horizontal gap states are max-plus prefix scans over the row:
//     H'[k]  = max(0, H[i-1][k] + s, F1, F2)                                   (everything that does not come from the left)
//     E1[k]  = max_{k' < k} (H'[k'] - q  - e  (k - k'))   E2 likewise with (q2, e2)   (a gap never pays to re-open right after a gap)
//     H[k]   = first maximum of (diagonal, E1, F1, E2, F2) in that order (ksw2's order), 0 = start when the maximum is <= 0
// One trace byte per cell (source of H + the four "gap extended" bits).  The end
// cell is the first maximum in anti-diagonal order (ksw2's extension keeps a maximum only when it is strictly exceeded), the walk
// back stops at the first H = 0: ties prefer the shorter alignment at both ends.  Cells outside the band count as H = 0.
#pragma once
#include "sp_kernels.cuh"
#include "sp_align.cuh"

namespace sp {

constexpr int AFF_NEG = -0x30000000;

struct AffinePairDev {
    long long t_off;      // first text byte of the window
    long long p_off;      // first pattern byte
    long long trace_off;  // bytes, inside the warp's trace slot
    long long cig_off;    // first entry of the backwards CIGAR region
    int32_t n, m;         // window length, pattern length
    int32_t centre;       // j - i of the middle of the band (window coordinates, 1-based cells)
    int32_t cig_len;
    int32_t out;
    int32_t pad_;
};

struct AffineParams {
    const uint8_t *tbases, *pbases;
    const AffinePairDev *pairs;
    uint8_t *trace;            // [n_slots][slot_bytes]
    long long slot_bytes;
    uint32_t *cigar, *dense;
    unsigned long long *dense_used;
    unsigned long long dense_cap;
    AlignRecDev *recs;
    int32_t *scores;
    int n_pairs, W;
    int a, b, q, e, q2, e2;
    int *next_pair;
};

template <int CELLS>
__global__ void __launch_bounds__(128) k9_affine_local(const AffineParams p) {
    __shared__ uint8_t lut[256];
    fill_code_lut(lut);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    uint8_t *tr = p.trace + static_cast<size_t>(slot) * p.slot_bytes;
    const int W = p.W, nb = 2 * W + 1, k0 = lane * CELLS;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(p.next_pair, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= p.n_pairs) break;
        const AffinePairDev pr = p.pairs[q];
        const uint8_t *T = p.tbases + pr.t_off, *P = p.pbases + pr.p_off;
        const int m = pr.m, n = pr.n, c = pr.centre;
        uint8_t *trace = tr + pr.trace_off;
        int Hp[CELLS], F1p[CELLS], F2p[CELLS];  // row i - 1, band coordinates of that row
#pragma unroll
        for (int x = 0; x < CELLS; ++x) { Hp[x] = 0; F1p[x] = AFF_NEG; F2p[x] = AFF_NEG; }
        int best = 0, best_sum = 0x7FFFFFFF, best_i = 0, best_j = 0;
        for (int i = 1; i <= m; ++i) {
            const uint32_t pc = lut[__ldg(P + i - 1)];
            const int jbase = i + c - W;  // column of band cell 0
            // vertical neighbours: cell k + 1 of the previous row
            const int Hn = __shfl_down_sync(0xffffffffu, Hp[0], 1), F1n = __shfl_down_sync(0xffffffffu, F1p[0], 1),
                      F2n = __shfl_down_sync(0xffffffffu, F2p[0], 1);
            int hq[CELLS], f1[CELLS], f2[CELLS];
            uint32_t tb[CELLS];
#pragma unroll
            for (int x = 0; x < CELLS; ++x) {
                const int k = k0 + x, j = jbase + k;
                const bool in = k < nb && j >= 1 && j <= n;
                // previous row, column j: band index k + 1 there; outside the band (or row 0) H = 0 and no open gap
                const bool up_in = k + 1 < nb && i > 1 && j >= 1 && j <= n;
                const int hu = up_in ? (x + 1 < CELLS ? Hp[x + 1] : (lane < 31 ? Hn : 0)) : 0;
                const int f1u = up_in ? (x + 1 < CELLS ? F1p[x + 1] : (lane < 31 ? F1n : AFF_NEG)) : AFF_NEG;
                const int f2u = up_in ? (x + 1 < CELLS ? F2p[x + 1] : (lane < 31 ? F2n : AFF_NEG)) : AFF_NEG;
                uint32_t t = 0;
                int o = hu - p.q - p.e, xx = f1u - p.e;
                if (xx > o) { f1[x] = xx; t |= 1u << 4; } else f1[x] = o;
                o = hu - p.q2 - p.e2; xx = f2u - p.e2;
                if (xx > o) { f2[x] = xx; t |= 1u << 6; } else f2[x] = o;
                // diagonal: previous row, column j - 1: same band index; column 0 / row 0 / outside the band = 0
                const int hd = (i > 1 && j >= 2 && k < nb) ? Hp[x] : 0;
                int s = -1;
                if (in) {
                    const uint32_t tc = lut[__ldg(T + j - 1)];
                    s = (pc == 4u || tc == 4u) ? -1 : (pc == tc ? p.a : -p.b);
                }
                hq[x] = in ? hd + s : AFF_NEG;  // diagonal candidate (kept apart from F for the source order)
                if (!in) { f1[x] = AFF_NEG; f2[x] = AFF_NEG; }
                tb[x] = t;
            }
            // H' = max(0, diag, F1, F2); y = H' + e k for the horizontal scans
            int pm1 = AFF_NEG, pm2 = AFF_NEG;  // running prefix maxima (exclusive) of H' + e k and H' + e2 k
            int e1[CELLS], e2v[CELLS], hprime[CELLS];
#pragma unroll
            for (int x = 0; x < CELLS; ++x) {
                const int k = k0 + x;
                hprime[x] = max(max(hq[x], 0), max(f1[x], f2[x]));
                if (hq[x] == AFF_NEG) hprime[x] = AFF_NEG;  // not a cell
                e1[x] = pm1; e2v[x] = pm2;                   // lane-local exclusive prefix
                if (hprime[x] > AFF_NEG) { pm1 = max(pm1, hprime[x] + p.e * k); pm2 = max(pm2, hprime[x] + p.e2 * k); }
            }
            int c1 = pm1, c2 = pm2;  // lane totals -> inclusive scan across lanes -> exclusive for this lane
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u1 = __shfl_up_sync(0xffffffffu, c1, d), u2 = __shfl_up_sync(0xffffffffu, c2, d);
                if (lane >= d) { c1 = max(c1, u1); c2 = max(c2, u2); }
            }
            int b1 = __shfl_up_sync(0xffffffffu, c1, 1), b2 = __shfl_up_sync(0xffffffffu, c2, 1);
            if (lane == 0) { b1 = AFF_NEG; b2 = AFF_NEG; }
            int Hc[CELLS];
#pragma unroll
            for (int x = 0; x < CELLS; ++x) {
                const int k = k0 + x, j = jbase + k;
                const int m1 = max(e1[x], b1), m2 = max(e2v[x], b2);
                // column 0 (H = 0) and the cell left of band cell 0 (outside the band: counted as H = 0) can open a gap
                const int left_out1 = (k == 0 || j == 1) ? -p.q - p.e : AFF_NEG, left_out2 = (k == 0 || j == 1) ? -p.q2 - p.e2 : AFF_NEG;
                e1[x] = hq[x] > AFF_NEG ? max(m1 > AFF_NEG ? m1 - p.q - p.e * k : AFF_NEG, left_out1) : AFF_NEG;
                e2v[x] = hq[x] > AFF_NEG ? max(m2 > AFF_NEG ? m2 - p.q2 - p.e2 * k : AFF_NEG, left_out2) : AFF_NEG;
                int h = hq[x], src = 0;
                if (e1[x] > h) { h = e1[x]; src = 1; }
                if (f1[x] > h) { h = f1[x]; src = 2; }
                if (e2v[x] > h) { h = e2v[x]; src = 3; }
                if (f2[x] > h) { h = f2[x]; src = 4; }
                if (h <= 0) { h = 0; src = 5; }
                Hc[x] = hq[x] > AFF_NEG ? h : 0;
                tb[x] |= static_cast<uint32_t>(src);
            }
            // "gap extended" bits of the horizontal states: E[k] came from E[k-1] - e rather than from H[k-1] - q - e
            const int Hl = __shfl_up_sync(0xffffffffu, Hc[CELLS - 1], 1), E1l = __shfl_up_sync(0xffffffffu, e1[CELLS - 1], 1),
                      E2l = __shfl_up_sync(0xffffffffu, e2v[CELLS - 1], 1);
#pragma unroll
            for (int x = 0; x < CELLS; ++x) {
                const int k = k0 + x, j = jbase + k;
                const bool in = k < nb && j >= 1 && j <= n;
                const bool left_in = k >= 1 && j >= 2;
                const int hl = left_in ? (x ? Hc[x - 1] : Hl) : 0;
                const int e1l = left_in ? (x ? e1[x - 1] : E1l) : AFF_NEG, e2l = left_in ? (x ? e2v[x - 1] : E2l) : AFF_NEG;
                if (in) {
                    if (e1l - p.e > hl - p.q - p.e) tb[x] |= 1u << 3;
                    if (e2l - p.e2 > hl - p.q2 - p.e2) tb[x] |= 1u << 5;
                    trace[static_cast<size_t>(i - 1) * nb + k] = static_cast<uint8_t>(tb[x]);
                    const int sum = i + j;
                    if (Hc[x] > best || (Hc[x] == best && Hc[x] > 0 && (sum < best_sum || (sum == best_sum && i < best_i)))) {
                        best = Hc[x]; best_sum = sum; best_i = i; best_j = j;
                    }
                }
                Hp[x] = Hc[x]; F1p[x] = f1[x]; F2p[x] = f2[x];
            }
        }
        // warp arg-max: (score desc, i + j asc, i asc)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, d), os = __shfl_xor_sync(0xffffffffu, best_sum, d);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, d), oj = __shfl_xor_sync(0xffffffffu, best_j, d);
            if (ob > best || (ob == best && ob > 0 && (os < best_sum || (os == best_sum && oi < best_i)))) { best = ob; best_sum = os; best_i = oi; best_j = oj; }
        }
        __threadfence_block();
        __syncwarp();
        // walk back (lane 0)
        AlignRecDev rec = {m, 0, 0, 0, 0, 0, 0, 0, 0};
        int ncig = 0;
        long long pos = pr.cig_off + pr.cig_len;
        if (lane == 0 && best > 0) {
            int i = best_i, j = best_j, state = 0, nm = 0;
            uint32_t cur_op = 0, cur_len = 0;
            auto emit = [&](uint32_t op) {
                if (op == cur_op) { ++cur_len; return; }
                if (cur_len) p.cigar[--pos] = (cur_len << 4) | cur_op;
                cur_op = op; cur_len = 1;
            };
            while (i > 0 && j > 0) {
                const int k = j - (i + c - W);
                if (k < 0 || k >= nb) break;  // left the band: the alignment starts here
                const uint32_t t = __ldcg(trace + static_cast<size_t>(i - 1) * nb + k);
                if (state == 0) {
                    const uint32_t src = t & 7u;
                    if (src == 5u) break;
                    if (src == 0u) {
                        const uint32_t pc = lut[P[i - 1]], tc = lut[T[j - 1]];
                        const bool eq = pc < 4u && pc == tc;
                        emit(eq ? CIG_EQ : CIG_X);
                        nm += !eq;
                        --i; --j;
                    } else {
                        state = static_cast<int>(src);
                    }
                } else if (state == 1 || state == 3) {  // deletion: a text base without a pattern base
                    const bool ext = state == 1 ? (t >> 3) & 1u : (t >> 5) & 1u;
                    emit(CIG_D); ++nm; --j;
                    if (!ext) state = 0;
                } else {  // insertion
                    const bool ext = state == 2 ? (t >> 4) & 1u : (t >> 6) & 1u;
                    emit(CIG_I); ++nm; --i;
                    if (!ext) state = 0;
                }
            }
            if (cur_len) p.cigar[--pos] = (cur_len << 4) | cur_op;
            rec.nm = nm; rec.p_start = i; rec.p_end = best_i; rec.t_start = j; rec.t_end = best_j;
            rec.dist = nm + (m - (best_i - i));
            ncig = static_cast<int>(pr.cig_off + pr.cig_len - pos);
        }
        __threadfence_block();
        ncig = __shfl_sync(0xffffffffu, ncig, 0);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        unsigned long long at = 0;
        if (lane == 0 && ncig > 0) at = atomicAdd(p.dense_used, static_cast<unsigned long long>(ncig));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at + static_cast<unsigned long long>(ncig) <= p.dense_cap)
            for (int x = lane; x < ncig; x += 32) p.dense[at + x] = __ldcg(p.cigar + pos + x);
        if (lane == 0) {
            rec.n_cigar = ncig; rec.cigar_off = static_cast<long long>(at);
            p.recs[pr.out] = rec;
            p.scores[pr.out] = best;
        }
        __syncwarp();
    }
}

}  // namespace sp
