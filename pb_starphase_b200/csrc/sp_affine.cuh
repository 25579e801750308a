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
// One trace byte per cell (source of H + the four "gap extended" bits), a row = 32 * CELLS bytes written with one vector store
// per lane.  Pattern and text codes never sit on the critical path: the text codes of the band live in registers and move one cell
// per row (one shuffle), the pattern base and the one new text base of a row are fetched 32 rows ahead.  Rows whose whole band lies
// inside the text take a path without bounds tests.  The end
// cell is the first maximum in anti-diagonal order (ksw2's extension keeps a maximum only when it is strictly exceeded), the walk
// back stops at the first H = 0: ties prefer the shorter alignment at both ends.  Cells outside the band count as H = 0.  The walk
// back is done by the whole warp: lane l looks at the l-th cell down the diagonal (or along the gap), so a run of up to 32 steps
// costs one round of loads.
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
    int32_t band;         // half width W of this pair's band
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
    int n_pairs;
    int a, b, q, e, q2, e2;
    int *next_pair;
};

template <int CELLS> struct K9Trace;
template <> struct K9Trace<2> { using type = uint16_t; };
template <> struct K9Trace<4> { using type = uint32_t; };
template <> struct K9Trace<8> { using type = uint2; };
template <> struct K9Trace<16> { using type = uint4; };

template <int CELLS>
__device__ __forceinline__ void k9_store_trace(uint8_t *row, int lane, const uint32_t (&tb)[CELLS]) {
    if constexpr (CELLS == 2) {
        reinterpret_cast<uint16_t *>(row)[lane] = static_cast<uint16_t>(tb[0] | (tb[1] << 8));
    } else {
        uint32_t w[CELLS / 4];
#pragma unroll
        for (int g = 0; g < CELLS / 4; ++g) w[g] = tb[4 * g] | (tb[4 * g + 1] << 8) | (tb[4 * g + 2] << 16) | (tb[4 * g + 3] << 24);
        if constexpr (CELLS == 4) reinterpret_cast<uint32_t *>(row)[lane] = w[0];
        else if constexpr (CELLS == 8) reinterpret_cast<uint2 *>(row)[lane] = make_uint2(w[0], w[1]);
        else reinterpret_cast<uint4 *>(row)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// State a lane carries from row to row: H / F1 / F2 of its CELLS band cells of the previous row (cells that are not cells of the
// problem hold H = 0 and no open gap, which is exactly what a neighbour outside the band or the matrix counts as), the text codes
// under its cells, and its best end cell so far.
template <int CELLS>
struct K9Lane {
    int Hp[CELLS], F1p[CELLS], F2p[CELLS];
    uint32_t tc[CELLS];
    int best, best_sum, best_i, best_j;
};

// One row.  INTERIOR: every band cell k < nb has a column inside [1, n].
template <int CELLS, bool INTERIOR>
__device__ __forceinline__ void k9_row(const AffineParams &p, K9Lane<CELLS> &L, int i, uint32_t pc, int jbase, int n, int nb, int lane, uint8_t *trace_row) {
    const int k0 = lane * CELLS;
    const int Hn = __shfl_down_sync(0xffffffffu, L.Hp[0], 1), F1n = __shfl_down_sync(0xffffffffu, L.F1p[0], 1),
              F2n = __shfl_down_sync(0xffffffffu, L.F2p[0], 1);  // lane 31 reads its own cell 0; its last cell is never a cell (nb is odd)
    const int oe1 = p.q + p.e, oe2 = p.q2 + p.e2;
    const int s_match = pc == 4u ? -1 : p.a, s_mis = pc == 4u ? -1 : -p.b;
    int hq[CELLS], f1[CELLS], f2[CELLS], e1[CELLS], e2v[CELLS];
    uint32_t tb[CELLS];
    bool in[CELLS];
    // the cell left of band cell 0 lies outside the band (H = 0) and can open a gap into the row: a virtual cell k = -1
    int pm1 = lane == 0 ? -p.e : AFF_NEG, pm2 = lane == 0 ? -p.e2 : AFF_NEG;
#pragma unroll
    for (int x = 0; x < CELLS; ++x) {
        const int k = k0 + x, j = jbase + k;
        in[x] = INTERIOR ? k < nb : (k < nb && j >= 1 && j <= n);
        const int hu = x + 1 < CELLS ? L.Hp[x + 1] : Hn, f1u = x + 1 < CELLS ? L.F1p[x + 1] : F1n, f2u = x + 1 < CELLS ? L.F2p[x + 1] : F2n;
        uint32_t t = 0;
        int o = hu - oe1, xx = f1u - p.e;
        if (xx > o) { o = xx; t |= 1u << 4; }
        f1[x] = in[x] ? o : AFF_NEG;
        o = hu - oe2; xx = f2u - p.e2;
        if (xx > o) { o = xx; t |= 1u << 6; }
        f2[x] = in[x] ? o : AFF_NEG;
        const uint32_t tcx = L.tc[x];
        const int s = tcx == pc ? s_match : (tcx == 4u ? -1 : s_mis);
        hq[x] = in[x] ? L.Hp[x] + s : AFF_NEG;  // diagonal candidate (kept apart from F for the source order)
        tb[x] = t;
        // H' = everything that does not come from the left; y = H' + e k feeds the horizontal scans (exclusive prefix maxima)
        const int hprime = max(max(hq[x], 0), max(f1[x], f2[x]));
        e1[x] = pm1; e2v[x] = pm2;
        int y1 = hprime + p.e * k, y2 = hprime + p.e2 * k;
        if (!in[x]) {
            // column 0 (H = 0) opens gaps too; any other non-cell is silent
            const bool col0 = !INTERIOR && j == 0 && k < nb;
            y1 = col0 ? p.e * k : AFF_NEG; y2 = col0 ? p.e2 * k : AFF_NEG;
        }
        pm1 = max(pm1, y1); pm2 = max(pm2, y2);
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
    int rowbest = 0, rowx = 0;
#pragma unroll
    for (int x = 0; x < CELLS; ++x) {
        const int k = k0 + x;
        // values far below AFF_NEG / 2 stand for "no such state": they never win against a real score
        e1[x] = in[x] ? max(e1[x], b1) - p.q - p.e * k : AFF_NEG;
        e2v[x] = in[x] ? max(e2v[x], b2) - p.q2 - p.e2 * k : AFF_NEG;
        int h = hq[x];
        uint32_t src = 0;
        if (e1[x] > h) { h = e1[x]; src = 1; }
        if (f1[x] > h) { h = f1[x]; src = 2; }
        if (e2v[x] > h) { h = e2v[x]; src = 3; }
        if (f2[x] > h) { h = f2[x]; src = 4; }
        if (h <= 0) { h = 0; src = 5; }
        Hc[x] = in[x] ? h : 0;
        tb[x] |= src;
        if (Hc[x] > rowbest) { rowbest = Hc[x]; rowx = x; }  // first maximum of the lane's cells = smallest column
    }
    // "gap extended" bits of the horizontal states: E[k] came from E[k-1] - e rather than from H[k-1] - q - e
    const int Hl = __shfl_up_sync(0xffffffffu, Hc[CELLS - 1], 1), E1l = __shfl_up_sync(0xffffffffu, e1[CELLS - 1], 1),
              E2l = __shfl_up_sync(0xffffffffu, e2v[CELLS - 1], 1);
#pragma unroll
    for (int x = 0; x < CELLS; ++x) {
        // left neighbour outside the band / the matrix: H = 0, no open gap -- what non-cells hold; lane 0 has no left lane
        const int hl = x ? Hc[x - 1] : (lane ? Hl : 0);
        const int e1l = x ? e1[x - 1] : (lane ? E1l : AFF_NEG), e2l = x ? e2v[x - 1] : (lane ? E2l : AFF_NEG);
        if (e1l > hl - p.q) tb[x] |= 1u << 3;
        if (e2l > hl - p.q2) tb[x] |= 1u << 5;
        L.Hp[x] = Hc[x]; L.F1p[x] = f1[x]; L.F2p[x] = f2[x];
    }
    k9_store_trace<CELLS>(trace_row, lane, tb);
    if (rowbest > 0) {
        const int j = jbase + k0 + rowx, sum = i + j;
        if (rowbest > L.best || (rowbest == L.best && (sum < L.best_sum || (sum == L.best_sum && i < L.best_i)))) {
            L.best = rowbest; L.best_sum = sum; L.best_i = i; L.best_j = j;
        }
    }
}

template <int CELLS>
__global__ void __launch_bounds__(128) k9_affine_local(const AffineParams p) {
    __shared__ uint8_t lut[256];
    fill_code_lut(lut);
    __syncthreads();
    constexpr int STRIDE = 32 * CELLS;  // trace bytes per row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    uint8_t *tr = p.trace + static_cast<size_t>(slot) * p.slot_bytes;
    const int k0 = lane * CELLS;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(p.next_pair, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= p.n_pairs) break;
        const AffinePairDev pr = p.pairs[q];
        const uint8_t *T = p.tbases + pr.t_off, *P = p.pbases + pr.p_off;
        const int m = pr.m, n = pr.n, c = pr.centre, W = pr.band, nb = 2 * W + 1;
        uint8_t *trace = tr + pr.trace_off;
        K9Lane<CELLS> L;
#pragma unroll
        for (int x = 0; x < CELLS; ++x) {
            L.Hp[x] = 0; L.F1p[x] = AFF_NEG; L.F2p[x] = AFF_NEG;
            const int j = 1 + c - W + k0 + x;  // row 1
            L.tc[x] = (j >= 1 && j <= n) ? lut[__ldg(T + j - 1)] : 4u;
        }
        L.best = 0; L.best_sum = 0x7FFFFFFF; L.best_i = 0; L.best_j = 0;
        // lane l of a 32-row block fetches the pattern base of row i0 + l and the text base that enters the band in that row
        // (under the last cell of lane 31), one block ahead
        auto fetch = [&](int i0) -> uint32_t {
            const int row = i0 + lane, j = row + c - W + 32 * CELLS - 1;
            const uint32_t pb = row <= m ? __ldg(P + row - 1) : static_cast<uint32_t>('N');
            const uint32_t tbyte = (row <= m && j >= 1 && j <= n) ? __ldg(T + j - 1) : static_cast<uint32_t>('N');
            return pb | (tbyte << 8);
        };
        uint32_t nxt = fetch(1);
        for (int i0 = 1; i0 <= m; i0 += 32) {
            const uint32_t cur = static_cast<uint32_t>(lut[nxt & 0xffu]) | (static_cast<uint32_t>(lut[nxt >> 8]) << 8);
            nxt = fetch(i0 + 32);
            const int rows = min(32, m - i0 + 1);
            for (int r = 0; r < rows; ++r) {
                const int i = i0 + r;
                const uint32_t pk = __shfl_sync(0xffffffffu, cur, r);
                if (i > 1) {  // the band moves one column to the right
                    const uint32_t tn = __shfl_down_sync(0xffffffffu, L.tc[0], 1);
#pragma unroll
                    for (int x = 0; x + 1 < CELLS; ++x) L.tc[x] = L.tc[x + 1];
                    L.tc[CELLS - 1] = lane == 31 ? (pk >> 8) : tn;
                }
                const int jbase = i + c - W;  // column of band cell 0
                uint8_t *trow = trace + static_cast<size_t>(i - 1) * STRIDE;
                if (jbase >= 1 && jbase + nb - 1 <= n) k9_row<CELLS, true>(p, L, i, pk & 0xffu, jbase, n, nb, lane, trow);
                else k9_row<CELLS, false>(p, L, i, pk & 0xffu, jbase, n, nb, lane, trow);
            }
        }
        int best = L.best, best_sum = L.best_sum, best_i = L.best_i, best_j = L.best_j;
        // warp arg-max: (score desc, i + j asc, i asc)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, d), os = __shfl_xor_sync(0xffffffffu, best_sum, d);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, d), oj = __shfl_xor_sync(0xffffffffu, best_j, d);
            if (ob > best || (ob == best && ob > 0 && (os < best_sum || (os == best_sum && oi < best_i)))) { best = ob; best_sum = os; best_i = oi; best_j = oj; }
        }
        __threadfence_block();
        __syncwarp();
        // walk back, the whole warp in step: every lane holds the same (i, j, state); lane 0 writes the run-length CIGAR backwards
        AlignRecDev rec = {m, 0, 0, 0, 0, 0, 0, 0, 0};
        int ncig = 0;
        long long pos = pr.cig_off + pr.cig_len;
        if (best > 0) {
            int i = best_i, j = best_j, state = 0, nm = 0;
            uint32_t cur_op = 0, cur_len = 0;
            auto emit = [&](uint32_t op, uint32_t len) {
                if (op == cur_op) { cur_len += len; return; }
                if (cur_len) { --pos; if (lane == 0) p.cigar[pos] = (cur_len << 4) | cur_op; }
                cur_op = op; cur_len = len;
            };
            while (i > 0 && j > 0) {
                const int k = j - (i + c - W);
                if (k < 0 || k >= nb) break;  // left the band: the alignment starts here
                if (state == 0) {
                    // lane l: cell (i - l, j - l), same band index
                    const bool ok = i - lane >= 1 && j - lane >= 1;
                    uint32_t src = 5u;
                    bool eq = false;
                    if (ok) {
                        src = __ldcg(trace + static_cast<size_t>(i - lane - 1) * STRIDE + k) & 7u;
                        const uint32_t pc = lut[P[i - lane - 1]], tc = lut[T[j - lane - 1]];
                        eq = pc < 4u && pc == tc;
                    }
                    const uint32_t stop = __ballot_sync(0xffffffffu, src != 0u);
                    const uint32_t eqm = __ballot_sync(0xffffffffu, eq);
                    const int run = stop ? __ffs(stop) - 1 : 32;
                    uint32_t bits = eqm;
                    for (int rem = run; rem > 0;) {
                        const uint32_t b = bits & 1u, flip = b ? ~bits : bits;
                        const int len = min(rem, flip ? __ffs(flip) - 1 : 32);
                        emit(b ? CIG_EQ : CIG_X, static_cast<uint32_t>(len));
                        nm += b ? 0 : len;
                        bits = len < 32 ? bits >> len : 0u;
                        rem -= len;
                    }
                    i -= run; j -= run;
                    if (run < 32) {
                        const uint32_t s = __shfl_sync(0xffffffffu, src, run);
                        const bool okr = __shfl_sync(0xffffffffu, static_cast<int>(ok), run) != 0;
                        if (!okr) continue;  // row 0 or column 0 reached
                        if (s == 5u) break;
                        state = static_cast<int>(s);
                    }
                } else if (state == 1 || state == 3) {  // deletion: text bases without a pattern base; lane l: cell (i, j - l)
                    const bool ok = k - lane >= 0 && j - lane >= 1;
                    bool ext = false;
                    if (ok) {
                        const uint32_t t = __ldcg(trace + static_cast<size_t>(i - 1) * STRIDE + (k - lane));
                        ext = state == 1 ? (t >> 3) & 1u : (t >> 5) & 1u;
                    }
                    const uint32_t stop = __ballot_sync(0xffffffffu, !ok || !ext);
                    const int f = stop ? __ffs(stop) - 1 : 32;
                    const bool closes = f < 32 && __shfl_sync(0xffffffffu, static_cast<int>(ok), f & 31) != 0;
                    const int steps = f < 32 ? (closes ? f + 1 : f) : 32;
                    if (steps > 0) emit(CIG_D, static_cast<uint32_t>(steps));
                    nm += steps; j -= steps;
                    if (closes) state = 0;
                } else {  // insertion: pattern bases without a text base; lane l: cell (i - l, j)
                    const bool ok = i - lane >= 1 && k + lane < nb;
                    bool ext = false;
                    if (ok) {
                        const uint32_t t = __ldcg(trace + static_cast<size_t>(i - lane - 1) * STRIDE + (k + lane));
                        ext = state == 2 ? (t >> 4) & 1u : (t >> 6) & 1u;
                    }
                    const uint32_t stop = __ballot_sync(0xffffffffu, !ok || !ext);
                    const int f = stop ? __ffs(stop) - 1 : 32;
                    const bool closes = f < 32 && __shfl_sync(0xffffffffu, static_cast<int>(ok), f & 31) != 0;
                    const int steps = f < 32 ? (closes ? f + 1 : f) : 32;
                    if (steps > 0) emit(CIG_I, static_cast<uint32_t>(steps));
                    nm += steps; i -= steps;
                    if (closes) state = 0;
                }
            }
            if (cur_len) { --pos; if (lane == 0) p.cigar[pos] = (cur_len << 4) | cur_op; }
            rec.nm = nm; rec.p_start = i; rec.p_end = best_i; rec.t_start = j; rec.t_end = best_j;
            rec.dist = nm + (m - (best_i - i));
            ncig = static_cast<int>(pr.cig_off + pr.cig_len - pos);
        }
        __threadfence_block();
        __syncwarp();
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
