// sp_consensus.cuh -- K7: batched extension of growing consensus sequences against a read set (row N1 of SURVEY.md 8f).
//
// The reference builds its HLA / CYP2D6 consensuses with waffle_con's dynamic-WFA search (src/hla/caller.rs:1097-1219,
// :727-755; src/cyp2d6/caller.rs:145-280): candidate consensus prefixes are extended one symbol at a time and every read keeps
// its edit distance to the growing prefix; reads vote for the next symbol with the bases that follow their best-scoring
// prefixes.  The data-parallel part of that search is "extend candidate X by symbol s for every read", and that is this
// kernel: one warp per (task, read).  A track holds, for every read, the last DP column of
//     E[i] = edit distance between the consensus so far and the read prefix r[0, i)
// restricted to a band of 2W + 1 rows around the read's nominal diagonal i = L - offset (HiFi reads drift by a few indels; the
// window of uncertainty of the offset widens the band).  Appending symbol s turns column L into column L + 1:
//     E'[i] = min(E[i-1] + (r[i-1] != s),  E[i] + 1,  E'[i-1] + 1)
// In band coordinates (k = i - (L - offset) + W) the diagonal neighbour keeps its k, the horizontal one is k + 1, and the
// vertical dependency is a min-plus prefix scan over k (per lane sequentially, across lanes with five shuffles).
// Row 0 (no read base used yet) costs max(0, L - (offset + window / 2)): the read may start anywhere inside its offset window
// for free.  A read becomes active at L = max(0, offset - window / 2); before that it has no cost and no vote.
// Per (task, read) the kernel reports ed = min_i E'[i] (end-free in the read: it continues), the set of next read bases at
// the rows reaching that minimum (the read's vote), whether the read is consumed at such a row, and the running minimum of
// E'[|r|] over all columns so far (the read's final cost once the consensus has passed its end).
//
// Integer DP on the INT32 pipe; the work per step is tiny (reads x band cells), so the kernel is written for batches: many
// tasks -- candidate extensions of many consensus problems -- per launch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sp {

constexpr int CONS_INF = 0x3FFFFFFF;
constexpr uint8_t VOTE_FINISHED = 1u << 5;  // the read is consumed at a row reaching the minimum
constexpr uint8_t VOTE_INACTIVE = 1u << 6;  // the consensus has not reached the read's window yet

struct ConsParams {
    const uint8_t *codes;      // read bases as codes 0..3 (ACGT), 4 = other; concatenated
    const long long *roffs;    // [n_reads + 1]
    const int32_t *offset;     // [n_reads] nominal start of the read inside the consensus; < 0 = exactly at its start, no window
    int32_t *band;             // [n_tracks][n_reads][band_cells] last column, band coordinates
    int32_t *best_full;        // [n_tracks][n_reads] min over columns so far of E[|r|]
    int32_t *track_len;        // [n_tracks] consensus length of the track
    const int32_t *src, *dst;  // [n_tasks]
    const uint8_t *sym;        // [n_tasks] code of the appended symbol (0..4), 255 = report only
    int32_t *out_ed;           // [n_tasks][n_reads]
    uint8_t *out_votes;        // [n_tasks][n_reads]
    int32_t *out_full;         // [n_tasks][n_reads]
    int n_reads, n_tasks, W, half_window;
};

// one warp per (task, read); CELLS = ceil((2W + 1) / 32) consecutive band cells per lane
template <int CELLS>
__global__ void __launch_bounds__(128) k7_extend(const ConsParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.n_tasks * p.n_reads) return;
    const int task = warp / p.n_reads, r = warp - task * p.n_reads;
    const int src = p.src[task], dst = p.dst[task];
    const uint32_t s = p.sym[task];
    const bool extend = s != 255u;
    const int nb = 2 * p.W + 1;
    const int L0 = p.track_len[src];        // consensus length before the step
    const int L = L0 + (extend ? 1 : 0);    // after
    const int off_raw = p.offset[r];             // < 0: the read starts exactly at the consensus start (add_sequence without offset)
    const int off = max(off_raw, 0);
    const int hw = off_raw < 0 ? 0 : p.half_window;
    const int m = static_cast<int>(p.roffs[r + 1] - p.roffs[r]);
    const uint8_t *R = p.codes + p.roffs[r];
    const int32_t *old = p.band + (static_cast<size_t>(src) * p.n_reads + r) * nb;
    int32_t *nw = p.band + (static_cast<size_t>(dst) * p.n_reads + r) * nb;
    const size_t o = static_cast<size_t>(task) * p.n_reads + r;
    const int start = max(0, off - hw);  // first column at which the read is active
    int bf = p.best_full[static_cast<size_t>(src) * p.n_reads + r];

    if (L < start) {  // not active yet: the state stays "column `start` not reached"
        if (lane == 0) {
            p.out_ed[o] = 0; p.out_votes[o] = VOTE_INACTIVE; p.out_full[o] = CONS_INF;
            p.best_full[static_cast<size_t>(dst) * p.n_reads + r] = CONS_INF;
        }
        return;
    }
    int e[CELLS];
    const int k0 = lane * CELLS;
    // row index of band cell k at column L:  i = (L - off) + (k - W)
    const int ibase = (L - off) - p.W;
    // column L0 of the parent track; the activating column (L0 == start) is E[i] = i by definition -- read bases before the
    // window start are unaligned -- and is synthesised instead of stored, so a fresh track needs no initial columns
    auto old_at = [&](int k) -> int {
        if (k < 0 || k >= nb) return CONS_INF;
        const int i_old = (L0 - off) + (k - p.W);
        if (i_old < 0 || i_old > m) return CONS_INF;
        return L0 == start ? i_old : old[k];
    };
    if (!extend) {  // report the state of column L0 (>= start here)
#pragma unroll
        for (int c = 0; c < CELLS; ++c) e[c] = old_at(k0 + c);
    } else if (L == start) {  // this step reaches the read's window: the activating column
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int k = k0 + c, i = ibase + k;
            e[c] = (k < nb && i >= 0 && i <= m) ? i : CONS_INF;
        }
    } else {
        // base[k] = min(diagonal, horizontal); then the vertical min-plus scan
        const int row0_cost = max(0, L - (off + hw));  // E'[0]: the read may start anywhere inside its window for free
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int k = k0 + c, i = ibase + k;
            int v = CONS_INF;
            if (k < nb && i >= 0 && i <= m) {
                if (i == 0) {
                    v = row0_cost;
                } else {
                    const int diag = old_at(k);       // E[i-1] of column L0 has the same band index
                    const int horiz = old_at(k + 1);  // E[i]   of column L0
                    const int sub = (s < 4u && R[i - 1] == s) ? 0 : 1;
                    v = min(diag < CONS_INF ? diag + sub : CONS_INF, horiz < CONS_INF ? horiz + 1 : CONS_INF);
                }
            }
            e[c] = v;
        }
        // x[k] = e[k] - k; inclusive prefix min over k; e[k] = prefmin + k
        int run = CONS_INF;
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int x = e[c] < CONS_INF ? e[c] - (k0 + c) : CONS_INF;
            run = min(run, x);
            e[c] = run;  // lane-local prefix min of x for now
        }
        int carry = run;  // lane total, then inclusive scan across lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, carry, d);
            if (lane >= d) carry = min(carry, up);
        }
        int prev = __shfl_up_sync(0xffffffffu, carry, 1);  // prefix min of all lower lanes
        if (lane == 0) prev = CONS_INF;
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int k = k0 + c, i = ibase + k;
            const int x = min(e[c], prev);
            e[c] = (k < nb && i >= 0 && i <= m && x < CONS_INF) ? x + k : CONS_INF;
        }
    }
    // store the column, reduce: minimum, votes at the minimum, distance with the read consumed
    int mn = CONS_INF, full = CONS_INF;
#pragma unroll
    for (int c = 0; c < CELLS; ++c) {
        const int k = k0 + c, i = ibase + k;
        if (k < nb) nw[k] = e[c];
        mn = min(mn, e[c]);
        if (k < nb && i == m) full = e[c];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        full = min(full, __shfl_xor_sync(0xffffffffu, full, d));
    }
    bf = min(bf, full);
    uint32_t votes = 0;
#pragma unroll
    for (int c = 0; c < CELLS; ++c) {
        const int k = k0 + c, i = ibase + k;
        if (k < nb && e[c] == mn && mn < CONS_INF && i >= 0 && i <= m) votes |= i == m ? VOTE_FINISHED : (1u << R[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) votes |= __shfl_xor_sync(0xffffffffu, votes, d);
    if (lane == 0) {
        p.out_ed[o] = mn; p.out_votes[o] = static_cast<uint8_t>(votes); p.out_full[o] = bf;
        p.best_full[static_cast<size_t>(dst) * p.n_reads + r] = bf;
    }
}

// fresh tracks: empty consensus, no read consumed yet
__global__ void k7_reset(int32_t *best_full, int32_t *track_len, int track, int n_reads) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) best_full[static_cast<size_t>(track) * n_reads + r] = CONS_INF;
    if (r == 0) track_len[track] = 0;
}

// sets track_len[dst] = track_len[src] + 1 (or keeps it for report-only tasks) after the extension kernel has read the old lengths
__global__ void k7_bump_lengths(const ConsParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_tasks) return;
    p.track_len[p.dst[t]] = p.track_len[p.src[t]] + (p.sym[t] != 255u ? 1 : 0);
}

}  // namespace sp
