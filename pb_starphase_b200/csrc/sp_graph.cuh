// sp_graph.cuh -- K8: end-to-end alignment of sequences to variant graphs (row N3 of SURVEY.md 8f).
//
// Cyp2d6Extractor::assign_haplotype (src/cyp2d6/haplotyper.rs:371-468 of the reference) aligns every consensus to a graph of the
// CYP2D6 backbone with one bubble per database variant (hiphase's WFAGraph) and reads the allele of every variant off the nodes
// the optimal alignments pass through.  K8 is the forward half: the unit-cost DP over the graph, one warp per (graph, sequence)
// problem.  The graph arrives linearised: position = one character, predecessors = the previous character of its node or the
// last characters of the predecessor nodes (-1 = the start column), `diag` = the row the position is expected to align to.
//     D[pos][i] = min( min over preds q of { D[q][i-1] + (g[pos] != s[i-1]),  D[q][i] + 1 },  D[pos][i-1] + 1 )
// for the rows |i - diag[pos]| <= W.  Inside a node the predecessor is the previous position and the band moves down by one
// row, so the diagonal neighbour is the lane's own previous value and the horizontal one the next cell (one shuffle): no memory
// traffic except the column store; at node starts the predecessor columns are read back from the column matrix.  The vertical
// term is the same min-plus prefix scan as in K7.  Every column is written out: the host walks the matrix backwards to mark the
// cells, and with them the nodes, that lie on optimal alignments (host/sp_host_graph.cpp).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sp {

constexpr int GRAPH_INF = 0x3FFFFFFF;

struct GraphParams {
    const uint8_t *gcodes;       // graph characters as codes 0..3, 4 = other
    const long long *goff;       // [n_problems + 1] positions of each problem
    const int32_t *pred_off;     // [total_positions + 1]
    const int32_t *preds;        // problem-local position ids, -1 = start column
    const int32_t *diag;         // [total_positions]
    const int32_t *end_off;      // [n_problems + 1]
    const int32_t *ends;
    const uint8_t *scodes;       // sequences as codes
    const long long *soff;       // [n_problems + 1]
    int32_t *columns;            // [total_positions][2W + 1]
    int32_t *score;              // [n_problems]
    int n_problems, W;
};

template <int CELLS>
__global__ void __launch_bounds__(128) k8_graph_forward(const GraphParams p) {
    const int prob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (prob >= p.n_problems) return;
    const int nb = 2 * p.W + 1, W = p.W;
    const long long g0 = p.goff[prob];
    const int npos = static_cast<int>(p.goff[prob + 1] - g0);
    const uint8_t *S = p.scodes + p.soff[prob];
    const int m = static_cast<int>(p.soff[prob + 1] - p.soff[prob]);
    const int k0 = lane * CELLS;
    int e[CELLS];  // column of the previous position (valid when the fast path applies)
    int prev_diag = 0, prev_pos = -2;
    // value of row i in the column of predecessor q (-1 = start column: D = i for i <= W)
    auto col_at = [&](int q, int i) -> int {
        if (i < 0 || i > m) return GRAPH_INF;
        if (q < 0) return i <= W ? i : GRAPH_INF;
        const int k = i - p.diag[g0 + q] + W;
        if (k < 0 || k >= nb) return GRAPH_INF;
        return p.columns[(g0 + q) * nb + k];
    };
    for (int pos = 0; pos < npos; ++pos) {
        const int d = p.diag[g0 + pos];
        const uint32_t ch = p.gcodes[g0 + pos];
        const int pb = p.pred_off[g0 + pos], pe = p.pred_off[g0 + pos + 1];
        const bool fast = pe - pb == 1 && p.preds[pb] == prev_pos && d == prev_diag + 1 && prev_pos >= 0;
        int nw[CELLS];
        if (fast) {
            // diagonal neighbour = own previous value (same band index), horizontal = the next band cell of the previous column
            const int from_next_lane = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
            for (int c = 0; c < CELLS; ++c) {
                const int k = k0 + c, i = d - W + k;
                int v = GRAPH_INF;
                if (k < nb && i >= 0 && i <= m) {
                    const int horiz = c + 1 < CELLS ? e[c + 1] : (lane < 31 ? from_next_lane : GRAPH_INF);
                    const int hz = (k + 1 < nb && horiz < GRAPH_INF) ? horiz + 1 : GRAPH_INF;
                    int dg = GRAPH_INF;
                    if (i > 0 && e[c] < GRAPH_INF) dg = e[c] + ((ch < 4u && S[i - 1] == ch) ? 0 : 1);
                    v = min(hz, dg);
                }
                nw[c] = v;
            }
        } else {
            __syncwarp();  // the columns this position reads were written by this warp
#pragma unroll
            for (int c = 0; c < CELLS; ++c) {
                const int k = k0 + c, i = d - W + k;
                int v = GRAPH_INF;
                if (k < nb && i >= 0 && i <= m) {
                    for (int q = pb; q < pe; ++q) {
                        const int pr = p.preds[q];
                        const int h = col_at(pr, i);
                        if (h < GRAPH_INF) v = min(v, h + 1);
                        if (i > 0) {
                            const int dg = col_at(pr, i - 1);
                            if (dg < GRAPH_INF) v = min(v, dg + ((ch < 4u && S[i - 1] == ch) ? 0 : 1));
                        }
                    }
                }
                nw[c] = v;
            }
        }
        // vertical: x[k] = v[k] - k, inclusive prefix min, back to values
        int run = GRAPH_INF;
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int x = nw[c] < GRAPH_INF ? nw[c] - (k0 + c) : GRAPH_INF;
            run = min(run, x);
            nw[c] = run;
        }
        int carry = run;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, carry, dd);
            if (lane >= dd) carry = min(carry, up);
        }
        int before = __shfl_up_sync(0xffffffffu, carry, 1);
        if (lane == 0) before = GRAPH_INF;
#pragma unroll
        for (int c = 0; c < CELLS; ++c) {
            const int k = k0 + c, i = d - W + k;
            const int x = min(nw[c], before);
            e[c] = (k < nb && i >= 0 && i <= m && x < GRAPH_INF) ? x + k : GRAPH_INF;
            if (k < nb) p.columns[(g0 + pos) * nb + k] = e[c];
        }
        prev_diag = d; prev_pos = pos;
    }
    __syncwarp();
    // best end: min over the end positions of D[end][m]
    int best = GRAPH_INF;
    for (int q = p.end_off[prob] + lane; q < p.end_off[prob + 1]; q += 32) best = min(best, col_at(p.ends[q], m));
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, dd));
    if (lane == 0) p.score[prob] = best;
}

}  // namespace sp
