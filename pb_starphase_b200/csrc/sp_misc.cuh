// sp_misc.cuh -- K5 (row top-k) and K6 (allele-vector match) kernels (sm_100a only).
#pragma once
#include "sp_kernels.cuh"

namespace sp {

// ------------------------------------------------------------------------------------------
// K5: the k best patterns of every text (row top-k of the distance matrix, optionally of distance + a per-pattern bias),
// ties by lower pattern index.
// Stands in for minimap2's best_n hit list at the realigner (src/hla/realigner.rs:116-146): only these candidates
// go on to the traceback, and only R x k records cross PCIe instead of the R x A matrix.
// One thread per text; the allele-major layout D[p * ld + t] makes the loads of a warp contiguous.
// ------------------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(128) k5_row_topk(const T *__restrict__ D, long long ld, int nt, int np, int k,
                                                   const int32_t *__restrict__ bias,  // optional per-pattern addend of the ranking key
                                                   uint32_t weight,                   // multiplier of the distance inside the ranking key
                                                   int32_t *__restrict__ idx, int32_t *__restrict__ dist) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    uint32_t bd[K], bi[K];
#pragma unroll
    for (int q = 0; q < K; ++q) { bd[q] = 0xFFFFFFFFu; bi[q] = 0xFFFFFFFFu; }
    for (int a = 0; a < np; ++a) {
        const uint32_t v = static_cast<uint32_t>(D[static_cast<long long>(a) * ld + t]) * weight + (bias ? static_cast<uint32_t>(bias[a]) : 0u);
        if (v < bd[K - 1]) {  // strict: on ties the earlier pattern stays ahead
            bd[K - 1] = v; bi[K - 1] = static_cast<uint32_t>(a);
#pragma unroll
            for (int q = K - 1; q > 0; --q)
                if (bd[q] < bd[q - 1]) {
                    const uint32_t xd = bd[q], xi = bi[q];
                    bd[q] = bd[q - 1]; bi[q] = bi[q - 1];
                    bd[q - 1] = xd; bi[q - 1] = xi;
                }
        }
    }
#pragma unroll
    for (int q = 0; q < K; ++q)
        if (q < k) {
            const bool have = bi[q] != 0xFFFFFFFFu;
            idx[static_cast<long long>(t) * k + q] = have ? static_cast<int32_t>(bi[q]) : -1;
            dist[static_cast<long long>(t) * k + q] = have ? static_cast<int32_t>(D[static_cast<long long>(bi[q]) * ld + t]) : -1;
        }
}

// ------------------------------------------------------------------------------------------
// K6: allele-vector match (src/cyp2d6/haplotyper.rs:470-517).  Rows of 0/1/2/3 site states become bit planes
// (one warp per row, __ballot_sync per 32 sites); a (sequence, haplotype) pair is then W = ceil(V / 32) words of
// AND / OR / POPC.  is_match = (seq == hap) for seq in {0, 1}, true for 2, false for 3.
// ------------------------------------------------------------------------------------------
// planes[row][k][w]: k = 0 (state == 0), 1 (state == 1), 2 (state == 2); a value above `max_state` sets *bad
__global__ void k6_pack_states(const uint8_t *__restrict__ states, int n_rows, int n_var, int W, int n_planes, int max_state,
                               uint32_t *__restrict__ planes, int *__restrict__ bad) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    for (int w = 0; w < W; ++w) {
        const int v = w * 32 + lane;
        const int st = v < n_var ? states[static_cast<size_t>(row) * n_var + v] : 255;
        if (v < n_var && st > max_state) atomicExch(bad, 1);
        for (int k = 0; k < n_planes; ++k) {
            const uint32_t bits = __ballot_sync(0xffffffffu, st == k);
            if (lane == 0) planes[(static_cast<size_t>(row) * n_planes + k) * W + w] = bits;
        }
    }
}

__global__ void __launch_bounds__(256) k6_variant_match(const uint32_t *__restrict__ seq_planes, const uint32_t *__restrict__ hap_planes,
                                                         const uint32_t *__restrict__ vi_plane, int n_seq, int n_hap, int W,
                                                         uint32_t *__restrict__ vi_match, uint32_t *__restrict__ all_match) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(n_seq) * n_hap) return;
    const int s = static_cast<int>(t / n_hap), h = static_cast<int>(t - static_cast<long long>(s) * n_hap);
    const uint32_t *s0 = seq_planes + static_cast<size_t>(s) * 3 * W, *s1 = s0 + W, *s2 = s1 + W;
    const uint32_t *h1 = hap_planes + (static_cast<size_t>(h) * 2 + 1) * W, *h0 = h1 - W;
    uint32_t all = 0, vi = 0;
    for (int w = 0; w < W; ++w) {
        const uint32_t m = (s0[w] & h0[w]) | (s1[w] & h1[w]) | s2[w];
        all += __popc(m);
        vi += __popc(m & vi_plane[1 * W + w]);
    }
    all_match[t] = all;
    vi_match[t] = vi;
}

// ------------------------------------------------------------------------------------------
// Derived text sets (row N4 of SURVEY.md 8f): spliced / reverse-complemented sequences built on the device from a resident set.
// Output sequence q = the concatenation of its intervals of source sequence src_index[q], reverse-complemented as a whole when
// revcomp[q] is set.  One thread per output base: two binary searches (sequence, interval) and one byte.
// ------------------------------------------------------------------------------------------
struct DeriveParams {
    const uint8_t *src_bases;
    const long long *src_offs;
    const int32_t *src_index;   // [n_out]
    const long long *iv_off;    // [n_out + 1] first interval of every output sequence
    const int32_t *iv_begin;    // [n_iv]
    const long long *iv_prefix; // [n_iv] output bases of the sequence before this interval (forward order)
    const uint8_t *revcomp;     // [n_out]
    const long long *out_offs;  // [n_out + 1]
    uint8_t *out_bases;
    long long n_out, total;
};

__device__ __forceinline__ uint8_t complement_base(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;  // N and anything else stays
    }
}

__global__ void __launch_bounds__(256) derive_texts(const DeriveParams p) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    long long lo = 0, hi = p.n_out;  // the sequence that holds output base i: last q with out_offs[q] <= i
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (p.out_offs[mid] <= i) lo = mid; else hi = mid;
    }
    const long long q = lo, len = p.out_offs[q + 1] - p.out_offs[q];
    const bool rc = p.revcomp[q] != 0;
    const long long pos = rc ? len - 1 - (i - p.out_offs[q]) : i - p.out_offs[q];
    long long a = p.iv_off[q], b = p.iv_off[q + 1];  // last interval whose prefix <= pos
    while (b - a > 1) {
        const long long mid = (a + b) >> 1;
        if (p.iv_prefix[mid] <= pos) a = mid; else b = mid;
    }
    const uint8_t c = p.src_bases[p.src_offs[p.src_index[q]] + p.iv_begin[a] + (pos - p.iv_prefix[a])];
    p.out_bases[i] = rc ? complement_base(c) : c;
}

}  // namespace sp
