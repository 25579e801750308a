// sp_kernels.cuh -- device code of libstarphase_gpu.so (sm_100a only).
//
// K1  k1_infix      batched infix edit distance, pattern-stationary systolic warps
//                   (replaces the per-allele minimap2 calls of src/hla/caller.rs:1413-1500,
//                    src/hla/realigner.rs:116-146, src/cyp2d6/chaining.rs:48-94,
//                    src/cyp2d6/haplotyper.rs:193-249 of the reference)
// K2  k2_pair_minsum  S[i,j] = sum_r min(D[r,i], D[r,j]) + per-CTA top-k by (S,i,j)
//                   (north_star pair scoring; CYP2D6 form of src/cyp2d6/chaining.rs:409-534)
// plus the pack kernels that turn ASCII sequences into the device formats.
//
// See DESIGN.md §4 for the layouts and the op-count model.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sp {

// ------------------------------------------------------------------------------------------
// Geometry shared by host and device
// ------------------------------------------------------------------------------------------
constexpr int K1_WARPS = 8;                 // warps (= pattern bins) per CTA
constexpr int K1_THREADS = K1_WARPS * 32;
constexpr int K1_CHUNK = 8;                 // text columns per chunk (one uint2 of byte codes)
constexpr int PEQ_ROWS = 6;                 // A,C,G,T, N(other), pv-init
__host__ __device__ constexpr int blob_words(int U) { return PEQ_ROWS * 32 * U + 64; }
// One mask row of a bin is 32 lanes x U words, stored as word groups so that each lane fetches its U words with the
// fewest conflict-free shared-memory loads: U/4 groups of four (LDS.128), then a pair (LDS.64) if U % 4 >= 2, then a
// single word (LDS.32) if U is odd.  Inside a group the lanes are contiguous.
__host__ __device__ constexpr int grp_start(int U, int u) {
    return u < (U / 4) * 4 ? (u / 4) * 4 : ((U % 4) >= 2 && u < (U / 4) * 4 + 2) ? (U / 4) * 4 : (U / 4) * 4 + ((U % 4) >= 2 ? 2 : 0);
}
__host__ __device__ constexpr int grp_width(int U, int u) {
    return u < (U / 4) * 4 ? 4 : ((U % 4) >= 2 && u < (U / 4) * 4 + 2) ? 2 : 1;
}
// word index of (lane, u) inside one mask row
__host__ __device__ constexpr int row_word(int U, int lane, int u) {
    return 32 * grp_start(U, u) + lane * grp_width(U, u) + (u - grp_start(U, u));
}
// info1 bit layout
constexpr uint32_t INFO_FIRST = 1u << 30;
constexpr uint32_t INFO_LAST = 1u << 31;
constexpr uint32_t INFO_LEN_MASK = (1u << 30) - 1;
constexpr uint32_t NO_PATTERN = 0xFFFFFFFFu;

struct K1Params {
    const uint32_t *blobs;          // [n_groups][K1_WARPS][blob_words(U)]
    const uint2 *text;              // chunk stream of all tiles
    const int32_t *tile_chunk_off;  // [n_tiles + 1], even
    const int32_t *tile_text0;      // [n_tiles] first text index of the tile
    void *out;                      // D[p * ld + t], u16 or i32
    int32_t *out_end;               // end columns (same layout, i32) or nullptr
    long long ld;
    int n_groups, n_tiles;
    int out16;
    int prefix_mode;                // SP_PREFIX: top row delta +1
    uint32_t one, m1, sixteen;      // +1, -1 (0xFFFFFFFF), 16: passed at run time so `x * one + y` stays an IMAD (FMA pipe)
    int *next_item;                 // work counter (zeroed by the host before the launch): items are handed out dynamically;
                                    // nullptr: one CTA per item (grid = items)
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy executed by the TMA unit; completes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#include "sp_addchain.inc"

// A,C,G,T (either case) -> 0..3; everything else (N, *, ...) -> 4 = matches nothing
__device__ __forceinline__ uint32_t base_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// ------------------------------------------------------------------------------------------
// K1: one text column for one lane (32*U pattern rows as a U-word big integer).
// Myers 1999 with Hyyro's hin/hout at lane boundaries only; inside the lane the U words are
// linked by the adder carry (IADD3.X) and funnel shifts (SHF.L.W), 10 ALU ops per word.
// ------------------------------------------------------------------------------------------
template <int U>
__device__ __forceinline__ void load_row(const uint32_t *row, int lane, uint32_t (&v)[U]) {
    constexpr int N4 = U / 4;
#pragma unroll
    for (int q = 0; q < N4; ++q) {
        const uint4 x = *reinterpret_cast<const uint4 *>(row + q * 128 + lane * 4);
        v[4 * q + 0] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
    if (U % 4 >= 2) {
        const uint2 x = *reinterpret_cast<const uint2 *>(row + N4 * 128 + lane * 2);
        v[4 * N4 + 0] = x.x; v[4 * N4 + 1] = x.y;
    }
    if (U % 2) v[U - 1] = row[32 * (U - 1) + lane];
}

// K1 inner step, "arithmetic form".
// The ALU pipe (LOP3 / IADD3 / SHF: 64 lanes/clk/SM) bounds this kernel while the FMA pipe idles
// (profiles/r01_k1_ncu_full.txt: alu 98 %, fma 2 %), and a full-rate IMAD issues for free beside a
// saturated ALU pipe (profiles/r01_pipe_bench.txt: 8 LOP3 = 16.1 cycles, 8 LOP3 + 8 IMAD = 16.3).
// Three of Myers' boolean operations are therefore computed as integer add/sub on the FMA pipe, using
// disjointness / subset facts of the delta vectors (Hyyro's D0 form; npv = ~Pv is the stored state), and the
// horizontal-minus vector is not computed at all:
//     t   = Eq & Pv,  sum = t + Pv + [hin < 0]      LOP3, IADD3.X (borrow chain t - npv - 1 + carry, SubChainY)
//     C   = sum ^ t ^ Pv                             LOP3   the carry-IN vector of that addition.  Its carry-OUT vector is
//                                                           t | (C & Pv & ~Eq) = Pv & D0 = Mh, so C IS Mh shifted up one row
//                                                           (with hin < 0 in bit 0): Mhs costs neither arithmetic nor a shift
//     D0  = C | Eq | Mv                              LOP3
//     nG  = ~(D0 | Pv)                               LOP3        G' = D0 | Phs                 LOP3
//     Ph  = Mv | nG   = Mv + nG   (Mv is a subset of D0)                                      IMAD
//     Mv' = Phs & D0  = D0 + Phs - G'                                                         2 IMAD
//     ~Pv' = ~(Mhs | ~G') = G' - C            (C is a subset of D0)                           IMAD
// leaving per 32-row word 5 LOP3 + 1 IADD3.X + 1 SHF (Phs) on the ALU pipe (the textbook form: 7 + 1 + 2) and 4 IMAD on
// the FMA pipe; Mh itself is formed for the lane's last word only (2 IMAD), whose top bit is the delta that leaves the lane.
// The multipliers +1 / -1 are kernel parameters so ptxas cannot turn the IMADs back into IADD3.
template <int U, bool TRACK_END, bool KEEP_D0 = false>
__device__ __forceinline__ void column_step(const uint32_t *peq, int lane, uint32_t code, uint32_t one, uint32_t m1,
                                            uint32_t (&npv)[U], uint32_t (&mv)[U], uint32_t &X, uint32_t &Y,
                                            uint32_t &cph, uint32_t &cmh, int &score, int &best, int &col,
                                            int &best_col, uint32_t *d0_keep = nullptr) {
    uint32_t eq[U], t[U], sum[U];
    load_row<U>(peq + code * (32 * U), lane, eq);
    // hin < 0 (Hyyro: the row above already paid for this column) enters as the carry-in of the Myers add (SubChainY takes it from
    // the top bit of Y) instead of being OR-ed into Eq bit 0: bit 0 of the carry-in vector is then that bit, the diagonal-zero
    // vector gets it through c[0], and words 1 .. U-1 receive the carry of the word below -- exactly the bit a shift of Mh would
    // have moved there
#pragma unroll
    for (int u = 0; u < U; ++u) t[u] = eq[u] & ~npv[u];
    SubChainY<U>::run(sum, t, npv, Y);
    uint32_t ph[U], c[U], d0[U];
    uint32_t mh_last = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        c[u] = ~(sum[u] ^ t[u] ^ npv[u]);          // carry-in vector = Mh shifted up one row  (LOP3)
        d0[u] = c[u] | eq[u] | mv[u];              //                                          (LOP3)
        const uint32_t ng = ~d0[u] & npv[u];       // ~(D0 | Pv)                               (LOP3)
        ph[u] = ng * one + mv[u];                  //                                          (IMAD)
        if (u == U - 1) mh_last = ng * one + (npv[u] * m1 + d0[u]);  // Mh of the lane's last word only: its top bit leaves the lane (2 IMAD)
        if (KEEP_D0) d0_keep[u] = d0[u];           // K4 keeps the diagonal-zero vector for the traceback
    }
    // horizontal delta of the lane's last row: carry for the next lane, score for a last lane
    cph = __funnelshift_l(ph[U - 1], cph, 1);
    cmh = __funnelshift_l(mh_last, cmh, 1);
    if (TRACK_END) {  // per-column score and arg-min; without end columns the caller does it per chunk from cph/cmh
        score += static_cast<int>(ph[U - 1] >> 31) - static_cast<int>(mh_last >> 31);
        ++col;
        if (score < best) { best = score; best_col = col; }
    }
#pragma unroll
    for (int u = U - 1; u >= 0; --u) {
        const uint32_t phs = __funnelshift_l(u ? ph[u - 1] : X, ph[u], 1);
        const uint32_t b1 = phs * one + d0[u];     // D0 + Phs          (IMAD)
        const uint32_t g = d0[u] | phs;            //                   (LOP3)
        mv[u] = g * m1 + b1;                       // Phs & D0          (IMAD)
        npv[u] = c[u] * m1 + g;                    // ~Pv' = G - Mhs    (IMAD)
    }
    X <<= 1;
    Y <<= 1;
}

// Score bookkeeping without end columns: the 8 horizontal deltas of a chunk arrive as two 4-column nibble pairs
// (plus bits, minus bits); a 256-entry table gives each nibble pair's total and its lowest running prefix, so the
// running score and its minimum advance once per chunk with two table reads (LSU) and a few adds (FMA pipe)
// instead of three ALU-pipe instructions per column.
struct ScoreLut {
    int8_t delta[256];   // index = plus4 | minus4 << 4, bit 3 = first column of the four
    int8_t minpre[256];  // lowest partial sum after 1..4 columns
};
__device__ __forceinline__ void fill_score_lut(ScoreLut *lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        int s = 0, mn = 127;
        for (int k = 3; k >= 0; --k) {
            s += ((i >> k) & 1) - ((i >> (4 + k)) & 1);
            mn = min(mn, s);
        }
        lut->delta[i] = static_cast<int8_t>(s);
        lut->minpre[i] = static_cast<int8_t>(mn);
    }
}

template <int U, bool TRACK_END>
__device__ __forceinline__ void k1_warp_run(const uint32_t *blob, const uint2 *s_text, int nch, int text0,
                                            const K1Params &p, const ScoreLut *lut) {
    const int lane = threadIdx.x & 31;
    const uint32_t pat = blob[PEQ_ROWS * 32 * U + lane];
    if (__all_sync(0xffffffffu, pat == NO_PATTERN)) return;  // filler warp of a partly used group
    const uint32_t info1 = blob[PEQ_ROWS * 32 * U + 32 + lane];
    const bool first = (info1 & INFO_FIRST) != 0;
    const bool last = (info1 & INFO_LAST) != 0;
    const int m = static_cast<int>(info1 & INFO_LEN_MASK);
    const uint32_t cin_first = p.prefix_mode ? 0x00FFu : 0u;

    uint32_t npv[U], mv[U];  // ~Pv, Mv
    load_row<U>(blob + 5 * (32 * U), lane, npv);
#pragma unroll
    for (int u = 0; u < U; ++u) { npv[u] = ~npv[u]; mv[u] = 0; }
    int score = m, best = m, col = 0, best_col = 0;
    uint32_t carry_out = 0;
    int tcount = 0;

    const int nsteps = nch + 31;
    for (int s = 0; s < nsteps; ++s) {
        uint32_t cin = __shfl_up_sync(0xffffffffu, carry_out, 1);
        if (first) cin = cin_first;
        const int idx = s - lane;
        if (static_cast<unsigned>(idx) < static_cast<unsigned>(nch)) {
            // one byte load per column (LSU pipe) instead of shift+mask on the ALU pipe, which is the bound
            // (inline PTX so the compiler does not fuse them back into one wide load plus eight extractions)
            const uint32_t tb = smem_u32(s_text + idx);
            uint32_t codes[K1_CHUNK];
            static_assert(K1_CHUNK == 8, "eight explicit byte loads below");
            asm("ld.shared.u8 %0, [%8];\n\tld.shared.u8 %1, [%8+1];\n\tld.shared.u8 %2, [%8+2];\n\tld.shared.u8 %3, [%8+3];\n\t"
                "ld.shared.u8 %4, [%8+4];\n\tld.shared.u8 %5, [%8+5];\n\tld.shared.u8 %6, [%8+6];\n\tld.shared.u8 %7, [%8+7];"
                : "=r"(codes[0]), "=r"(codes[1]), "=r"(codes[2]), "=r"(codes[3]), "=r"(codes[4]), "=r"(codes[5]),
                  "=r"(codes[6]), "=r"(codes[7])
                : "r"(tb));
            const uint32_t end_flag = codes[K1_CHUNK - 1] & 0x80u;
            codes[K1_CHUNK - 1] &= 0x7Fu;
            uint32_t X = cin << 24, Y = cin << 16;
            uint32_t cphA = 0, cmhA = 0, cphB = 0, cmhB = 0;  // columns 0-3 and 4-7
#pragma unroll
            for (int c = 0; c < K1_CHUNK; ++c)
                column_step<U, TRACK_END>(blob, lane, codes[c], p.one, p.m1, npv, mv, X, Y, c < 4 ? cphA : cphB,
                                          c < 4 ? cmhA : cmhB, score, best, col, best_col);
            const uint32_t iA = cmhA * p.sixteen + cphA, iB = cmhB * p.sixteen + cphB;  // IMAD: table indices
            carry_out = (cphA * p.sixteen + cphB) | ((iA & 0xF0u) << 8) | ((iB & 0xF0u) << 4);
            if (!TRACK_END) {
                const int dA = lut->delta[iA], mA = lut->minpre[iA], dB = lut->delta[iB], mB = lut->minpre[iB];
                const int sA = score + dA;
                best = min(best, min(score + mA, sA + mB));
                score = sA + dB;
            }
            if (end_flag) {  // last chunk of a text: emit + reset
                if (last) {
                    const long long o = static_cast<long long>(pat) * p.ld + (text0 + tcount);
                    if (p.out16) reinterpret_cast<uint16_t *>(p.out)[o] = static_cast<uint16_t>(best);
                    else reinterpret_cast<int32_t *>(p.out)[o] = best;
                    if (TRACK_END) p.out_end[o] = best_col;
                }
                ++tcount;
                load_row<U>(blob + 5 * (32 * U), lane, npv);
#pragma unroll
                for (int u = 0; u < U; ++u) { npv[u] = ~npv[u]; mv[u] = 0; }
                score = m; best = m; col = 0; best_col = 0;
            }
        }
    }
}

#ifndef SP_K1_MIN_BLOCKS
#define SP_K1_MIN_BLOCKS 2
#endif
// lane widths above 16 exist for patterns longer than 16,384 rows only (DRB1-sized genomic alleles): one CTA per SM, ~190 registers
__host__ __device__ constexpr int k1_min_blocks(int U) { return U <= 16 ? SP_K1_MIN_BLOCKS : 1; }

template <int U, bool TRACK_END>
__global__ void __launch_bounds__(K1_THREADS, k1_min_blocks(U)) k1_infix(const K1Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ ScoreLut lut;
    constexpr int BW = blob_words(U);
    uint32_t *s_blob = reinterpret_cast<uint32_t *>(smem_raw);
    uint2 *s_text = reinterpret_cast<uint2 *>(smem_raw + K1_WARPS * BW * 4);
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    fill_score_lut(&lut);
    __syncthreads();

    uint32_t parity = 0;
    const int n_items = p.n_groups * p.n_tiles;
    __shared__ int s_item;
    // items (pattern group x text tile) are handed out from an atomic counter: tiles hold whole texts and differ in length,
    // and with few reads per call (the cohort's 64 per gene) a static round-robin left the last round a fifth full
    // Without a counter (a context that shares its GPU, sp_ctx_share_device) the launch brings one CTA per item: the hardware's CTA
    // scheduler hands the items out, and every retiring CTA is a slot that a short kernel of another context can take
    for (bool once = true;; once = false) {
        if (p.next_item) {
            if (threadIdx.x == 0) s_item = atomicAdd(p.next_item, 1);
            __syncthreads();
        }
        const int item = p.next_item ? s_item : (once ? static_cast<int>(blockIdx.x) : n_items);
        if (item >= n_items) break;
        const int g = item / p.n_tiles;
        const int tile = item - g * p.n_tiles;
        const int c0 = p.tile_chunk_off[tile];
        const int nch = p.tile_chunk_off[tile + 1] - c0;
        if (threadIdx.x == 0) {
            fence_proxy_async();
            const uint32_t bytes_blob = K1_WARPS * BW * 4;
            const uint32_t bytes_text = static_cast<uint32_t>(nch) * 8u;
            mbar_expect_tx(&mbar, bytes_blob + bytes_text);
            tma_bulk_g2s(s_blob, p.blobs + static_cast<size_t>(g) * K1_WARPS * BW, bytes_blob, &mbar);
            tma_bulk_g2s(s_text, p.text + c0, bytes_text, &mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1u;
        k1_warp_run<U, TRACK_END>(s_blob + warp * BW, s_text, nch, p.tile_text0[tile], p, &lut);
        __syncthreads();  // everyone is done with this item's shared memory
    }
}

// ------------------------------------------------------------------------------------------
// K3 span recovery: start column of the optimal placement that ends at the column K1 reported.
// For every (text t, pattern p): reversed P against reversed T[0 .. end), anchored at the window start
// (SP_PREFIX boundary), smallest end column c of a best placement => start = end - c, the rightmost
// start among optimal placements ending at `end`.  Replaces the m.query_start / m.target_start reads of
// src/cyp2d6/chaining.rs:69-81 and src/cyp2d6/haplotyper.rs:203-249.  One warp per pair, one pattern
// per bin (lane width SPAN_U), text bytes read straight from global memory (the pair count is small).
// ------------------------------------------------------------------------------------------
constexpr int SPAN_U = 16;       // patterns up to 16,384 rows
constexpr int SPAN_U_LONG = 24;  // up to 24,576

struct SpanParams {
    const uint32_t *blobs;   // [np] bins, reversed rows, prefix pad rows
    const uint8_t *tbases;   // ASCII texts
    const long long *toffs;
    const int32_t *D;        // forward distances  [p * ld + t]
    const int32_t *E;        // forward end columns [p * ld + t]
    int32_t *S;              // out: start columns  [p * ld + t]
    long long ld;
    int nt, np;
    uint32_t one, m1;
    int max_dist_permille;   // pairs with D * 1000 > |P| * this get S = -1 and no reverse pass; < 0: every pair
    const int32_t *plen;     // [np] pattern lengths
    unsigned long long *next_pair;  // work counter (zeroed by the host): pairs are handed out one at a time
};

template <int U>
__global__ void __launch_bounds__(K1_THREADS, 1) k3_span_starts(const SpanParams p) {
    constexpr int BW = blob_words(U);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *blob = reinterpret_cast<uint32_t *>(smem_raw) + warp * BW;
    const long long total = static_cast<long long>(p.nt) * p.np;
    int cur_pat = -1;
    for (;;) {  // dynamic hand-out, pattern-major: the pass lengths differ by orders of magnitude between pairs
        unsigned long long qq = 0;
        if (lane == 0) qq = atomicAdd(p.next_pair, 1ull);
        const long long q = static_cast<long long>(__shfl_sync(0xffffffffu, qq, 0));
        if (q >= total) break;
        const int pat = static_cast<int>(q / p.nt), t = static_cast<int>(q - static_cast<long long>(pat) * p.nt);
        {
            const long long o0 = static_cast<long long>(pat) * p.ld + t;
            const long long mlen = p.plen[pat];
            if (p.max_dist_permille >= 0 && static_cast<long long>(p.D[o0]) * 1000 > mlen * p.max_dist_permille) {
                if (lane == 0) p.S[o0] = -1;  // too far apart for an aligner to have reported anything: no span
                continue;
            }
        }
        if (pat != cur_pat) {
            __syncwarp();
            const uint32_t *src = p.blobs + static_cast<size_t>(pat) * BW;
            for (int i = lane; i < BW; i += 32) blob[i] = src[i];
            __syncwarp();
            cur_pat = pat;
        }
        const uint32_t info1 = blob[PEQ_ROWS * 32 * U + 32 + lane];
        const bool first = (info1 & INFO_FIRST) != 0, last = (info1 & INFO_LAST) != 0;
        const int m = static_cast<int>(info1 & INFO_LEN_MASK);
        const long long o = static_cast<long long>(pat) * p.ld + t;
        const int e = p.E[o], d = p.D[o];
        const int m_all = __shfl_sync(0xffffffffu, m, 0);  // lane 0 always belongs to the pattern (or it is empty)
        const bool empty = blob[PEQ_ROWS * 32 * U] == NO_PATTERN;
        if (empty) { if (lane == 0) p.S[o] = e; continue; }
        const int wlen = min(e, m_all + d);
        const int nch = (wlen + K1_CHUNK - 1) / K1_CHUNK;
        const uint8_t *T = p.tbases + p.toffs[t];
        uint32_t npv[U], mv[U];
        load_row<U>(blob + 5 * (32 * U), lane, npv);
#pragma unroll
        for (int u = 0; u < U; ++u) { npv[u] = ~npv[u]; mv[u] = 0; }
        int score = m, best = m, col = 0, best_col = 0;
        uint32_t carry_out = 0;
        const int nsteps = nch + 31;
        for (int s = 0; s < nsteps; ++s) {
            uint32_t cin = __shfl_up_sync(0xffffffffu, carry_out, 1);
            if (first) cin = 0x00FFu;  // anchored: the row above the pattern costs one per column
            const int idx = s - lane;
            if (static_cast<unsigned>(idx) < static_cast<unsigned>(nch)) {
                uint32_t X = cin << 24, Y = cin << 16;
                uint32_t cph = 0, cmh = 0;
#pragma unroll 1
                for (int c = 0; c < K1_CHUNK; ++c) {
                    const int j = idx * K1_CHUNK + c;
                    const uint32_t code = j < wlen ? base_code(T[e - 1 - j]) : 4u;
                    column_step<U, true>(blob, lane, code, p.one, p.m1, npv, mv, X, Y, cph, cmh, score, best, col, best_col);
                }
                carry_out = cph | (cmh << 8);
            }
        }
        if (last) p.S[o] = e - best_col;
        if (nch == 0 && lane == 0) p.S[o] = e;  // empty window: the placement is empty, start == end
    }
}

// ------------------------------------------------------------------------------------------
// K3 chain windows: B[c][r] = min over windows s of chain c of sum_t W[r][t][chain_c[s + t]], or 2 * worst_r
// when the chain is shorter than the read's segment count -- the inner loop of containment_score
// (src/cyp2d6/chaining.rs:683-731) hoisted out of the chain-pair loop: for a pair (i, j) the reference's
// best_score is min(B[i][r], B[j][r]), so K2 on B gives the ED term of every pair (chaining.rs:470-485).
// One thread per (read, chain); reads contiguous in the output (the K2 layout).
// ------------------------------------------------------------------------------------------
struct ChainWinParams {
    const int32_t *chain_off;    // [n_chains + 1]
    const int32_t *chain_items;  // consensus (haplotype) indices
    const int32_t *seg_off;      // [n_reads + 1] first segment row of each read in W
    const uint32_t *W;           // [n_segments][n_haps] edit distances
    int32_t *B;                  // [n_chains][ld]
    long long ld;
    int n_chains, n_reads, n_haps;
};

#ifndef SP_NO_GLOBAL_KERNELS
__global__ void __launch_bounds__(256) k3_chain_windows(const ChainWinParams p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads) return;
    const int s0 = p.seg_off[r], w = p.seg_off[r + 1] - s0;
    for (int c = blockIdx.y; c < p.n_chains; c += gridDim.y) {  // grid.y is capped at 65,535: stride over the chains
    const int c0 = p.chain_off[c], len = p.chain_off[c + 1] - c0;
    uint32_t best;
    if (len < w) {
        unsigned long long worst = 0;  // 2 * sum_t max_k W[r][t][k]
        for (int t = 0; t < w; ++t) {
            uint32_t mx = 0;
            for (int k = 0; k < p.n_haps; ++k) mx = max(mx, p.W[static_cast<long long>(s0 + t) * p.n_haps + k]);
            worst += mx;
        }
        best = static_cast<uint32_t>(min(2ull * worst, 0x7FFFFFFFull));
    } else {
        best = 0xFFFFFFFFu;
        for (int s = 0; s + w <= len; ++s) {
            uint32_t tot = 0;
            for (int t = 0; t < w; ++t)
                tot += p.W[static_cast<long long>(s0 + t) * p.n_haps + p.chain_items[c0 + s + t]];
            best = min(best, tot);
        }
        best = min(best, 0x7FFFFFFFu);
    }
    p.B[static_cast<long long>(c) * p.ld + r] = static_cast<int32_t>(best);
    }
}

// ------------------------------------------------------------------------------------------
// Pack kernels
// ------------------------------------------------------------------------------------------

// one thread per (bin, lane, word): builds the six 32-row masks of that word
__global__ void pack_patterns(const uint8_t *__restrict__ bases, const long long *__restrict__ offs,
                              const int32_t *__restrict__ lane_pat, const int32_t *__restrict__ lane_row0,
                              const uint32_t *__restrict__ lane_info1, uint32_t *__restrict__ blobs, int n_bins,
                              int U, int prefix_mode, int reverse) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(n_bins) * 32 * U;
    if (idx >= total) return;
    const int bin = static_cast<int>(idx / (32 * U));
    const int rem = static_cast<int>(idx - static_cast<long long>(bin) * 32 * U);
    const int lane = rem / U, u = rem % U;
    const int BW = blob_words(U);
    const int pat = lane_pat[bin * 32 + lane];
    uint32_t mask[PEQ_ROWS] = {0, 0, 0, 0, 0, 0};
    if (pat >= 0) {
        const uint8_t *P = bases + offs[pat];
        const int m = static_cast<int>(offs[pat + 1] - offs[pat]);
        const int row0 = lane_row0[bin * 32 + lane] + 32 * u;
        for (int b = 0; b < 32; ++b) {
            const int r = row0 + b;
            const uint32_t bit = 1u << b;
            if (r < 0) {  // pad row above the pattern: wildcard (infix) / pass-through (prefix)
                if (!prefix_mode) { mask[0] |= bit; mask[1] |= bit; mask[2] |= bit; mask[3] |= bit; mask[4] |= bit; }
            } else {
                const uint32_t c = base_code(P[reverse ? m - 1 - r : r]);
                if (c < 4) mask[c] |= bit;
                mask[5] |= bit;
            }
        }
    }
    uint32_t *blob = blobs + static_cast<size_t>(bin) * BW;
#pragma unroll
    for (int k = 0; k < PEQ_ROWS; ++k) blob[k * 32 * U + row_word(U, lane, u)] = mask[k];
    if (u == 0) {
        blob[PEQ_ROWS * 32 * U + lane] = pat >= 0 ? static_cast<uint32_t>(pat) : NO_PATTERN;
        blob[PEQ_ROWS * 32 * U + 32 + lane] = pat >= 0 ? lane_info1[bin * 32 + lane] : INFO_FIRST;
    }
}

// one CTA per text; 8 byte-codes per chunk; bit 7 of the last byte of the last chunk ends the text
__global__ void pack_texts(const uint8_t *__restrict__ bases, const long long *__restrict__ offs,
                           const int32_t *__restrict__ text_chunk0, const int32_t *__restrict__ text_nch,
                           uint2 *__restrict__ out, int n_texts) {
    for (int t = blockIdx.x; t < n_texts; t += gridDim.x) {
        const uint8_t *T = bases + offs[t];
        const long long n = offs[t + 1] - offs[t];
        const int nc = text_nch[t];
        uint2 *dst = out + text_chunk0[t];
        for (int k = threadIdx.x; k < nc; k += blockDim.x) {
            uint32_t w[2] = {0, 0};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const long long colx = static_cast<long long>(k) * 8 + c;
                uint32_t code = colx < n ? base_code(T[colx]) : 4u;
                w[c >> 2] |= code << (8 * (c & 3));
            }
            if (k == nc - 1) w[1] |= 0x80000000u;
            dst[k] = make_uint2(w[0], w[1]);
        }
    }
}

#endif  // SP_NO_GLOBAL_KERNELS

// D[p*ld + t] (u16 / i32) -> host-order rows[t * np + p] (int32, or u16 for the 16-bit read-back)
template <typename T, typename O>
__global__ void dmatrix_to_rows(const T *__restrict__ D, long long ld, int nt, int np, O *__restrict__ rows) {
    __shared__ O tile[32][33];
    const int t0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < np && t < nt) ? static_cast<O>(D[static_cast<long long>(p) * ld + t]) : O(0);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, p = p0 + threadIdx.x;
        if (t < nt && p < np) rows[static_cast<long long>(t) * np + p] = tile[threadIdx.x][i];
    }
}

// host-order int32 rows[r * A + a] -> D[a*ld + r] int32 (for the *_host K2 entry points)
#ifndef SP_NO_GLOBAL_KERNELS
__global__ void rows_to_dmatrix(const int32_t *__restrict__ rows, int R, int A, long long ld, int32_t *__restrict__ D) {
    __shared__ int32_t tile[32][33];
    const int a0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, a = a0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && a < A) ? rows[static_cast<long long>(r) * A + a] : 0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int a = a0 + i, r = r0 + threadIdx.x;
        if (a < A && r < R) D[static_cast<long long>(a) * ld + r] = tile[threadIdx.x][i];
    }
}

#endif  // SP_NO_GLOBAL_KERNELS

// ------------------------------------------------------------------------------------------
// K2: pair min-sum.  CTA = 64 x 64 pair tile (i-tile I, j-tile J >= I), 256 threads x (4 x 4).
// Optional second matrix: key = (S1, S2, i, j), the (cDNA, DNA) lexicographic order of
// HlaMappingScore (src/hla/mapping.rs:111-117) carried over to pair sums.
// ------------------------------------------------------------------------------------------
constexpr int K2_TILE = 64;
constexpr int K2_RC = 32;        // reads per shared-memory stage
constexpr int K2_THREADS = 256;
constexpr int K2_MAXK = 64;

struct PairKey {
    unsigned long long score;
    unsigned long long score2;
    unsigned long long ij;  // i << 32 | j
};
__device__ __forceinline__ bool key_less(const PairKey &a, const PairKey &b) {
    if (a.score != b.score) return a.score < b.score;
    if (a.score2 != b.score2) return a.score2 < b.score2;
    return a.ij < b.ij;
}

struct K2Params {
    const void *D;        // [A][ld], u16 or i32 (primary)
    const void *D2;       // secondary matrix, same geometry, or nullptr
    long long ld, ld2;
    int R, A;
    int i_begin, i_end;   // row range owned by this call
    int tile_i0;          // first i-tile index (i_begin / 64)
    int n_tiles_j;        // ceil(A / 64)
    int k;                // top-k per CTA (0 in full mode)
    PairKey *cand;        // [n_ctas][k]   (top-k mode)
    unsigned long long *S; // [A][A]       (full mode)
};

// linear CTA index -> (I, J) with J >= I over rows I in [tile_i0, tile_i1)
__device__ __forceinline__ void k2_decode_tile(int bid, int tile_i0, int n_tiles_j, int &I, int &J) {
    int I_ = tile_i0;
    int rem = bid;
    while (rem >= n_tiles_j - I_) { rem -= n_tiles_j - I_; ++I_; }
    I = I_;
    J = I_ + rem;
}

template <typename T> struct K2Acc { using type = unsigned long long; };
template <> struct K2Acc<uint16_t> { using type = uint32_t; };  // host checks R * 65535 < 2^32

template <typename T, bool DUAL>
__device__ __forceinline__ void k2_stage(const K2Params &p, int I, int J, int r0, int32_t (*sA)[K2_TILE + 4],
                                         int32_t (*sB)[K2_TILE + 4], int32_t (*sA2)[K2_TILE + 4],
                                         int32_t (*sB2)[K2_TILE + 4]) {
    const T *D = reinterpret_cast<const T *>(p.D);
    const T *D2 = reinterpret_cast<const T *>(p.D2);
    for (int e = threadIdx.x; e < K2_TILE * K2_RC; e += K2_THREADS) {
        const int a = e / K2_RC, r = e % K2_RC;
        const int ga = I * K2_TILE + a, gb = J * K2_TILE + a, gr = r0 + r;
        int32_t va = 0, vb = 0, va2 = 0, vb2 = 0;
        if (gr < p.R) {
            if (ga < p.A) {
                va = static_cast<int32_t>(D[static_cast<long long>(ga) * p.ld + gr]);
                if (DUAL) va2 = static_cast<int32_t>(D2[static_cast<long long>(ga) * p.ld2 + gr]);
            }
            if (gb < p.A) {
                vb = static_cast<int32_t>(D[static_cast<long long>(gb) * p.ld + gr]);
                if (DUAL) vb2 = static_cast<int32_t>(D2[static_cast<long long>(gb) * p.ld2 + gr]);
            }
        }
        sA[r][a] = va;
        sB[r][a] = vb;
        if (DUAL) { sA2[r][a] = va2; sB2[r][a] = vb2; }
    }
}

template <typename T, bool FULL, bool DUAL>
__global__ void __launch_bounds__(K2_THREADS) k2_pair_minsum(const K2Params p) {
    using Acc = typename K2Acc<T>::type;
    __shared__ __align__(16) int32_t sA[K2_RC][K2_TILE + 4];
    __shared__ __align__(16) int32_t sB[K2_RC][K2_TILE + 4];
    __shared__ __align__(16) int32_t sA2[DUAL ? K2_RC : 1][K2_TILE + 4];
    __shared__ __align__(16) int32_t sB2[DUAL ? K2_RC : 1][K2_TILE + 4];
    __shared__ PairKey s_red[K2_THREADS / 32];
    __shared__ PairKey s_last;

    int I, J;
    k2_decode_tile(blockIdx.x, p.tile_i0, p.n_tiles_j, I, J);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    Acc acc[4][4], acc2[DUAL ? 4 : 1][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            acc[a][b] = 0;
            if (DUAL) acc2[a][b] = 0;
        }

    for (int r0 = 0; r0 < p.R; r0 += K2_RC) {
        // stage: D[(I*64 + a) * ld + r0 + r] -> sA[r][a]   (reads contiguous in global memory)
        k2_stage<T, DUAL>(p, I, J, r0, sA, sB, sA2, sB2);
        __syncthreads();
        uint32_t part[4][4], part2[DUAL ? 4 : 1][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                part[a][b] = 0;
                if (DUAL) part2[a][b] = 0;
            }
#pragma unroll 8
        for (int r = 0; r < K2_RC; ++r) {
            const int4 va = *reinterpret_cast<const int4 *>(&sA[r][ty * 4]);
            const int4 vb = *reinterpret_cast<const int4 *>(&sB[r][tx * 4]);
            const int xa[4] = {va.x, va.y, va.z, va.w};
            const int xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) part[a][b] += static_cast<uint32_t>(min(xa[a], xb[b]));
            if (DUAL) {
                const int4 wa = *reinterpret_cast<const int4 *>(&sA2[r][ty * 4]);
                const int4 wb = *reinterpret_cast<const int4 *>(&sB2[r][tx * 4]);
                const int ya[4] = {wa.x, wa.y, wa.z, wa.w};
                const int yb[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) part2[a][b] += static_cast<uint32_t>(min(ya[a], yb[b]));
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                acc[a][b] += part[a][b];
                if (DUAL) acc2[a][b] += part2[a][b];
            }
        __syncthreads();
    }

    // validity: i <= j, inside the matrix, i inside the owned row range
    PairKey keys[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gi = I * K2_TILE + ty * 4 + a, gj = J * K2_TILE + tx * 4 + b;
            const bool ok = gi < p.A && gj < p.A && gi <= gj && gi >= p.i_begin && gi < p.i_end;
            if (FULL) {
                if (ok) p.S[static_cast<long long>(gi) * p.A + gj] = acc[a][b];
            } else {
                keys[a * 4 + b].score = ok ? static_cast<unsigned long long>(acc[a][b]) : ~0ull;
                keys[a * 4 + b].score2 = ok ? (DUAL ? static_cast<unsigned long long>(acc2[DUAL ? a : 0][b]) : 0ull) : ~0ull;
                keys[a * 4 + b].ij = ok ? ((static_cast<unsigned long long>(gi) << 32) | static_cast<unsigned>(gj)) : ~0ull;
            }
        }
    if (FULL) return;

    // k rounds of block-wide lexicographic argmin over keys strictly greater than the last one taken
    PairKey lastk;
    lastk.score = 0; lastk.score2 = 0; lastk.ij = 0;
    bool have_last = false;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int round = 0; round < p.k; ++round) {
        PairKey best;
        best.score = ~0ull; best.score2 = ~0ull; best.ij = ~0ull;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const bool gt = !have_last || key_less(lastk, keys[q]);
            if (gt && key_less(keys[q], best)) best = keys[q];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            PairKey o;
            o.score = __shfl_xor_sync(0xffffffffu, best.score, off);
            o.score2 = __shfl_xor_sync(0xffffffffu, best.score2, off);
            o.ij = __shfl_xor_sync(0xffffffffu, best.ij, off);
            if (key_less(o, best)) best = o;
        }
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            PairKey b = s_red[0];
            for (int w = 1; w < K2_THREADS / 32; ++w)
                if (key_less(s_red[w], b)) b = s_red[w];
            s_last = b;
            p.cand[static_cast<long long>(blockIdx.x) * p.k + round] = b;
        }
        __syncthreads();
        lastk = s_last;
        have_last = true;
    }
}

// c1[q] = #{r : (D[i_q][r], D2[i_q][r]) <= (D[j_q][r], D2[j_q][r])} for the final records; one CTA per record
template <typename T>
__global__ void k2_count_c1(const T *__restrict__ D, long long ld, const T *__restrict__ D2, long long ld2, int R,
                            const uint32_t *__restrict__ ij, uint32_t *__restrict__ c1) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint32_t i = ij[2 * blockIdx.x], j = ij[2 * blockIdx.x + 1];
    uint32_t cnt = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const T x = D[static_cast<long long>(i) * ld + r], y = D[static_cast<long long>(j) * ld + r];
        bool le = x <= y;
        if (D2 != nullptr && x == y) le = D2[static_cast<long long>(i) * ld2 + r] <= D2[static_cast<long long>(j) * ld2 + r];
        cnt += le;
    }
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) c1[blockIdx.x] = s_cnt;
}

// ------------------------------------------------------------------------------------------
// Integer pipe microbenchmark (roofline denominator)
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters) {
    uint32_t a[8], b = threadIdx.x * 2654435761u + 1u, c = blockIdx.x + 0x9E3779B9u;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i * 7919u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else if (KIND == 1) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else if (KIND == 2) {
                    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else if (KIND == 3) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                } else if (KIND == 4) {
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else {
                    if (i & 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

}  // namespace sp
