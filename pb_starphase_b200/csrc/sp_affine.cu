// sp_affine.cu -- K9 host side: sp_align_affine_resident (include/starphase_gpu.h).
#include "sp_internal.cuh"

#define SP_NO_GLOBAL_KERNELS
#include "sp_kernels.cuh"
#include "sp_affine.cuh"

using namespace sp;

namespace {
template <int CELLS>
int k9_occupancy() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k9_affine_local<CELLS>, 128, 0) != cudaSuccess) occ = 1;
    return std::max(occ, 1);
}
}  // namespace

extern "C" sp_status sp_align_affine_resident(sp_ctx *ctx, const sp_targets *texts, const sp_targets *patterns, int64_t n_pairs,
                                              const int32_t *pair_text, const int32_t *pair_pattern, const int32_t *win_begin,
                                              const int32_t *win_end, const int32_t *band_centre, int32_t band, const sp_affine_costs *costs,
                                              sp_align_rec *recs, int32_t *scores, uint32_t *cigar, int64_t cigar_cap, int64_t *cigar_used) {
    if (!ctx) return SP_ERR_INVALID;
    if (!texts || !patterns || !costs) return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: NULL argument");
    if ((win_begin == nullptr) != (win_end == nullptr)) return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: win_begin and win_end go together");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_text || !pair_pattern || !recs || !scores)) || cigar_cap < 0 || (cigar_cap > 0 && !cigar))
        return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: bad argument");
    if (band < 1 || band > 255) return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: band must be in [1, 255]");
    if (costs->a < 1 || costs->b < 0 || costs->q < 0 || costs->e < 1 || costs->q2 < 0 || costs->e2 < 1 || costs->a > 100 || costs->b > 1000 ||
        costs->q > 10000 || costs->q2 > 10000 || costs->e > 1000 || costs->e2 > 1000)
        return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: costs out of range");
    if (cigar_used) *cigar_used = 0;
    if (n_pairs == 0) return SP_OK;
    if (n_pairs > 0x7FFFFFF0ll) return fail(ctx, SP_ERR_RANGE, "sp_align_affine_resident: too many pairs");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nb = 2 * band + 1;
    std::vector<AffinePairDev> pairs(static_cast<size_t>(n_pairs));
    int64_t cig_total = 0, max_trace = 16;
    for (int64_t q = 0; q < n_pairs; ++q) {
        const int64_t t = pair_text[q], pi = pair_pattern[q];
        if (t < 0 || t >= texts->n || pi < 0 || pi >= patterns->n) return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: pair index outside the sequence sets");
        const int64_t m = patterns->h_offs[static_cast<size_t>(pi) + 1] - patterns->h_offs[static_cast<size_t>(pi)];
        int64_t n = texts->h_offs[static_cast<size_t>(t) + 1] - texts->h_offs[static_cast<size_t>(t)], t_off = texts->h_offs[static_cast<size_t>(t)];
        if (win_begin) {
            if (win_begin[q] < 0 || win_end[q] < win_begin[q] || win_end[q] > n) return fail(ctx, SP_ERR_INVALID, "sp_align_affine_resident: window outside its text");
            t_off += win_begin[q];
            n = win_end[q] - win_begin[q];
        }
        if (m > 0x3FFFFFFF || n > 0x3FFFFFFF) return fail(ctx, SP_ERR_TOO_LONG, "sp_align_affine_resident: sequence too long");
        AffinePairDev &d = pairs[static_cast<size_t>(q)];
        d.t_off = t_off; d.p_off = patterns->h_offs[static_cast<size_t>(pi)];
        d.n = static_cast<int32_t>(n); d.m = static_cast<int32_t>(m);
        d.centre = band_centre ? band_centre[q] : 0;
        d.trace_off = 0;
        d.cig_off = cig_total;
        d.cig_len = static_cast<int32_t>(std::min<int64_t>(m + std::min<int64_t>(n, m + 2ll * band) + 2, 0x7FFFFFF0ll));
        d.out = static_cast<int32_t>(q); d.pad_ = 0;
        cig_total += d.cig_len;
        max_trace = std::max<int64_t>(max_trace, (m * nb + 15) / 16 * 16);
    }
    const int cells = (nb + 31) / 32;
    const int occ = cells <= 3 ? k9_occupancy<3>() : cells <= 5 ? k9_occupancy<5>() : cells <= 9 ? k9_occupancy<9>() : k9_occupancy<16>();
    int64_t n_slots = std::min<int64_t>(n_pairs, static_cast<int64_t>(ctx->num_sms) * occ * 4);
    const int64_t budget = 8ll << 30;
    n_slots = std::max<int64_t>(1, std::min(n_slots, budget / max_trace));
    const int grid = static_cast<int>((n_slots + 3) / 4);
    uint8_t *d_trace = nullptr;
    uint32_t *d_cigar = nullptr, *d_dense = nullptr;
    AffinePairDev *d_pairs = nullptr;
    AlignRecDev *d_recs = nullptr;
    int32_t *d_scores = nullptr;
    unsigned long long *d_used = nullptr;
    auto cleanup = [&]() { dev_free(ctx, d_pairs); dev_free(ctx, d_recs); dev_free(ctx, d_scores); dev_free(ctx, d_used); };
    auto cu = [&](cudaError_t e, const char *what) -> sp_status {
        if (e != cudaSuccess)
            return fail(ctx, e == cudaErrorMemoryAllocation ? SP_ERR_NOMEM : SP_ERR_CUDA, std::string("sp_align_affine_resident: ") + what + ": " + cudaGetErrorString(e));
        return SP_OK;
    };
#define SP_TRY(x)                                   \
    do {                                            \
        sp_status s__ = (x);                        \
        if (s__ != SP_OK) { cleanup(); return s__; } \
    } while (0)
    SP_TRY(cu(ctx_scratch(ctx, static_cast<size_t>(grid) * 4 * static_cast<size_t>(max_trace), reinterpret_cast<void **>(&d_trace)), "trace scratch"));
    SP_TRY(cu(ctx_pool(ctx, 1, static_cast<size_t>(cig_total) * 4, reinterpret_cast<void **>(&d_cigar)), "cigar pool"));
    SP_TRY(cu(ctx_pool(ctx, 3, static_cast<size_t>(std::max<int64_t>(cigar_cap, 4)) * 4, reinterpret_cast<void **>(&d_dense)), "dense cigar pool"));
    SP_TRY(cu(dev_malloc(ctx, &d_pairs, pairs.size() * sizeof(AffinePairDev)), "cudaMalloc"));
    SP_TRY(cu(dev_malloc(ctx, &d_recs, static_cast<size_t>(n_pairs) * sizeof(AlignRecDev)), "cudaMalloc"));
    SP_TRY(cu(dev_malloc(ctx, &d_scores, static_cast<size_t>(n_pairs) * 4), "cudaMalloc"));
    SP_TRY(cu(dev_malloc(ctx, &d_used, sizeof(unsigned long long)), "cudaMalloc"));
    SP_TRY(cu(cudaMemsetAsync(d_used, 0, sizeof(unsigned long long), ctx->stream), "memset"));
    SP_TRY(cu(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream), "memset"));
    SP_TRY(cu(cudaMemcpyAsync(d_pairs, pairs.data(), pairs.size() * sizeof(AffinePairDev), cudaMemcpyHostToDevice, ctx->stream), "H2D"));
    AffineParams prm;
    prm.tbases = texts->d_bases; prm.pbases = patterns->d_bases; prm.pairs = d_pairs; prm.trace = d_trace; prm.slot_bytes = max_trace;
    prm.cigar = d_cigar; prm.dense = d_dense; prm.dense_used = d_used; prm.dense_cap = static_cast<unsigned long long>(cigar_cap);
    prm.recs = d_recs; prm.scores = d_scores; prm.n_pairs = static_cast<int>(n_pairs); prm.W = band;
    prm.a = costs->a; prm.b = costs->b; prm.q = costs->q; prm.e = costs->e; prm.q2 = costs->q2; prm.e2 = costs->e2;
    prm.next_pair = ctx->d_counter;
    ev_begin(ctx, 4);
    if (cells <= 3) k9_affine_local<3><<<grid, 128, 0, ctx->stream>>>(prm);
    else if (cells <= 5) k9_affine_local<5><<<grid, 128, 0, ctx->stream>>>(prm);
    else if (cells <= 9) k9_affine_local<9><<<grid, 128, 0, ctx->stream>>>(prm);
    else k9_affine_local<16><<<grid, 128, 0, ctx->stream>>>(prm);
    ev_end(ctx, 4);
    ++ctx->launches;
    SP_TRY(cu(cudaGetLastError(), "k9_affine_local launch"));
    std::vector<AlignRecDev> hrec(static_cast<size_t>(n_pairs));
    unsigned long long used = 0;
    SP_TRY(cu(cudaMemcpyAsync(hrec.data(), d_recs, hrec.size() * sizeof(AlignRecDev), cudaMemcpyDeviceToHost, ctx->stream), "D2H recs"));
    SP_TRY(cu(cudaMemcpyAsync(scores, d_scores, static_cast<size_t>(n_pairs) * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H scores"));
    SP_TRY(cu(cudaMemcpyAsync(&used, d_used, sizeof(used), cudaMemcpyDeviceToHost, ctx->stream), "D2H"));
    SP_TRY(cu(cudaStreamSynchronize(ctx->stream), "k9_affine_local"));
    if (cigar_used) *cigar_used = static_cast<int64_t>(used);
    if (used > static_cast<unsigned long long>(cigar_cap)) {
        cleanup();
        return fail(ctx, SP_ERR_RANGE, "sp_align_affine_resident: cigar buffer too small: " + std::to_string(used) + " entries needed");
    }
    if (used > 0) {
        SP_TRY(cu(cudaMemcpyAsync(cigar, d_dense, static_cast<size_t>(used) * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H cigar"));
        SP_TRY(cu(cudaStreamSynchronize(ctx->stream), "D2H cigar"));
    }
    for (int64_t q = 0; q < n_pairs; ++q) {
        const AlignRecDev &r = hrec[static_cast<size_t>(q)];
        sp_align_rec &o = recs[q];
        o.dist = r.dist; o.nm = r.nm; o.p_start = r.p_start; o.p_end = r.p_end; o.t_start = r.t_start; o.t_end = r.t_end;
        o.n_cigar = r.n_cigar; o._pad = 0; o.cigar_off = r.cigar_off;
    }
#undef SP_TRY
    cleanup();
    return SP_OK;
}
