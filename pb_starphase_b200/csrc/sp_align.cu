// sp_align.cu -- K4 host side: sp_align_pairs / sp_align_windows (include/starphase_gpu.h).
#include "sp_internal.cuh"

#define SP_NO_GLOBAL_KERNELS  // the non-template kernels of sp_kernels.cuh are instantiated in starphase_gpu.cu
#include "sp_kernels.cuh"
#include "sp_align.cuh"

using namespace sp;

// ------------------------------------------------------------------------------------------
// K4: traceback alignment of selected pairs (what the host reads from a minimap2::Mapping)
// ------------------------------------------------------------------------------------------
static_assert(sizeof(sp_align_rec) == sizeof(AlignRecDev), "sp_align_rec layout");

extern "C" sp_status sp_align_pairs(sp_ctx *ctx, const sp_seqset *targets, const sp_seqset *patterns, int64_t n_pairs,
                                    const int32_t *pair_target, const int32_t *pair_pattern, sp_align_rec *recs,
                                    uint32_t *cigar, int64_t cigar_cap, int64_t *cigar_used) {
    return sp_align_windows(ctx, targets, patterns, n_pairs, pair_target, pair_pattern, nullptr, nullptr, recs, cigar, cigar_cap, cigar_used);
}

extern "C" sp_status sp_align_windows(sp_ctx *ctx, const sp_seqset *targets, const sp_seqset *patterns, int64_t n_pairs,
                                      const int32_t *pair_target, const int32_t *pair_pattern, const int32_t *win_begin,
                                      const int32_t *win_end, sp_align_rec *recs, uint32_t *cigar, int64_t cigar_cap,
                                      int64_t *cigar_used) {
    if (!ctx) return SP_ERR_INVALID;
    if ((win_begin == nullptr) != (win_end == nullptr)) return fail(ctx, SP_ERR_INVALID, "sp_align_windows: win_begin and win_end go together");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_target || !pair_pattern || !recs)) || cigar_cap < 0 ||
        (cigar_cap > 0 && !cigar))
        return fail(ctx, SP_ERR_INVALID, "sp_align_pairs: bad argument");
    if (cigar_used) *cigar_used = 0;
    sp_status st = check_seqset(ctx, targets, "targets");
    if (st == SP_OK) st = check_seqset(ctx, patterns, "patterns");
    if (st != SP_OK) return st;
    if (n_pairs == 0) return SP_OK;
    if (n_pairs > 0x7FFFFFF0ll) return fail(ctx, SP_ERR_RANGE, "sp_align_pairs: too many pairs");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    PhaseTimer tm;

    // only the sequences some pair names travel to the device
    std::vector<int32_t> t_local(static_cast<size_t>(targets->n), -1), p_local(static_cast<size_t>(patterns->n), -1);
    std::vector<int64_t> t_ids, p_ids;
    std::vector<int32_t> pt(static_cast<size_t>(n_pairs)), pp(static_cast<size_t>(n_pairs));
    for (int64_t q = 0; q < n_pairs; ++q) {
        const int64_t t = pair_target[q], pi = pair_pattern[q];
        if (t < 0 || t >= targets->n || pi < 0 || pi >= patterns->n)
            return fail(ctx, SP_ERR_INVALID, "sp_align_pairs: pair index outside the sequence sets");
        if (t_local[static_cast<size_t>(t)] < 0) { t_local[static_cast<size_t>(t)] = static_cast<int32_t>(t_ids.size()); t_ids.push_back(t); }
        if (p_local[static_cast<size_t>(pi)] < 0) { p_local[static_cast<size_t>(pi)] = static_cast<int32_t>(p_ids.size()); p_ids.push_back(pi); }
        pt[static_cast<size_t>(q)] = t_local[static_cast<size_t>(t)];
        pp[static_cast<size_t>(q)] = p_local[static_cast<size_t>(pi)];
    }
    auto gather = [](const sp_seqset *s, const std::vector<int64_t> &ids, std::vector<uint8_t> &bases, std::vector<int64_t> &offs) {
        offs.assign(1, 0);
        for (int64_t id : ids) {
            bases.insert(bases.end(), s->bases + s->offsets[id], s->bases + s->offsets[id + 1]);
            offs.push_back(static_cast<int64_t>(bases.size()));
        }
        if (bases.empty()) bases.push_back(0);
    };
    std::vector<uint8_t> tb, pbs;
    std::vector<int64_t> to, po;
    gather(targets, t_ids, tb, to);
    gather(patterns, p_ids, pbs, po);
    const int64_t np = static_cast<int64_t>(p_ids.size());
    const int64_t rows = 32ll * ALN_U;
    int64_t max_slot_words = 4;
    std::vector<long long> cig_off(static_cast<size_t>(n_pairs) + 1, 0);
    for (int64_t i = 0; i < np; ++i)
        if (po[static_cast<size_t>(i) + 1] - po[static_cast<size_t>(i)] > SP_MAX_PATTERN_LEN)
            return fail(ctx, SP_ERR_TOO_LONG, "sp_align_pairs: pattern exceeds SP_MAX_PATTERN_LEN");
    for (int64_t q = 0; q < n_pairs; ++q) {
        const int64_t m = po[static_cast<size_t>(pp[static_cast<size_t>(q)]) + 1] - po[static_cast<size_t>(pp[static_cast<size_t>(q)])];
        int64_t n = to[static_cast<size_t>(pt[static_cast<size_t>(q)]) + 1] - to[static_cast<size_t>(pt[static_cast<size_t>(q)])];
        if (n > 0x7FFFFF00ll) return fail(ctx, SP_ERR_TOO_LONG, "sp_align_pairs: text too long");
        if (win_begin) {
            if (win_begin[q] < 0 || win_end[q] < win_begin[q] || win_end[q] > n)
                return fail(ctx, SP_ERR_INVALID, "sp_align_windows: window outside its text");
            n = win_end[q] - win_begin[q];
        }
        const int64_t ncols = std::min(n, 2 * m);  // window = m + d columns, d <= m
        const int64_t nl = (m + rows - 1) / rows, pad = nl * rows - m;
        const int64_t Wp = nl * ALN_U - ((pad >> 5) & ~3ll);
        max_slot_words = std::max(max_slot_words, ncols * 2 * Wp);
        cig_off[static_cast<size_t>(q) + 1] = cig_off[static_cast<size_t>(q)] + m + ncols + 1;
    }
    max_slot_words = (max_slot_words + 3) / 4 * 4;
    tm.mark("gather + plan");
    // one scratch slot per warp in flight (the context's grow-only scratch: no malloc / free per call), capped at a quarter of
    // the free HBM and 16 GB; the slots are spread over all SMs, 1..K1_WARPS warps per CTA
    int64_t n_slots = std::min<int64_t>(n_pairs, 2ll * ctx->num_sms * K1_WARPS);
    size_t free_b = 0, total_b = 0;
    SP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const int64_t budget_bytes = std::max<int64_t>(4ll << 30, std::min<int64_t>(16ll << 30, static_cast<int64_t>((free_b + ctx->pool_bytes[0]) / 4)));
    const int64_t budget_words = budget_bytes / 4;
    n_slots = std::max<int64_t>(1, std::min(n_slots, budget_words / max_slot_words));
    // hysteresis: the budget follows the free memory, which moves with the stream-ordered pool; re-growing a multi-GB scratch
    // costs ~100 ms (cudaFree + cudaMalloc), so a pool that already holds at least half of the wanted slots is used as it is
    {
        const int64_t have_slots = static_cast<int64_t>(ctx->pool_bytes[0] / 4) / max_slot_words;
        if (have_slots < n_slots && have_slots * 2 >= n_slots) n_slots = have_slots;
    }
    if (max_slot_words > (24ll << 30) / 4) return fail(ctx, SP_ERR_NOMEM, "sp_align_pairs: traceback scratch of one pair exceeds 24 GB");
    const int warps_per_cta = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(K1_WARPS, (n_slots + ctx->num_sms - 1) / ctx->num_sms)));
    const int grid = static_cast<int>((n_slots + warps_per_cta - 1) / warps_per_cta);

    sp_seqset tset = {tb.data(), to.data(), static_cast<int64_t>(t_ids.size())};
    sp_seqset pset = {pbs.data(), po.data(), np};
    uint8_t *d_tb = nullptr, *d_pb = nullptr; long long *d_to = nullptr, *d_po = nullptr, *d_cig_off = nullptr, *d_out_off = nullptr;
    int32_t *d_lane_pat = nullptr, *d_lane_row0 = nullptr, *d_pt = nullptr, *d_pp = nullptr, *d_wb = nullptr, *d_we = nullptr;
    uint32_t *d_lane_info1 = nullptr, *d_blobs = nullptr, *d_cigar = nullptr, *d_scratch = nullptr, *d_dense = nullptr;
    AlignRecDev *d_recs = nullptr;
    auto cleanup = [&]() {
        dev_free(ctx, d_wb); dev_free(ctx, d_we);
        dev_free(ctx, d_tb); dev_free(ctx, d_pb); dev_free(ctx, d_to); dev_free(ctx, d_po); dev_free(ctx, d_cig_off); dev_free(ctx, d_out_off);
        dev_free(ctx, d_lane_pat); dev_free(ctx, d_lane_row0); dev_free(ctx, d_pt); dev_free(ctx, d_pp); dev_free(ctx, d_lane_info1);
        dev_free(ctx, d_recs);  // d_scratch, d_cigar, d_blobs and d_dense live in the context's pools
    };
    auto cu = [&](cudaError_t e, const char *what) -> sp_status {
        if (e != cudaSuccess)
            return fail(ctx, e == cudaErrorMemoryAllocation ? SP_ERR_NOMEM : SP_ERR_CUDA,
                        std::string("sp_align_pairs: ") + what + ": " + cudaGetErrorString(e));
        return SP_OK;
    };
#define SP_TRY(x)                                   \
    do {                                            \
        sp_status s__ = (x);                        \
        if (s__ != SP_OK) { cleanup(); return s__; } \
    } while (0)
    auto up = [&](void **dst, const void *src, size_t bytes) -> sp_status {
        sp_status s = cu(dev_malloc(ctx, dst, std::max<size_t>(bytes, 16)), "cudaMalloc");
        if (s == SP_OK && bytes) s = cu(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D");
        return s;
    };
    SP_TRY(upload_seqset(ctx, &tset, &d_tb, &d_to));
    SP_TRY(upload_seqset(ctx, &pset, &d_pb, &d_po));
    const size_t tab = static_cast<size_t>(np) * 32;
    std::vector<int32_t> lane_pat(tab, -1), lane_row0(tab, 0);
    std::vector<uint32_t> lane_info1(tab, INFO_FIRST);
    for (int64_t i = 0; i < np; ++i) {
        const int64_t m = po[static_cast<size_t>(i) + 1] - po[static_cast<size_t>(i)];
        if (m == 0) continue;
        const int64_t nl = (m + rows - 1) / rows, pad = nl * rows - m;
        for (int64_t li = 0; li < nl; ++li) {
            const size_t o = static_cast<size_t>(i) * 32 + static_cast<size_t>(li);
            lane_pat[o] = static_cast<int32_t>(i);
            lane_row0[o] = static_cast<int32_t>(li * rows - pad);
            lane_info1[o] = static_cast<uint32_t>(m) | (li == 0 ? INFO_FIRST : 0u) | (li == nl - 1 ? INFO_LAST : 0u);
        }
    }
    SP_TRY(up(reinterpret_cast<void **>(&d_lane_pat), lane_pat.data(), tab * 4));
    SP_TRY(up(reinterpret_cast<void **>(&d_lane_row0), lane_row0.data(), tab * 4));
    SP_TRY(up(reinterpret_cast<void **>(&d_lane_info1), lane_info1.data(), tab * 4));
    SP_TRY(up(reinterpret_cast<void **>(&d_pt), pt.data(), pt.size() * 4));
    SP_TRY(up(reinterpret_cast<void **>(&d_pp), pp.data(), pp.size() * 4));
    if (win_begin) {
        SP_TRY(up(reinterpret_cast<void **>(&d_wb), win_begin, static_cast<size_t>(n_pairs) * 4));
        SP_TRY(up(reinterpret_cast<void **>(&d_we), win_end, static_cast<size_t>(n_pairs) * 4));
    }
    SP_TRY(up(reinterpret_cast<void **>(&d_cig_off), cig_off.data(), cig_off.size() * sizeof(long long)));
    SP_TRY(cu(ctx_pool(ctx, 2, static_cast<size_t>(np) * blob_words(ALN_U) * 4, reinterpret_cast<void **>(&d_blobs)), "blob pool"));
    SP_TRY(cu(ctx_pool(ctx, 1, static_cast<size_t>(cig_off.back()) * 4, reinterpret_cast<void **>(&d_cigar)), "cigar pool"));
    SP_TRY(cu(dev_malloc(ctx, reinterpret_cast<void **>(&d_recs), static_cast<size_t>(n_pairs) * sizeof(AlignRecDev)), "cudaMalloc recs"));
    SP_TRY(cu(ctx_scratch(ctx, static_cast<size_t>(grid) * warps_per_cta * max_slot_words * 4, reinterpret_cast<void **>(&d_scratch)),
              "traceback scratch"));
    tm.mark("upload + cudaMalloc");
    SP_TRY(sp_internal_pack_blobs(ctx, d_pb, d_po, d_lane_pat, d_lane_row0, d_lane_info1, d_blobs, static_cast<int>(np), ALN_U, 0, 0));
    {
        AlignParams prm;
        prm.blobs = d_blobs; prm.tbases = d_tb; prm.toffs = d_to; prm.pair_t = d_pt; prm.pair_p = d_pp;
        prm.win_begin = d_wb; prm.win_end = d_we;
        prm.cig_off = d_cig_off; prm.cigar = d_cigar; prm.scratch = d_scratch; prm.slot_words = max_slot_words;
        prm.recs = d_recs; prm.n_pairs = static_cast<int>(n_pairs); prm.one = 1u; prm.m1 = 0xFFFFFFFFu; prm.seed_a = 1u; prm.seed_b = 0xFFFFFFFFu;
        const size_t smem = static_cast<size_t>(warps_per_cta) * blob_words(ALN_U) * 4;
        SP_TRY(cu(cudaFuncSetAttribute(k4_align, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_WARPS * blob_words(ALN_U) * 4), "k4_align smem"));
        ev_begin(ctx, 4);
        k4_align<<<grid, 32 * warps_per_cta, smem, ctx->stream>>>(prm);
        ev_end(ctx, 4);
        ++ctx->launches;
        SP_TRY(cu(cudaGetLastError(), "k4_align launch"));
    }
    std::vector<AlignRecDev> hrec(static_cast<size_t>(n_pairs));
    SP_TRY(cu(cudaMemcpyAsync(hrec.data(), d_recs, hrec.size() * sizeof(AlignRecDev), cudaMemcpyDeviceToHost, ctx->stream), "D2H recs"));
    SP_TRY(cu(cudaStreamSynchronize(ctx->stream), "k4_align"));
    tm.mark("pack + k4_align + recs D2H");
    std::vector<long long> out_off(static_cast<size_t>(n_pairs) + 1, 0);
    for (int64_t q = 0; q < n_pairs; ++q) out_off[static_cast<size_t>(q) + 1] = out_off[static_cast<size_t>(q)] + hrec[static_cast<size_t>(q)].n_cigar;
    const int64_t total = out_off.back();
    if (cigar_used) *cigar_used = total;
    if (total > cigar_cap) {
        cleanup();
        return fail(ctx, SP_ERR_RANGE, "sp_align_pairs: cigar buffer too small: " + std::to_string(total) + " entries needed");
    }
    if (total > 0) {
        SP_TRY(up(reinterpret_cast<void **>(&d_out_off), out_off.data(), out_off.size() * sizeof(long long)));
        SP_TRY(cu(ctx_pool(ctx, 3, static_cast<size_t>(total) * 4, reinterpret_cast<void **>(&d_dense)), "dense cigar pool"));
        k4_compact_cigar<<<static_cast<unsigned>(n_pairs), 128, 0, ctx->stream>>>(d_recs, d_cigar, d_out_off, d_dense);
        ++ctx->launches;
        SP_TRY(cu(cudaGetLastError(), "k4_compact_cigar launch"));
        SP_TRY(cu(cudaMemcpyAsync(cigar, d_dense, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H cigar"));
        SP_TRY(cu(cudaStreamSynchronize(ctx->stream), "k4_compact_cigar"));
    }
    for (int64_t q = 0; q < n_pairs; ++q) {
        const AlignRecDev &r = hrec[static_cast<size_t>(q)];
        sp_align_rec &o = recs[q];
        o.dist = r.dist; o.nm = r.nm; o.p_start = r.p_start; o.p_end = r.p_end; o.t_start = r.t_start; o.t_end = r.t_end;
        o.n_cigar = r.n_cigar; o._pad = 0; o.cigar_off = out_off[static_cast<size_t>(q)];
    }
#undef SP_TRY
    tm.mark("compact + cigar D2H");
    cleanup();
    tm.mark("cudaFree");
    return SP_OK;
}

