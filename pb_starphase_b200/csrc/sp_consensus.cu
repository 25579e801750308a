// sp_consensus.cu -- K7 host side: sp_consensus_* (include/starphase_gpu.h), row N1 of SURVEY.md 8f.
#include "sp_internal.cuh"

#include <cstring>

#include "sp_consensus.cuh"

using namespace sp;

struct sp_consensus {
    sp_ctx *ctx = nullptr;
    int n_reads = 0, n_tracks = 0, W = 0, half_window = 0, cells = 0;
    uint8_t *d_codes = nullptr;
    long long *d_roffs = nullptr;
    int32_t *d_offset = nullptr, *d_band = nullptr, *d_best_full = nullptr, *d_track_len = nullptr;
    // grow-only task buffers
    int32_t *d_src = nullptr, *d_dst = nullptr, *d_ed = nullptr, *d_full = nullptr;
    uint8_t *d_sym = nullptr, *d_votes = nullptr;
    int task_cap = 0;
    // sp_consensus_run: node state in / out, step log, step count
    // one device block and its page-locked mirror: [n_steps, pad x 3 | ed | full | votes] for 2 x n_reads read-sides; the log apart
    uint8_t *d_run = nullptr, *h_run = nullptr, *d_run_log = nullptr, *h_run_log = nullptr;
    int run_log_cap = 0;
    bool run_attr_set[4] = {false, false, false, false};
    int max_read_len = 0;
};

static uint8_t code_of(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '*': return 5;  // wildcard (waffle_con's CdwfaConfig::wildcard as the CYP2D6 caller sets it, src/cyp2d6/caller.rs:148)
        default: return 4;
    }
}

extern "C" sp_status sp_consensus_create(sp_ctx *ctx, const sp_seqset *reads, const int32_t *offsets, int32_t offset_window, int32_t band,
                                         int32_t max_tracks, sp_consensus **out) {
    if (!ctx) return SP_ERR_INVALID;
    if (!out) return fail(ctx, SP_ERR_INVALID, "sp_consensus_create: out is NULL");
    *out = nullptr;
    sp_status st = check_seqset(ctx, reads, "reads");
    if (st != SP_OK) return st;
    if (offset_window < 0 || band < 1 || max_tracks < 1) return fail(ctx, SP_ERR_INVALID, "sp_consensus_create: bad window / band / track count");
    const int W = band + offset_window / 2;  // the band has to hold the drift of the read and the uncertainty of where it starts
    if (2 * W + 1 > 32 * 32) return fail(ctx, SP_ERR_RANGE, "sp_consensus_create: band + offset_window / 2 must stay below 512");
    if (reads->n > 0x7FFFFF) return fail(ctx, SP_ERR_RANGE, "sp_consensus_create: too many reads");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    sp_consensus *c = new (std::nothrow) sp_consensus();
    if (!c) return fail(ctx, SP_ERR_NOMEM, "out of host memory");
    c->ctx = ctx; c->n_reads = static_cast<int>(reads->n); c->n_tracks = max_tracks; c->W = W; c->half_window = offset_window / 2;
    c->cells = (2 * W + 1 + 31) / 32;
    for (int64_t i = 0; i < reads->n; ++i) c->max_read_len = std::max<int>(c->max_read_len, static_cast<int>(std::min<int64_t>(reads->offsets[i + 1] - reads->offsets[i], 0x7FFFFFFF)));
    const int64_t base0 = reads->n ? reads->offsets[0] : 0, nbytes = reads->n ? reads->offsets[reads->n] - base0 : 0;
    std::vector<uint8_t> codes(static_cast<size_t>(std::max<int64_t>(nbytes, 1)));
    for (int64_t i = 0; i < nbytes; ++i) codes[static_cast<size_t>(i)] = code_of(reads->bases[base0 + i]);
    std::vector<long long> roffs(static_cast<size_t>(reads->n) + 1, 0);
    std::vector<int32_t> offs(static_cast<size_t>(std::max<int64_t>(reads->n, 1)), -1);
    for (int64_t i = 0; i <= reads->n; ++i) roffs[static_cast<size_t>(i)] = reads->n ? reads->offsets[i] - base0 : 0;
    for (int64_t i = 0; i < reads->n; ++i) {
        offs[static_cast<size_t>(i)] = offsets ? std::max(offsets[i], -1) : -1;
    }
    const size_t nr = static_cast<size_t>(std::max(c->n_reads, 1)), nb = static_cast<size_t>(2 * W + 1);
    cudaError_t e = dev_malloc(ctx, &c->d_codes, codes.size());
    if (e == cudaSuccess) e = dev_malloc(ctx, &c->d_roffs, roffs.size() * sizeof(long long));
    if (e == cudaSuccess) e = dev_malloc(ctx, &c->d_offset, offs.size() * 4);
    if (e == cudaSuccess) e = dev_malloc(ctx, &c->d_band, static_cast<size_t>(max_tracks) * nr * nb * 4);
    if (e == cudaSuccess) e = dev_malloc(ctx, &c->d_best_full, static_cast<size_t>(max_tracks) * nr * 4);
    if (e == cudaSuccess) e = dev_malloc(ctx, &c->d_track_len, static_cast<size_t>(max_tracks) * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_codes, codes.data(), codes.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_roffs, roffs.data(), roffs.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_offset, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_track_len, 0, static_cast<size_t>(max_tracks) * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        sp_consensus_destroy(c);
        return fail(ctx, e == cudaErrorMemoryAllocation ? SP_ERR_NOMEM : SP_ERR_CUDA, std::string("sp_consensus_create: ") + cudaGetErrorString(e));
    }
    for (int t = 0; t < max_tracks; ++t) {
        k7_reset<<<(c->n_reads + 127) / 128 + 1, 128, 0, ctx->stream>>>(c->d_best_full, c->d_track_len, t, c->n_reads);
        ++ctx->launches;
    }
    SP_CUDA(ctx, cudaGetLastError());
    *out = c;
    return SP_OK;
}

extern "C" void sp_consensus_destroy(sp_consensus *c) {
    if (!c) return;
    sp_ctx *ctx = c->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, c->d_codes); dev_free(ctx, c->d_roffs); dev_free(ctx, c->d_offset); dev_free(ctx, c->d_band);
    dev_free(ctx, c->d_best_full); dev_free(ctx, c->d_track_len);
    dev_free(ctx, c->d_src); dev_free(ctx, c->d_dst); dev_free(ctx, c->d_ed); dev_free(ctx, c->d_full); dev_free(ctx, c->d_sym); dev_free(ctx, c->d_votes);
    dev_free(ctx, c->d_run); dev_free(ctx, c->d_run_log);
    cudaFreeHost(c->h_run); cudaFreeHost(c->h_run_log);
    delete c;
}

extern "C" int32_t sp_consensus_num_reads(const sp_consensus *c) { return c ? c->n_reads : 0; }
extern "C" int32_t sp_consensus_num_tracks(const sp_consensus *c) { return c ? c->n_tracks : 0; }

extern "C" sp_status sp_consensus_reset(sp_consensus *c, int32_t track) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (track < 0 || track >= c->n_tracks) return fail(ctx, SP_ERR_INVALID, "sp_consensus_reset: track out of range");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    k7_reset<<<(c->n_reads + 127) / 128 + 1, 128, 0, ctx->stream>>>(c->d_best_full, c->d_track_len, track, c->n_reads);
    ++ctx->launches;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

template <int CELLS>
static void launch_k7(const ConsParams &prm, cudaStream_t stream) {
    const long long warps = static_cast<long long>(prm.n_tasks) * prm.n_reads;
    const unsigned grid = static_cast<unsigned>((warps * 32 + 127) / 128);
    k7_extend<CELLS><<<grid, 128, 0, stream>>>(prm);
}

extern "C" sp_status sp_consensus_extend(sp_consensus *c, int32_t n_tasks, const int32_t *src, const uint8_t *symbols, const int32_t *dst,
                                         int32_t *ed, uint8_t *votes, int32_t *full) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (n_tasks < 0 || (n_tasks > 0 && (!src || !dst || !ed || !votes || !full))) return fail(ctx, SP_ERR_INVALID, "sp_consensus_extend: bad argument");
    if (n_tasks == 0 || c->n_reads == 0) return SP_OK;
    std::vector<uint8_t> sym(static_cast<size_t>(n_tasks), 255);
    std::vector<char> is_dst(static_cast<size_t>(c->n_tracks), 0);
    for (int q = 0; q < n_tasks; ++q) {
        if (src[q] < 0 || src[q] >= c->n_tracks || dst[q] < 0 || dst[q] >= c->n_tracks) return fail(ctx, SP_ERR_INVALID, "sp_consensus_extend: track out of range");
        if (is_dst[static_cast<size_t>(dst[q])]) return fail(ctx, SP_ERR_INVALID, "sp_consensus_extend: two tasks write the same track");
        is_dst[static_cast<size_t>(dst[q])] = 1;
        if (symbols) sym[static_cast<size_t>(q)] = symbols[q] == 0 ? 255 : std::min<uint8_t>(code_of(symbols[q]), 4);  // symbol 0: report the track's state only
    }
    for (int q = 0; q < n_tasks; ++q)
        if (src[q] != dst[q] && is_dst[static_cast<size_t>(src[q])])
            return fail(ctx, SP_ERR_INVALID, "sp_consensus_extend: a track is read by one task and written by another");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_tasks > c->task_cap) {
        cudaStreamSynchronize(ctx->stream);
        dev_free(ctx, c->d_src); dev_free(ctx, c->d_dst); dev_free(ctx, c->d_ed); dev_free(ctx, c->d_full); dev_free(ctx, c->d_sym); dev_free(ctx, c->d_votes);
        c->d_src = c->d_dst = c->d_ed = c->d_full = nullptr; c->d_sym = c->d_votes = nullptr;
        const int cap = std::max(16, n_tasks * 2);
        const size_t per = static_cast<size_t>(cap) * c->n_reads;
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_src, static_cast<size_t>(cap) * 4));
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_dst, static_cast<size_t>(cap) * 4));
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_sym, static_cast<size_t>(cap)));
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_ed, per * 4));
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_full, per * 4));
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_votes, per));
        c->task_cap = cap;
    }
    const size_t per = static_cast<size_t>(n_tasks) * c->n_reads;
    SP_CUDA(ctx, cudaMemcpyAsync(c->d_src, src, static_cast<size_t>(n_tasks) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(c->d_dst, dst, static_cast<size_t>(n_tasks) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(c->d_sym, sym.data(), static_cast<size_t>(n_tasks), cudaMemcpyHostToDevice, ctx->stream));
    ConsParams prm;
    prm.codes = c->d_codes; prm.roffs = c->d_roffs; prm.offset = c->d_offset; prm.band = c->d_band; prm.best_full = c->d_best_full;
    prm.track_len = c->d_track_len; prm.src = c->d_src; prm.dst = c->d_dst; prm.sym = c->d_sym; prm.out_ed = c->d_ed; prm.out_votes = c->d_votes;
    prm.out_full = c->d_full; prm.n_reads = c->n_reads; prm.n_tasks = n_tasks; prm.W = c->W; prm.half_window = c->half_window;
    switch (c->cells) {
#define SP_CASE(n) case n: launch_k7<n>(prm, ctx->stream); break
        SP_CASE(1); SP_CASE(2); SP_CASE(3); SP_CASE(4); SP_CASE(5); SP_CASE(6); SP_CASE(7); SP_CASE(8); SP_CASE(9); SP_CASE(10); SP_CASE(11);
        SP_CASE(12); SP_CASE(13); SP_CASE(14); SP_CASE(15); SP_CASE(16); SP_CASE(17); SP_CASE(18); SP_CASE(19); SP_CASE(20); SP_CASE(21);
        SP_CASE(22); SP_CASE(23); SP_CASE(24); SP_CASE(25); SP_CASE(26); SP_CASE(27); SP_CASE(28); SP_CASE(29); SP_CASE(30); SP_CASE(31);
        SP_CASE(32);
#undef SP_CASE
        default: return fail(ctx, SP_ERR_RANGE, "sp_consensus_extend: band too wide");
    }
    k7_bump_lengths<<<(n_tasks + 127) / 128, 128, 0, ctx->stream>>>(prm);
    ctx->launches += 2;
    SP_CUDA(ctx, cudaGetLastError());
    SP_CUDA(ctx, cudaMemcpyAsync(ed, c->d_ed, per * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(full, c->d_full, per * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(votes, c->d_votes, per, cudaMemcpyDeviceToHost, ctx->stream));
    SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SP_OK;
}

// ---- sp_consensus_run: K7 in a loop on the device (k7_run) ----
namespace {
struct RunPlan {
    int cells = 0, n_cta = 0;
    size_t smem = 0, smem_max = 0;  // of this launch; of the larger of the 1- and 2-sided launches (the attribute is set once)
};
// cells per lane rounded up to {2, 4, 8, 16}; cluster size and shared memory of one CTA; n_cta = 0: does not fit
RunPlan plan_run(const sp_consensus *c, int n_sides) {
    RunPlan pl;
    const int nb = 2 * c->W + 1, items = c->n_reads * n_sides;
    if (nb > 511 || items < 1 || items > 2048 || c->max_read_len > 60000) return pl;  // columns are kept as u16
    pl.cells = nb < 64 ? 2 : nb < 128 ? 4 : nb < 256 ? 8 : 16;
    const int nwarps = K7_RUN_THREADS / 32;
    int n_cta = 1;
    while (n_cta < 8 && n_cta * nwarps < items) n_cta *= 2;
    if (const char *force = getenv("SP_K7_RUN_CTAS")) n_cta = std::max(1, std::min(8, atoi(force)));  // measurement hook
    const int per_warp = (items + n_cta * nwarps - 1) / (n_cta * nwarps);
    // K7Item + column (u16) + code ring (2 bytes per cell) per slot; ed / full / votes double-buffered per item
    pl.smem = static_cast<size_t>(nwarps) * per_warp * (sizeof(K7Item) + static_cast<size_t>(pl.cells) * 32 * 4) + static_cast<size_t>(items) * 18 + 16;
    const size_t room = static_cast<size_t>(std::max(c->ctx->smem_optin - 1024, 0));  // the kernel's few static bytes come out of the same budget
    if (pl.smem > room) return pl;
    pl.smem_max = room;
    pl.n_cta = n_cta;
    return pl;
}
template <int CELLS>
cudaError_t launch_run(const ConsRunParams &prm, const RunPlan &pl, cudaStream_t stream, bool &attr_set) {
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(k7_run<CELLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem_max));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(pl.n_cta)); cfg.blockDim = dim3(K7_RUN_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(pl.n_cta); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k7_run<CELLS>, prm);
}
}  // namespace

extern "C" int32_t sp_consensus_run_supported(const sp_consensus *c, int32_t n_sides) {
    if (!c || n_sides < 1 || n_sides > 2) return 0;
    return plan_run(c, n_sides).n_cta > 0 ? 1 : 0;
}

extern "C" sp_status sp_consensus_run(sp_consensus *c, int32_t n_sides, const int32_t *src, const int32_t *dst, int32_t *ed, uint8_t *votes,
                                      int32_t *full, int32_t min_count, int32_t min_af_permille, int64_t cost_limit, int64_t size_limit,
                                      int64_t cost_cap, int32_t max_steps, uint8_t *steps, int32_t *n_steps) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (n_sides < 1 || n_sides > 2 || !src || !dst || !ed || !votes || !full || !steps || !n_steps || max_steps < 1 || min_count < 0 ||
        min_af_permille < 0 || min_af_permille > 1000)
        return fail(ctx, SP_ERR_INVALID, "sp_consensus_run: bad argument");
    *n_steps = 0;
    for (int s = 0; s < n_sides; ++s)
        if (src[s] < 0 || src[s] >= c->n_tracks || dst[s] < 0 || dst[s] >= c->n_tracks) return fail(ctx, SP_ERR_INVALID, "sp_consensus_run: track out of range");
    if (n_sides == 2 && (dst[0] == dst[1] || dst[0] == src[1] || dst[1] == src[0]))
        return fail(ctx, SP_ERR_INVALID, "sp_consensus_run: the two sides need separate destination tracks");
    if (c->n_reads == 0) return SP_OK;
    const RunPlan pl = plan_run(c, n_sides);
    if (pl.n_cta == 0) return fail(ctx, SP_ERR_RANGE, "sp_consensus_run: the read set does not fit on chip (sp_consensus_run_supported)");
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t items = static_cast<size_t>(c->n_reads) * n_sides, cap = static_cast<size_t>(c->n_reads) * 2;
    const size_t o_ed = 16, o_full = o_ed + cap * 4, o_votes = o_full + cap * 4, block = o_votes + (cap + 3) / 4 * 4;
    constexpr int kLogHead = 256;  // log bytes fetched with the state; longer runs take one more copy
    if (!c->d_run) {
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_run, block));
        SP_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&c->h_run), block, cudaHostAllocDefault));
    }
    if (max_steps > c->run_log_cap) {
        cudaStreamSynchronize(ctx->stream);
        dev_free(ctx, c->d_run_log);
        cudaFreeHost(c->h_run_log);
        c->d_run_log = nullptr; c->h_run_log = nullptr; c->run_log_cap = 0;
        const int want = std::max(max_steps, kLogHead);
        SP_CUDA(ctx, dev_malloc(ctx, &c->d_run_log, static_cast<size_t>(want)));
        SP_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&c->h_run_log), static_cast<size_t>(want), cudaHostAllocDefault));
        c->run_log_cap = want;
    }
    std::memset(c->h_run, 0, 16);
    std::memcpy(c->h_run + o_ed, ed, items * 4);
    std::memcpy(c->h_run + o_full, full, items * 4);
    std::memcpy(c->h_run + o_votes, votes, items);
    SP_CUDA(ctx, cudaMemcpyAsync(c->d_run, c->h_run, block, cudaMemcpyHostToDevice, ctx->stream));
    ConsRunParams prm;
    prm.codes = c->d_codes; prm.roffs = c->d_roffs; prm.offset = c->d_offset; prm.band = c->d_band; prm.best_full = c->d_best_full;
    prm.track_len = c->d_track_len; prm.n_reads = c->n_reads; prm.W = c->W; prm.half_window = c->half_window; prm.n_sides = n_sides;
    for (int s = 0; s < 2; ++s) { prm.src[s] = src[s < n_sides ? s : 0]; prm.dst[s] = dst[s < n_sides ? s : 0]; }
    prm.ed = reinterpret_cast<int32_t *>(c->d_run + o_ed); prm.full = reinterpret_cast<int32_t *>(c->d_run + o_full); prm.votes = c->d_run + o_votes;
    prm.min_units = 12 * min_count; prm.permille = min_af_permille; prm.cost_limit = cost_limit; prm.size_limit = size_limit; prm.cost_cap = cost_cap; prm.max_steps = max_steps;
    prm.log = c->d_run_log; prm.n_steps = reinterpret_cast<int *>(c->d_run);
    const int ci = pl.cells == 2 ? 0 : pl.cells == 4 ? 1 : pl.cells == 8 ? 2 : 3;
    ev_begin(ctx, 4);
    cudaError_t e = ci == 0 ? launch_run<2>(prm, pl, ctx->stream, c->run_attr_set[0]) : ci == 1 ? launch_run<4>(prm, pl, ctx->stream, c->run_attr_set[1])
                    : ci == 2 ? launch_run<8>(prm, pl, ctx->stream, c->run_attr_set[2]) : launch_run<16>(prm, pl, ctx->stream, c->run_attr_set[3]);
    SP_CUDA(ctx, e);
    ++ctx->launches;
    ev_end(ctx, 4);
    const int head = std::min(max_steps, kLogHead);
    SP_CUDA(ctx, cudaMemcpyAsync(c->h_run, c->d_run, block, cudaMemcpyDeviceToHost, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(c->h_run_log, c->d_run_log, static_cast<size_t>(head), cudaMemcpyDeviceToHost, ctx->stream));
    SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int done = 0;
    std::memcpy(&done, c->h_run, sizeof(int));
    if (done > head) {
        SP_CUDA(ctx, cudaMemcpyAsync(c->h_run_log + head, c->d_run_log + head, static_cast<size_t>(done - head), cudaMemcpyDeviceToHost, ctx->stream));
        SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::memcpy(ed, c->h_run + o_ed, items * 4);
    std::memcpy(full, c->h_run + o_full, items * 4);
    std::memcpy(votes, c->h_run + o_votes, items);
    if (done > 0) std::memcpy(steps, c->h_run_log, static_cast<size_t>(done));
    if (getenv("SP_TIMING")) fprintf(stderr, "[sp_timing] k7_run %d sides %d steps %.1f us\n", n_sides, done, 1e3 * sp_last_kernel_ms(ctx, 4));
    *n_steps = done;
    return SP_OK;
}
