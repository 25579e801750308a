// sp_consensus.cu -- K7 host side: sp_consensus_* (include/starphase_gpu.h), row N1 of SURVEY.md 8f.
#include "sp_internal.cuh"

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
};

static uint8_t code_of(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
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
        if (symbols) sym[static_cast<size_t>(q)] = symbols[q] == 0 ? 255 : code_of(symbols[q]);  // symbol 0: report the track's state only
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
