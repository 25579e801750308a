// sp_comm.cu -- the scoring path over the GPUs of one box, behind the C ABI.
//
// One sp_comm per (GPU, sp_ctx); the ranks of a communicator may be processes (torchrun-style, one per GPU) or threads of one
// process (a single-process host such as the reference's Rust binary drives each context from its own thread).  The path
// shards without any exchange inside K1: the allele set is dealt to the ranks (sp_shard_plan), the read set is broadcast
// (ncclBroadcast), every rank scores its shard straight into its slot of an all-gather buffer, ncclAllGather replicates the
// distance matrix (51 MB at BASELINE configs[1]), a row permutation puts it into database order, K2 runs on equal-area row
// ranges of the pair triangle and the k-record lists are all-gathered and merged by the library's total order
// (score, score2, i, j) -- the answer is bit-identical for every world size.
//
// NCCL is loaded with dlopen at the first sp_comm_* call ($SP_NCCL_LIB, then libnccl.so.2 -- the copy the embedding process
// already loaded, e.g. torch's, else the system one): libstarphase_gpu.so has no link-time NCCL dependency and single-GPU
// hosts never touch it.
#include "sp_internal.cuh"

#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

#include "sp_comm_kernels.cuh"

using namespace sp;

namespace {

struct NcclApi {
    void *handle = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char *names[] = {getenv("SP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
        g_nccl.err = dlerror();
    }
    if (!g_nccl.handle) return;
    auto sym = [&](const char *name) -> void * {
        void *p = dlsym(g_nccl.handle, name);
        if (!p) { g_nccl.err = std::string("NCCL symbol missing: ") + name; }
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(sym("ncclBroadcast"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.Broadcast ||
        !g_nccl.GetErrorString) {
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
    }
}

const NcclApi *nccl(sp_ctx *ctx) {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.handle) {
        fail(ctx, SP_ERR_CUDA, "NCCL is not available (set SP_NCCL_LIB to libnccl.so.2): " + g_nccl.err);
        return nullptr;
    }
    return &g_nccl;
}

}  // namespace

struct sp_comm {
    sp_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // shard layout of the last sp_comm_score_allgather (re-used while the same shard handle comes back)
    const sp_patterns *plan_shard = nullptr;
    int64_t plan_total = 0, plan_slot_rows = 0;
    int32_t *d_row_global = nullptr;  // [world * slot_rows]
};

#define SP_NCCL(c, call)                                                                                          \
    do {                                                                                                          \
        ncclResult_t r__ = (call);                                                                                \
        if (r__ != ncclSuccess)                                                                                   \
            return fail((c)->ctx, SP_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__) + " (" __FILE__ ":" + \
                                                   std::to_string(__LINE__) + ")");                                 \
    } while (0)

static_assert(SP_COMM_ID_BYTES == sizeof(ncclUniqueId), "SP_COMM_ID_BYTES must match ncclUniqueId");

extern "C" sp_status sp_comm_unique_id(uint8_t *id) {
    if (!id) return fail(nullptr, SP_ERR_INVALID, "sp_comm_unique_id: id is NULL");
    const NcclApi *api = nccl(nullptr);
    if (!api) return SP_ERR_CUDA;
    ncclUniqueId uid;
    ncclResult_t r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) return fail(nullptr, SP_ERR_CUDA, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
    memcpy(id, &uid, sizeof(uid));
    return SP_OK;
}

extern "C" sp_status sp_comm_create(sp_ctx *ctx, const uint8_t *id, int rank, int world, sp_comm **out) {
    if (!ctx) return SP_ERR_INVALID;
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, SP_ERR_INVALID, "sp_comm_create: bad argument");
    *out = nullptr;
    sp_comm *c = new (std::nothrow) sp_comm();
    if (!c) return fail(ctx, SP_ERR_NOMEM, "out of host memory");
    c->ctx = ctx; c->rank = rank; c->world = world;
    if (world > 1) {
        const NcclApi *api = nccl(ctx);
        if (!api) { delete c; return SP_ERR_CUDA; }
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) { delete c; return fail(ctx, SP_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
        if (r != ncclSuccess) {
            delete c;
            return fail(ctx, SP_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
        }
    }
    *out = c;
    return SP_OK;
}

extern "C" void sp_comm_destroy(sp_comm *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    dev_free(c->ctx, c->d_row_global);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
}

extern "C" int sp_comm_rank(const sp_comm *c) { return c ? c->rank : 0; }
extern "C" int sp_comm_world(const sp_comm *c) { return c ? c->world : 0; }

// ------------------------------------------------------------------------------------------
// host-only planning
// ------------------------------------------------------------------------------------------
// Deals the patterns to the ranks in length order, snake-wise (0 1 .. w-1, w-1 .. 1 0, ...): every rank receives the same
// length distribution, so each shard packs into the same lane-width classes with the same share of real rows as the whole set
// (contiguous index ranges did not: one rank received a single-gene shard that packed 2.3 % worse and every rank waited for it).
extern "C" sp_status sp_shard_plan(const int64_t *lens, int64_t n, int world, int rank, int64_t *idx, int64_t *n_idx) {
    if (n < 0 || (n > 0 && !lens) || world < 1 || rank < 0 || rank >= world || !n_idx || (n > 0 && !idx)) return SP_ERR_INVALID;
    std::vector<int64_t> order(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        if (lens[i] < 0) return SP_ERR_INVALID;
        order[static_cast<size_t>(i)] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return lens[a] > lens[b]; });
    int64_t cnt = 0;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t round = k / world, pos = k % world;
        const int64_t owner = (round & 1) ? world - 1 - pos : pos;
        if (owner == rank) idx[cnt++] = order[static_cast<size_t>(k)];
    }
    std::sort(idx, idx + cnt);
    *n_idx = cnt;
    return SP_OK;
}

// Row range [lo, hi) of the pair triangle i <= j < n owned by `rank`: (nearly) equal pair counts per rank.
extern "C" sp_status sp_triangle_rows(int64_t n, int world, int rank, int64_t *lo, int64_t *hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return SP_ERR_INVALID;
    auto edge = [&](int b) -> int64_t {
        if (b <= 0) return 0;
        if (b >= world) return n;
        // smallest i with (pairs in rows < i) >= b / world of all pairs; pairs in rows < i = i*n - i*(i-1)/2
        const long double total = static_cast<long double>(n) * (n + 1) / 2, want = total * b / world;
        int64_t a = 0, z = n;
        while (a < z) {
            const int64_t mid = (a + z) / 2;
            const long double have = static_cast<long double>(mid) * n - static_cast<long double>(mid) * (mid - 1) / 2;
            if (have >= want) z = mid; else a = mid + 1;
        }
        return a;
    };
    *lo = edge(rank);
    *hi = edge(rank + 1);
    return SP_OK;
}

// ------------------------------------------------------------------------------------------
// read-set broadcast
// ------------------------------------------------------------------------------------------
extern "C" sp_status sp_comm_bcast_targets(sp_comm *c, const sp_seqset *targets, int root, sp_targets **out) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (!out || root < 0 || root >= c->world) return fail(ctx, SP_ERR_INVALID, "sp_comm_bcast_targets: bad argument");
    *out = nullptr;
    const bool is_root = c->rank == root;
    if (is_root) {
        sp_status st = check_seqset(ctx, targets, "targets");
        if (st != SP_OK) return st;
    }
    if (c->world == 1) return sp_targets_create(ctx, targets, out);
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    long long *d_hdr = nullptr;
    long long hdr[2] = {0, 0};  // n, bytes
    if (is_root) {
        hdr[0] = targets->n;
        hdr[1] = targets->n ? targets->offsets[targets->n] - targets->offsets[0] : 0;
    }
    SP_CUDA(ctx, dev_malloc(ctx, &d_hdr, sizeof(hdr)));
    auto bail = [&](sp_status st) { dev_free(ctx, d_hdr); return st; };
    if (is_root && cudaMemcpyAsync(d_hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return bail(fail(ctx, SP_ERR_CUDA, "sp_comm_bcast_targets: H2D header"));
    if (g_nccl.Broadcast(d_hdr, d_hdr, 2, ncclInt64, root, c->comm, ctx->stream) != ncclSuccess)
        return bail(fail(ctx, SP_ERR_CUDA, "sp_comm_bcast_targets: ncclBroadcast header"));
    if (cudaMemcpyAsync(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, SP_ERR_CUDA, "sp_comm_bcast_targets: D2H header"));
    dev_free(ctx, d_hdr);
    d_hdr = nullptr;
    const int64_t n = hdr[0], nbytes = hdr[1];
    std::vector<int64_t> offs(static_cast<size_t>(n) + 1, 0);
    if (is_root)
        for (int64_t i = 0; i <= n; ++i) offs[static_cast<size_t>(i)] = n ? targets->offsets[i] - targets->offsets[0] : 0;
    uint8_t *d_bases = nullptr;
    long long *d_offs = nullptr;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    SP_CUDA(ctx, dev_malloc(ctx, &d_bases, static_cast<size_t>(std::max<int64_t>(nbytes, 16))));
    cudaError_t e = dev_malloc(ctx, &d_offs, offs.size() * sizeof(long long));
    auto bail2 = [&](sp_status st) { dev_free(ctx, d_bases); dev_free(ctx, d_offs); return st; };
    if (e != cudaSuccess) return bail2(fail(ctx, SP_ERR_NOMEM, "sp_comm_bcast_targets: cudaMalloc"));
    if (is_root) {
        e = cudaMemcpyAsync(d_offs, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && nbytes)
            e = cudaMemcpyAsync(d_bases, targets->bases + targets->offsets[0], static_cast<size_t>(nbytes), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return bail2(fail(ctx, SP_ERR_CUDA, std::string("sp_comm_bcast_targets: H2D: ") + cudaGetErrorString(e)));
    }
    if (g_nccl.Broadcast(d_offs, d_offs, offs.size(), ncclInt64, root, c->comm, ctx->stream) != ncclSuccess ||
        (nbytes && g_nccl.Broadcast(d_bases, d_bases, static_cast<size_t>(nbytes), ncclUint8, root, c->comm, ctx->stream) != ncclSuccess))
        return bail2(fail(ctx, SP_ERR_CUDA, "sp_comm_bcast_targets: ncclBroadcast"));
    if (!is_root) e = cudaMemcpyAsync(offs.data(), d_offs, offs.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail2(fail(ctx, SP_ERR_CUDA, std::string("sp_comm_bcast_targets: ") + cudaGetErrorString(e)));
    sp_status st = sp_internal_targets_adopt(ctx, d_bases, d_offs, offs.data(), n, out);
    if (st != SP_OK) return bail2(st);
    return SP_OK;
}

// ------------------------------------------------------------------------------------------
// sharded K1 + all-gather into database order
// ------------------------------------------------------------------------------------------
static sp_status plan_layout(sp_comm *c, const sp_patterns *shard, const int64_t *shard_idx, int64_t n_total) {
    sp_ctx *ctx = c->ctx;
    if (c->plan_shard == shard && c->plan_total == n_total && c->d_row_global) return SP_OK;
    dev_free(ctx, c->d_row_global);
    c->d_row_global = nullptr; c->plan_shard = nullptr;
    const int64_t n_mine = sp_patterns_count(shard);
    // slot size = the largest shard: all-gather of the counts
    long long *d_cnt = nullptr;
    SP_CUDA(ctx, dev_malloc(ctx, &d_cnt, static_cast<size_t>(c->world) * sizeof(long long)));
    const long long mine = n_mine;
    std::vector<long long> cnt(static_cast<size_t>(c->world), 0);
    cudaError_t e = cudaMemcpyAsync(d_cnt + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && g_nccl.AllGather(d_cnt + c->rank, d_cnt, 1, ncclInt64, c->comm, ctx->stream) != ncclSuccess) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d_cnt);
    if (e != cudaSuccess) return fail(ctx, SP_ERR_CUDA, "sp_comm_score_allgather: shard-size exchange failed");
    int64_t slot = 1, sum = 0;
    for (long long x : cnt) { slot = std::max<int64_t>(slot, x); sum += x; }
    if (sum != n_total) return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: the shards hold " + std::to_string(sum) + " patterns, n_total says " + std::to_string(n_total));
    if (n_total > 0x7FFFFFF0ll) return fail(ctx, SP_ERR_RANGE, "sp_comm_score_allgather: too many patterns");
    // database index of every slot row: all-gather of the index lists, then a host check that they form a permutation
    const size_t rows = static_cast<size_t>(c->world) * static_cast<size_t>(slot);
    std::vector<int32_t> mine_idx(static_cast<size_t>(slot), -1), all(rows, -1);
    for (int64_t i = 0; i < n_mine; ++i) {
        if (shard_idx[i] < 0 || shard_idx[i] >= n_total) return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: shard index outside [0, n_total)");
        mine_idx[static_cast<size_t>(i)] = static_cast<int32_t>(shard_idx[i]);
    }
    SP_CUDA(ctx, dev_malloc(ctx, &c->d_row_global, rows * sizeof(int32_t)));
    int32_t *slot_ptr = c->d_row_global + static_cast<size_t>(c->rank) * slot;
    e = cudaMemcpyAsync(slot_ptr, mine_idx.data(), mine_idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && g_nccl.AllGather(slot_ptr, c->d_row_global, static_cast<size_t>(slot), ncclInt32, c->comm, ctx->stream) != ncclSuccess) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(all.data(), c->d_row_global, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SP_ERR_CUDA, "sp_comm_score_allgather: shard-index exchange failed");
    std::vector<uint8_t> seen(static_cast<size_t>(n_total), 0);
    for (int32_t g : all)
        if (g >= 0) {
            if (seen[static_cast<size_t>(g)]) return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: pattern " + std::to_string(g) + " is in two shards");
            seen[static_cast<size_t>(g)] = 1;
        }
    c->plan_shard = shard; c->plan_total = n_total; c->plan_slot_rows = slot;
    return SP_OK;
}

extern "C" sp_status sp_comm_score_allgather(sp_comm *c, const sp_targets *t, const sp_patterns *shard, const int64_t *shard_idx,
                                             int64_t n_total, int elem_bits, sp_dmatrix **out) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (!t || !shard || !out || (sp_patterns_count(shard) > 0 && !shard_idx) || n_total < 0)
        return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: bad argument");
    if (elem_bits != 16 && elem_bits != 32) return fail(ctx, SP_ERR_INVALID, "elem_bits must be 16 or 32");
    *out = nullptr;
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (c->world == 1) {
        // one rank: the shard is the database (any order given by shard_idx is honoured through the same permutation path below
        // only when it is not the identity)
        bool identity = sp_patterns_count(shard) == n_total;
        for (int64_t i = 0; identity && i < n_total; ++i) identity = shard_idx[i] == i;
        if (identity) return sp_score_device(ctx, t, shard, elem_bits, 0, out);
    }
    const size_t eb = static_cast<size_t>(elem_bits / 8);
    const int64_t nt = sp_targets_count(t), ld = (nt + 63) / 64 * 64;
    int64_t slot = 0;
    if (c->world > 1) {
        sp_status st = plan_layout(c, shard, shard_idx, n_total);
        if (st != SP_OK) return st;
        slot = c->plan_slot_rows;
    } else {
        // world 1 with a non-identity order: build the row map locally
        slot = std::max<int64_t>(1, sp_patterns_count(shard));
        if (sp_patterns_count(shard) != n_total) return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: world 1 needs the whole database");
        if (!(c->plan_shard == shard && c->plan_total == n_total && c->d_row_global)) {
            dev_free(ctx, c->d_row_global); c->d_row_global = nullptr;
            std::vector<int32_t> map(static_cast<size_t>(slot), -1);
            std::vector<uint8_t> seen(static_cast<size_t>(n_total), 0);
            for (int64_t i = 0; i < n_total; ++i) {
                if (shard_idx[i] < 0 || shard_idx[i] >= n_total || seen[static_cast<size_t>(shard_idx[i])])
                    return fail(ctx, SP_ERR_INVALID, "sp_comm_score_allgather: shard_idx is not a permutation");
                seen[static_cast<size_t>(shard_idx[i])] = 1;
                map[static_cast<size_t>(i)] = static_cast<int32_t>(shard_idx[i]);
            }
            SP_CUDA(ctx, dev_malloc(ctx, &c->d_row_global, map.size() * sizeof(int32_t)));
            SP_CUDA(ctx, cudaMemcpyAsync(c->d_row_global, map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            c->plan_shard = shard; c->plan_total = n_total; c->plan_slot_rows = slot;
        }
    }
    // gather buffer: world slots of `slot` rows; this rank scores straight into its slot
    char *gbuf = nullptr;
    const size_t slot_bytes = static_cast<size_t>(slot) * static_cast<size_t>(ld) * eb;
    SP_CUDA(ctx, dev_malloc(ctx, &gbuf, std::max<size_t>(slot_bytes * static_cast<size_t>(c->world), 16)));
    sp_dmatrix *view = nullptr, *full = nullptr;
    auto bail = [&](sp_status st) { sp_dmatrix_destroy(view); sp_dmatrix_destroy(full); dev_free(ctx, gbuf); return st; };
    char *my_slot = gbuf + slot_bytes * static_cast<size_t>(c->rank);
    sp_status st = sp_dmatrix_wrap(ctx, my_slot, nt, slot, ld, elem_bits, &view);
    if (st != SP_OK) return bail(st);
    if (cudaMemsetAsync(my_slot, 0, slot_bytes, ctx->stream) != cudaSuccess) return bail(fail(ctx, SP_ERR_CUDA, "sp_comm_score_allgather: memset"));
    if (sp_patterns_count(shard) > 0) {
        sp_dmatrix *mine = nullptr;  // view of exactly the shard's rows (sp_score_into checks the row count)
        st = sp_dmatrix_wrap(ctx, my_slot, nt, sp_patterns_count(shard), ld, elem_bits, &mine);
        if (st == SP_OK) st = sp_score_into(ctx, t, shard, mine, 0);
        sp_dmatrix_destroy(mine);
        if (st != SP_OK) return bail(st);
    }
    if (c->world > 1 && slot_bytes > 0 &&
        g_nccl.AllGather(my_slot, gbuf, slot_bytes, ncclUint8, c->comm, ctx->stream) != ncclSuccess)
        return bail(fail(ctx, SP_ERR_CUDA, "sp_comm_score_allgather: ncclAllGather failed"));
    // database order
    full = new (std::nothrow) sp_dmatrix();
    if (!full) return bail(fail(ctx, SP_ERR_NOMEM, "out of host memory"));
    full->ctx = ctx; full->nt = nt; full->np = n_total; full->ld = ld; full->elem_bits = elem_bits; full->owned = true;
    const size_t full_bytes = std::max<size_t>(static_cast<size_t>(n_total) * static_cast<size_t>(ld) * eb, 16);
    if (dev_malloc(ctx, &full->d, full_bytes) != cudaSuccess) { full->d = nullptr; return bail(fail(ctx, SP_ERR_NOMEM, "sp_comm_score_allgather: cudaMalloc")); }
    if (n_total > 0 && nt > 0) {
        const long long row_vec4 = static_cast<long long>(ld) * static_cast<long long>(eb) / 16;
        comm_rows_to_database_order<<<static_cast<unsigned>(static_cast<size_t>(c->world) * static_cast<size_t>(slot)), 256, 0, ctx->stream>>>(
            reinterpret_cast<const uint4 *>(gbuf), static_cast<uint4 *>(full->d), c->d_row_global, row_vec4);
        ++ctx->launches;
        if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, SP_ERR_CUDA, "comm_rows_to_database_order launch failed"));
    }
    sp_dmatrix_destroy(view);
    dev_free(ctx, gbuf);  // stream-ordered: released after the permutation has read it
    *out = full;
    return SP_OK;
}

// ------------------------------------------------------------------------------------------
// sharded K2 + merge of the per-rank lists
// ------------------------------------------------------------------------------------------
extern "C" sp_status sp_comm_pair_minsum_topk(sp_comm *c, const sp_dmatrix *d, const sp_dmatrix *d2, int k, sp_pair_rec *out, int *n_out) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    if (!d || !out || !n_out || k < 1 || k > 64) return fail(ctx, SP_ERR_INVALID, "sp_comm_pair_minsum_topk: bad argument");
    *n_out = 0;
    int64_t lo = 0, hi = 0;
    sp_triangle_rows(d->np, c->world, c->rank, &lo, &hi);
    std::vector<sp_pair_rec> mine(static_cast<size_t>(k));
    int n_mine = 0;
    sp_status st = sp_pair_minsum_topk(ctx, d, d2, lo, hi, k, mine.data(), &n_mine);
    if (st != SP_OK) return st;
    if (c->world == 1) {
        std::copy(mine.begin(), mine.begin() + n_mine, out);
        *n_out = n_mine;
        return SP_OK;
    }
    for (int q = n_mine; q < k; ++q) { mine[static_cast<size_t>(q)] = sp_pair_rec{~0ull, ~0ull, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0}; }
    const size_t bytes = static_cast<size_t>(k) * sizeof(sp_pair_rec);
    char *d_buf = nullptr;
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    SP_CUDA(ctx, dev_malloc(ctx, &d_buf, bytes * static_cast<size_t>(c->world)));
    std::vector<sp_pair_rec> all(static_cast<size_t>(k) * static_cast<size_t>(c->world));
    cudaError_t e = cudaMemcpyAsync(d_buf + bytes * static_cast<size_t>(c->rank), mine.data(), bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && g_nccl.AllGather(d_buf + bytes * static_cast<size_t>(c->rank), d_buf, bytes, ncclUint8, c->comm, ctx->stream) != ncclSuccess)
        e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(all.data(), d_buf, bytes * static_cast<size_t>(c->world), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d_buf);
    if (e != cudaSuccess) return fail(ctx, SP_ERR_CUDA, "sp_comm_pair_minsum_topk: record exchange failed");
    all.erase(std::remove_if(all.begin(), all.end(), [](const sp_pair_rec &r) { return r.i == 0xFFFFFFFFu; }), all.end());
    std::sort(all.begin(), all.end(), [](const sp_pair_rec &a, const sp_pair_rec &b) {
        if (a.score != b.score) return a.score < b.score;
        if (a.score2 != b.score2) return a.score2 < b.score2;
        if (a.i != b.i) return a.i < b.i;
        return a.j < b.j;
    });
    const size_t kk = std::min<size_t>(static_cast<size_t>(k), all.size());
    std::copy(all.begin(), all.begin() + static_cast<std::ptrdiff_t>(kk), out);
    *n_out = static_cast<int>(kk);
    return SP_OK;
}

// Block until every rank of the communicator has reached this call (an all-gather of one byte per rank).
extern "C" sp_status sp_comm_barrier(sp_comm *c) {
    if (!c) return SP_ERR_INVALID;
    sp_ctx *ctx = c->ctx;
    SP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (c->world > 1) {
        char *d = nullptr;
        SP_CUDA(ctx, dev_malloc(ctx, &d, static_cast<size_t>(std::max(c->world, 16))));
        ncclResult_t r = g_nccl.AllGather(d + c->rank, d, 1, ncclUint8, c->comm, ctx->stream);
        dev_free(ctx, d);
        if (r != ncclSuccess) return fail(ctx, SP_ERR_CUDA, std::string("sp_comm_barrier: ") + g_nccl.GetErrorString(r));
    }
    SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SP_OK;
}
