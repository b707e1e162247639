// S1 across GPUs inside the C-ABI (SURVEY §8(b) kv_all_to_all): DefaultShuffler.shuffle (base.py:416-433) moves
// every map output to the reducer that owns its partition through files and a Queue; here one process per GPU
// holds an NCCL communicator and ONE call
//   1. splits the records by owner = mix64(key) % world into destination-contiguous runs (the stable partition
//      level of kv.cu: a key-sorted input arrives as one SORTED RUN per source rank),
//   2. all-gathers the row of counts together with a small caller header (line counts, flags ...: what used to be
//      separate all-reduces) — the only host synchronisation of the exchange,
//   3. moves the payload with grouped ncclSend / ncclRecv straight between the kv buffers (NVLink 5 / NVSwitch).
// NCCL is bound at run time (dlopen of libnccl.so.2: in a process that imported torch that is torch's own copy),
// so the library loads, and everything single-GPU works, on a machine without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <vector>

#include "common.cuh"

struct dampr_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    dampr_ctx *ctx = nullptr;
    u64 *d_row = nullptr;   // device: this rank's row (world counts + header)
    u64 *d_all = nullptr;   // device: world rows
    size_t row_cap = 0;     // u64 per row the buffers were sized for
};

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string why;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.why = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed");
            return;
        }
#define BIND(field, sym)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));     \
    if (!api.field) {                                                           \
        api.why = std::string("NCCL symbol missing: ") + sym;                   \
        api.lib = nullptr;                                                      \
        return;                                                                 \
    }
        BIND(GetUniqueId, "ncclGetUniqueId")
        BIND(CommInitRank, "ncclCommInitRank")
        BIND(CommDestroy, "ncclCommDestroy")
        BIND(AllGather, "ncclAllGather")
        BIND(Send, "ncclSend")
        BIND(Recv, "ncclRecv")
        BIND(GroupStart, "ncclGroupStart")
        BIND(GroupEnd, "ncclGroupEnd")
        BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    });
    return api.lib ? &api : nullptr;
}

int nccl_fail(dampr_ctx *ctx, ncclResult_t r, const char *what) {
    NcclApi *a = nccl_api();
    char b[384];
    snprintf(b, sizeof b, "%s failed: %s", what, a && a->GetErrorString ? a->GetErrorString(r) : "?");
    if (ctx) ctx->err = b;
    return DAMPR_ERR_CUDA;
}

#define NCCL_TRY(ctx, expr)                                  \
    do {                                                     \
        ncclResult_t _r = (expr);                            \
        if (_r != ncclSuccess) return nccl_fail(ctx, _r, #expr); \
    } while (0)

}  // namespace

extern "C" {

int32_t dampr_comm_unique_id(uint8_t *out_id, uint32_t cap) {
    if (!out_id || cap < sizeof(ncclUniqueId)) return DAMPR_ERR_ARG;
    NcclApi *a = nccl_api();
    if (!a) return DAMPR_ERR_CUDA;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return DAMPR_ERR_CUDA;
    memcpy(out_id, &id, sizeof id);
    return DAMPR_OK;
}

int32_t dampr_comm_create(dampr_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id, uint32_t id_bytes,
                          dampr_comm **out) {
    ARG_CHECK(ctx, ctx && id && out, "null");
    ARG_CHECK(ctx, world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    ARG_CHECK(ctx, id_bytes >= sizeof(ncclUniqueId), "unique id too short");
    NcclApi *a = nccl_api();
    if (!a) {
        ctx->err = "NCCL is not available (dlopen libnccl.so.2)";
        return DAMPR_ERR_CUDA;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId nid;
    memcpy(&nid, id, sizeof nid);
    dampr_comm *c = new dampr_comm();
    c->rank = rank;
    c->world = world;
    c->ctx = ctx;
    ncclResult_t r = a->CommInitRank(&c->comm, world, nid, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(ctx, r, "ncclCommInitRank");
    }
    *out = c;
    return DAMPR_OK;
}

int32_t dampr_comm_destroy(dampr_comm *c) {
    if (!c) return DAMPR_OK;
    NcclApi *a = nccl_api();
    if (c->ctx) cudaSetDevice(c->ctx->device);
    if (c->d_row) cudaFree(c->d_row);
    if (c->d_all) cudaFree(c->d_all);
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
    return DAMPR_OK;
}

// Every record of `kv` to the owner of its key. On return *out holds the records this rank owns, the records
// of source rank s at [out_offsets[s], out_offsets[s + 1]) in the order rank s held them; headers_all (world x
// n_header, may be null when n_header is 0) holds every rank's header row. `kv` itself is left untouched.
int32_t dampr_kv_all_to_all(dampr_ctx *ctx, dampr_comm *c, dampr_kv *kv, const int64_t *header, int32_t n_header,
                            dampr_kv **out, uint64_t *out_offsets, int64_t *headers_all) {
    ARG_CHECK(ctx, ctx && c && kv && out && out_offsets, "null");
    ARG_CHECK(ctx, c->ctx == ctx, "communicator belongs to another context");
    ARG_CHECK(ctx, n_header >= 0 && n_header <= 64 && (n_header == 0 || (header && headers_all)), "bad header");
    NcclApi *a = nccl_api();
    if (!a) {
        ctx->err = "NCCL is not available";
        return DAMPR_ERR_CUDA;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int W = c->world;
    const size_t row = (size_t)W + (size_t)n_header;
    if (c->row_cap < row) {
        if (c->d_row) cudaFree(c->d_row);
        if (c->d_all) cudaFree(c->d_all);
        c->d_row = c->d_all = nullptr;
        c->row_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&c->d_row, row * 8));
        CUDA_TRY(ctx, cudaMalloc(&c->d_all, row * 8 * W));
        c->row_cap = row;
    }
    // 1. destination-contiguous send runs
    dampr_kv *parts = nullptr;
    std::vector<uint64_t> counts((size_t)W);
    int rc = dampr_kv_partition_by_owner(ctx, kv, W, &parts, counts.data());
    if (rc) return rc;
    struct PartsGuard {
        dampr_ctx *ctx;
        dampr_kv *p;
        ~PartsGuard() { if (p) dampr_kv_destroy(ctx, p); }
    } guard{ctx, parts};
    // 2. counts + header of every rank (one all-gather, one host synchronisation)
    u64 *h_row = (u64 *)host_pin(ctx, 0, row * 8 * (size_t)(W + 1));
    if (!h_row) return DAMPR_ERR_NOMEM;
    u64 *h_all = h_row + row;
    for (int d = 0; d < W; ++d) h_row[d] = counts[(size_t)d];
    for (int i = 0; i < n_header; ++i) h_row[W + i] = (u64)header[i];
    CUDA_TRY(ctx, cudaMemcpyAsync(c->d_row, h_row, row * 8, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, a->AllGather(c->d_row, c->d_all, row, ncclUint64, c->comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_all, c->d_all, row * 8 * W, cudaMemcpyDeviceToHost, ctx->stream));
    host_pin_used(ctx, 0);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<u64> recv((size_t)W), send_off((size_t)W + 1, 0);
    u64 total = 0;
    out_offsets[0] = 0;
    for (int s = 0; s < W; ++s) {
        recv[(size_t)s] = h_all[(size_t)s * row + (size_t)c->rank];
        total += recv[(size_t)s];
        out_offsets[s + 1] = total;
        send_off[(size_t)s + 1] = send_off[(size_t)s] + counts[(size_t)s];
        for (int i = 0; i < n_header; ++i) headers_all[(size_t)s * n_header + i] = (int64_t)h_all[(size_t)s * row + W + i];
    }
    // 3. payload
    dampr_kv *dst = nullptr;
    rc = dampr_kv_create(ctx, total ? total : 1, &dst);
    if (rc) return rc;
    dst->n = total;
    ncclResult_t r = a->GroupStart();
    if (r == ncclSuccess) {
        for (int p = 0; p < W && r == ncclSuccess; ++p) {
            if (counts[(size_t)p])
                r = a->Send(parts->rec + send_off[(size_t)p], counts[(size_t)p] * 2, ncclUint64, p, c->comm, ctx->stream);
            if (r == ncclSuccess && recv[(size_t)p])
                r = a->Recv(dst->rec + out_offsets[p], recv[(size_t)p] * 2, ncclUint64, p, c->comm, ctx->stream);
        }
        ncclResult_t r2 = a->GroupEnd();
        if (r == ncclSuccess) r = r2;
    }
    if (r != ncclSuccess) {
        dampr_kv_destroy(ctx, dst);
        return nccl_fail(ctx, r, "grouped ncclSend / ncclRecv");
    }
    // `parts` is released when the guard runs: the pool hands the block out again only to later work on this
    // stream, which is ordered after the sends
    *out = dst;
    return DAMPR_OK;
}

}  // extern "C"
