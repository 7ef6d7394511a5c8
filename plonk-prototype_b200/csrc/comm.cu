// comm.cu — the multi-GPU communicator inside the library: one process per GPU, NCCL over NVLink / NVSwitch.
//
// MSM shards by point range (partial results all-gathered), large NTTs shard as a four-step
// decomposition with one all-to-all, the sharded prover uses both.  Round 1 drove those collectives through host-language
// callbacks (torch.distributed from Python); a Rust or C++ host got no multi-GPU path without re-writing that glue.  Here
// the library owns an ncclComm_t per context and enqueues every collective on the context's own stream, so a sharded
// prove has no host synchronisation around its exchanges and any host language can drive it:
//     rank 0:  pb200_comm_unique_id(id)  → broadcast the 128 bytes by any means (MPI, a file, torch's store …)
//     all:     pb200_comm_init(ctx, id, rank, world)
//              pb200_preprocess_comm(...) / pb200_prove(...)  |  pb200_msm_g1_sharded_dev  |  pb200_ntt_sharded_dev
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already loaded by the host process — e.g. PyTorch's — or
// the system one), so single-GPU users of libpb200.so need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return &api;
#define PB_SYM(field, name)                                  \
    api.field = (decltype(api.field))dlsym(h, name);         \
    if (!api.field) return &api;
    PB_SYM(GetUniqueId, "ncclGetUniqueId")
    PB_SYM(CommInitRank, "ncclCommInitRank")
    PB_SYM(CommDestroy, "ncclCommDestroy")
    PB_SYM(AllGather, "ncclAllGather")
    PB_SYM(Send, "ncclSend")
    PB_SYM(Recv, "ncclRecv")
    PB_SYM(GroupStart, "ncclGroupStart")
    PB_SYM(GroupEnd, "ncclGroupEnd")
    PB_SYM(GetErrorString, "ncclGetErrorString")
#undef PB_SYM
    api.ok = true;
    return &api;
}

}  // namespace

struct pb200_comm {
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
    void *small_dev = nullptr;   // staging for host-buffer all-gathers (≤ 64 KiB per rank)
    size_t small_bytes = 0;
};

#define PB_NCCL(ctx, call)                                                                                                   \
    do {                                                                                                                     \
        ncclResult_t r__ = (call);                                                                                           \
        if (r__ != ncclSuccess) return pb_fail(ctx, PB200_ERR_CUDA, #call, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

extern "C" int pb200_comm_unique_id(unsigned char id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (!id_out) return PB200_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->ok) return PB200_ERR_NO_DEVICE;   // no NCCL in this process: multi-GPU entry points are unavailable
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return PB200_ERR_CUDA;
    memcpy(id_out, &id, 128);
    return 0;
}
extern "C" int pb200_comm_init(pb200_ctx *ctx, const unsigned char id[128], uint32_t rank, uint32_t world) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, id != nullptr && world >= 1 && rank < world && ctx->comm == nullptr);
    NcclApi *api = nccl_api();
    if (!api->ok) return pb_fail(ctx, PB200_ERR_NO_DEVICE, "pb200_comm_init", "libnccl.so.2 could not be loaded", __FILE__, __LINE__);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    pb200_comm *c = new pb200_comm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = api->CommInitRank(&c->comm, (int)world, uid, (int)rank);
    if (r != ncclSuccess) {
        delete c;
        return pb_fail(ctx, PB200_ERR_CUDA, "ncclCommInitRank", api->GetErrorString(r), __FILE__, __LINE__);
    }
    c->small_bytes = (size_t)(1 + world) << 16;
    if (cudaMalloc(&c->small_dev, c->small_bytes) != cudaSuccess) {
        api->CommDestroy(c->comm);
        delete c;
        return pb_fail(ctx, PB200_ERR_CUDA, "pb200_comm_init", "staging allocation failed", __FILE__, __LINE__);
    }
    ctx->comm = c;
    return 0;
}
extern "C" int pb200_comm_destroy(pb200_ctx *ctx) {
    if (!ctx) return PB200_ERR_ARG;
    if (!ctx->comm) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_api()->CommDestroy(ctx->comm->comm);
    cudaFree(ctx->comm->small_dev);
    delete ctx->comm;
    ctx->comm = nullptr;
    return 0;
}
extern "C" int pb200_comm_info(const pb200_ctx *ctx, uint32_t *rank, uint32_t *world) {
    if (!ctx || !ctx->comm) return PB200_ERR_ARG;
    if (rank) *rank = ctx->comm->rank;
    if (world) *world = ctx->comm->world;
    return 0;
}

// recv_dev = world × bytes, rank-major; enqueued on the context stream
extern "C" int pb200_allgather_dev(pb200_ctx *ctx, const void *send_dev, void *recv_dev, size_t bytes) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, ctx->comm != nullptr && send_dev != nullptr && recv_dev != nullptr);
    if (bytes == 0) return 0;
    PB_NCCL(ctx, nccl_api()->AllGather(send_dev, recv_dev, bytes, ncclUint8, ctx->comm->comm, ctx->stream));
    return 0;
}
// block h (bytes_per_peer bytes) of send_dev goes to rank h, block b of recv_dev comes from rank b; context stream
extern "C" int pb200_alltoall_dev(pb200_ctx *ctx, const void *send_dev, void *recv_dev, size_t bytes_per_peer) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, ctx->comm != nullptr && send_dev != nullptr && recv_dev != nullptr && send_dev != recv_dev);
    if (bytes_per_peer == 0) return 0;
    NcclApi *api = nccl_api();
    const pb200_comm *c = ctx->comm;
    PB_NCCL(ctx, api->GroupStart());
    for (uint32_t h = 0; h < c->world; h++) {
        ncclResult_t r = api->Send((const char *)send_dev + (size_t)h * bytes_per_peer, bytes_per_peer, ncclUint8, (int)h, c->comm, ctx->stream);
        if (r == ncclSuccess) r = api->Recv((char *)recv_dev + (size_t)h * bytes_per_peer, bytes_per_peer, ncclUint8, (int)h, c->comm, ctx->stream);
        if (r != ncclSuccess) {
            api->GroupEnd();
            return pb_fail(ctx, PB200_ERR_CUDA, "ncclSend/ncclRecv", api->GetErrorString(r), __FILE__, __LINE__);
        }
    }
    PB_NCCL(ctx, api->GroupEnd());
    return 0;
}
// Host-buffer all-gather (a few hundred bytes: partial commitments, IPC handles): staged through device memory, blocks.
int comm_allgather_host(pb200_ctx *ctx, const void *send, void *recv, size_t bytes) {
    PB_ARG(ctx, ctx->comm != nullptr && bytes <= ((size_t)1 << 16));
    pb200_comm *c = ctx->comm;
    char *s = (char *)c->small_dev, *r = s + ((size_t)1 << 16);
    PB_CUDA(ctx, cudaMemcpyAsync(s, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PB_NCCL(ctx, nccl_api()->AllGather(s, r, bytes, ncclUint8, c->comm, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(recv, r, bytes * c->world, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
// A cross-rank barrier IN STREAM ORDER (no host synchronisation): a one-word all-gather — every rank's kernels enqueued
// before it have finished before any rank's kernels enqueued after it start.
int comm_stream_barrier(pb200_ctx *ctx) {
    PB_ARG(ctx, ctx->comm != nullptr);
    pb200_comm *c = ctx->comm;
    char *s = (char *)c->small_dev + c->small_bytes - 4096;   // its own corner of the staging block
    PB_NCCL(ctx, nccl_api()->AllGather(s, s + 64, 4, ncclUint8, c->comm, ctx->stream));
    return 0;
}

// The sharded prover's commitments: every rank leaves its `batch` partial MSM results (36 words each) in the first 64 KiB of
// the staging block; they are all-gathered, added per batch element by one kernel and copied to the host once — no
// per-element round trips, no host collective.
uint32_t *comm_partials_buffer(pb200_ctx *ctx) { return ctx->comm ? (uint32_t *)ctx->comm->small_dev : nullptr; }
int comm_sum_partials(pb200_ctx *ctx, uint32_t batch, uint64_t *out_xyz_host) {
    PB_ARG(ctx, ctx->comm != nullptr && batch >= 1 && (size_t)batch * 144 <= ((size_t)1 << 15));
    pb200_comm *c = ctx->comm;
    char *s = (char *)c->small_dev, *r = s + ((size_t)1 << 16);
    uint32_t *sums = (uint32_t *)(s + ((size_t)1 << 15));   // upper half of the send block
    PB_NCCL(ctx, nccl_api()->AllGather(s, r, (size_t)batch * 144, ncclUint8, c->comm, ctx->stream));
    PB_TRY(tail_g1_sum_batch(ctx, (const uint32_t *)r, c->world, batch, sums));
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sums, (size_t)batch * 144, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_xyz_host, ctx->pinned, (size_t)batch * 144);
    return 0;
}

// ---- callbacks with the pb200_shard signatures, bound to a context's communicator (user = ctx) ------------------------------
static int cb_allgather(void *user, const void *send, void *recv, size_t bytes) {
    return comm_allgather_host((pb200_ctx *)user, send, recv, bytes);
}
static int cb_alltoall_dev(void *user, const void *send_dev, void *recv_dev, size_t bytes_per_peer) {
    return pb200_alltoall_dev((pb200_ctx *)user, send_dev, recv_dev, bytes_per_peer);
}
static int cb_allgather_dev(void *user, const void *send_dev, void *recv_dev, size_t bytes) {
    return pb200_allgather_dev((pb200_ctx *)user, send_dev, recv_dev, bytes);
}

extern "C" int pb200_preprocess_comm(pb200_ctx *ctx, const pb200_srs *srs_slice, const pb200_circuit *circuit, const uint8_t *transcript_label,
                                     size_t label_len, pb200_prover_key **out, uint8_t vk_commitments[15 * 48]) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, ctx->comm != nullptr);
    pb200_shard sh;
    sh.rank = ctx->comm->rank;
    sh.world = ctx->comm->world;
    sh.allgather = cb_allgather;
    sh.user = ctx;
    sh.alltoall_dev = cb_alltoall_dev;
    sh.allgather_dev = cb_allgather_dev;
    sh.flags = PB200_SHARD_STREAM_ORDERED;
    if (sh.world == 1) return pb200_preprocess(ctx, srs_slice, circuit, transcript_label, label_len, out, vk_commitments);
    return pb200_preprocess_sharded(ctx, srs_slice, circuit, transcript_label, label_len, &sh, out, vk_commitments);
}

// ---- sharded MSM: Σ over ranks of the local MSMs; every rank gets the total -----------------------------
extern "C" int pb200_msm_g1_sharded_dev(pb200_ctx *ctx, const pb200_srs *srs_slice, size_t offset, const uint64_t *scalars_mont_dev, size_t n,
                                        uint64_t out_xyz_mont[18]) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, ctx->comm != nullptr && out_xyz_mont != nullptr);
    uint64_t mine[18];
    PB_TRY(pb200_msm_g1_dev(ctx, srs_slice, offset, scalars_mont_dev, n, mine));
    const uint32_t world = ctx->comm->world;
    if (world == 1) {
        memcpy(out_xyz_mont, mine, sizeof(mine));
        return 0;
    }
    std::vector<uint64_t> all((size_t)world * 18);
    PB_TRY(comm_allgather_host(ctx, mine, all.data(), sizeof(mine)));   // 144 bytes per rank: the only collective of the MSM path
    return pb200_g1_sum(ctx, all.data(), world, out_xyz_mont);
}

// ---- sharded four-step NTT (the orchestration of plonk-prototype_b200/dist_ntt.py in C++) ---------------
// forward: column layout in `data` → row layout in `data`; inverse: row layout → column layout.  `tmp` is a second buffer of
// the same size (2^log_n / world scalars).  n = n1·m with n1 = 2^8 (or world if larger); everything on the context stream.
extern "C" int pb200_ntt_sharded_dev(pb200_ctx *ctx, uint64_t *data_dev, uint64_t *tmp_dev, uint32_t log_n, int inverse) {
    if (!ctx) return PB200_ERR_ARG;
    PB_ARG(ctx, ctx->comm != nullptr && data_dev != nullptr && tmp_dev != nullptr && data_dev != tmp_dev);
    const uint32_t world = ctx->comm->world, rank = ctx->comm->rank;
    PB_ARG(ctx, (world & (world - 1)) == 0);
    uint32_t log_g = 0;
    while ((1u << log_g) < world) log_g++;
    const uint32_t log_n1 = std::max(8u, log_g);
    PB_ARG(ctx, log_n < 32 && log_n >= log_n1 + log_g);
    const uint32_t log_m = log_n - log_n1, log_cl = log_m - log_g, log_rl = log_n1 - log_g;
    const uint32_t cl = 1u << log_cl, rl = 1u << log_rl, col0 = rank << log_cl;
    const size_t local = (size_t)1 << (log_n - log_g), peer_bytes = (local >> log_g) * 32;
    if (!inverse) {
        PB_TRY(pb200_ntt_columns_dev(ctx, data_dev, log_n, log_n1, log_cl, col0, 0));
        PB_TRY(pb200_alltoall_dev(ctx, data_dev, tmp_dev, peer_bytes));                  // block h (rl × cl) → rank h
        PB_TRY(pb200_block_transpose_dev(ctx, data_dev, tmp_dev, world, rl, cl));        // → rl rows of length m
        PB_TRY(pb200_ntt_batch_dev(ctx, data_dev, log_m, rl, 0, 0));
    } else {
        PB_TRY(pb200_ntt_batch_dev(ctx, data_dev, log_m, rl, 1, 0));
        PB_TRY(pb200_block_transpose_dev(ctx, tmp_dev, data_dev, rl, world, cl));
        PB_TRY(pb200_alltoall_dev(ctx, tmp_dev, data_dev, peer_bytes));
        PB_TRY(pb200_ntt_columns_dev(ctx, data_dev, log_n, log_n1, log_cl, col0, 1));
    }
    return 0;
}
