// fyx_comm.inl — multi-GPU exchange of the visible lists (included at the end of fyx_api.cu).
//
// One process per GPU, one fyx_ctx per process.  The node array is sharded by sub-tree
// so transforms / boxes / skinning never cross GPUs; the only exchange step of the path is the
// all-gather of the compacted visible-index lists, done with NCCL over NVLink 5 / NVSwitch:
//   1. ncclAllGather of the per-frustum counts (FYX_MAX_FRUSTA u32 per rank)
//   2. one grouped ncclAllGather per frustum of fixed max-count slots
//   3. a pack kernel that removes the slot padding (rank order)
// The reference has no counterpart (it is single-process).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): in a Python host that already imported torch
// this resolves to torch's bundled NCCL, otherwise to the system library; the C ABI itself has no
// link-time NCCL dependency, so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
#define SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd &&
             api.GetErrorString;
    return api;
}

#define NC(call)                                                                                            \
    do {                                                                                                    \
        ncclResult_t r__ = (call);                                                                          \
        if (r__ != ncclSuccess) return fail(c, FYX_ERR_NCCL, "%s failed: %s", #call, nccl().GetErrorString(r__)); \
    } while (0)

} // namespace

static void fyx_comm_destroy_internal(fyx_ctx *c)
{
    if (c->comm_stream) {
        cudaStreamSynchronize(c->comm_stream);
        cudaStreamDestroy(c->comm_stream);
        c->comm_stream = nullptr;
    }

    if (c->comm && nccl().ok) nccl().CommDestroy(static_cast<ncclComm_t>(c->comm));
    c->comm = nullptr;
}

extern "C" int32_t fyx_comm_get_unique_id(void *out_id128)
{
    if (!out_id128) return FYX_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(ncclUniqueId) == FYX_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!nccl().ok) return fail(nullptr, FYX_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, FYX_ERR_NCCL, "ncclGetUniqueId failed: %s", nccl().GetErrorString(r));
    memcpy(out_id128, &id, sizeof id);
    return FYX_OK;
}

extern "C" int32_t fyx_comm_init(fyx_ctx *c, int32_t nranks, int32_t rank, const void *id128)
{
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return FYX_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return fail(c, FYX_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CU(cudaSetDevice(c->device));
    fyx_comm_destroy_internal(c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    NC(nccl().CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->nranks = nranks;
    c->rank = rank;
    int32_t rc;
    if ((rc = dev_ensure(c, c->b_counts_packed, sizeof(uint32_t) * FYX_MAX_FRUSTA))) return rc;
    if ((rc = dev_ensure(c, c->b_counts_all, sizeof(uint32_t) * FYX_MAX_FRUSTA * nranks))) return rc;
    if (c->h_counts_all) cudaFreeHost(c->h_counts_all);
    CU(cudaHostAlloc(reinterpret_cast<void **>(&c->h_counts_all), sizeof(uint32_t) * FYX_MAX_FRUSTA * nranks, cudaHostAllocDefault));
    if (!c->comm_stream) {
        // highest priority: its few CTAs are placed as soon as a skinning CTA retires, so the exchange really
        // runs beside k_skin instead of after it (k_skin alone fills the register file of every SM)
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    return FYX_OK;
}

// Step 1 of the exchange, enqueued on stream `s`: pack the counters, all-gather them, start their copy to the host.
static int32_t allgather_begin(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    // a new exchange may only start once the previous one's pack kernels have read the shared count table
    // (same stream in every caller ⇒ stream order guarantees it)
    CU(cudaMemsetAsync(c->b_counts_packed.p, 0, sizeof(uint32_t) * FYX_MAX_FRUSTA, s));
    CU(cudaMemcpy2DAsync(c->b_counts_packed.p, sizeof(uint32_t), V.d_counts, sizeof(uint32_t) * kCountStride, sizeof(uint32_t), V.nf,
                         cudaMemcpyDeviceToDevice, s));
    NC(nccl().AllGather(c->b_counts_packed.p, c->b_counts_all.p, FYX_MAX_FRUSTA, ncclUint32, comm, s));
    CU(cudaMemcpyAsync(c->h_counts_all, c->b_counts_all.p, sizeof(uint32_t) * FYX_MAX_FRUSTA * c->nranks, cudaMemcpyDeviceToHost, s));
    return FYX_OK;
}

// Steps 2+3: wait (host) for the counts, then enqueue the payload all-gathers in max-count slots and the pack kernels.
static int32_t allgather_finish(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    const uint32_t nf = V.nf;
    const int R = c->nranks;
    CU(cudaStreamSynchronize(s));
    for (uint32_t f = 0; f < nf; ++f) V.h_counts[f] = c->h_counts_all[c->rank * FYX_MAX_FRUSTA + f];
    V.counts_on_host = true; // own counts are now known on the host as well
    uint32_t maxc[FYX_MAX_FRUSTA] = {};
    int32_t rc;
    for (uint32_t f = 0; f < nf; ++f) {
        uint64_t total = 0;
        for (int r = 0; r < R; ++r) {
            const uint32_t n = c->h_counts_all[r * FYX_MAX_FRUSTA + f];
            maxc[f] = std::max(maxc[f], n);
            total += n;
        }
        V.gath_count[f] = (uint32_t)total;
        if ((rc = dev_ensure(c, V.b_gath_pad[f], sizeof(uint32_t) * std::max<size_t>((size_t)maxc[f] * R, 1)))) return rc;
        if ((rc = dev_ensure(c, V.b_gath[f], sizeof(uint32_t) * std::max<size_t>(total, 1)))) return rc;
        // the send buffer must hold maxc entries: visible lists are sized for every renderable node of THIS
        // shard, which may be fewer than another rank's count (rare: grow it with everything quiesced)
        if ((size_t)maxc[f] * sizeof(uint32_t) > V.b_vis[f].bytes) {
            CU(cudaDeviceSynchronize());
            if ((rc = dev_ensure(c, V.b_vis[f], sizeof(uint32_t) * (size_t)maxc[f], true))) return rc;
            CU(cudaDeviceSynchronize());
            c->cp.out[f] = V.b_vis[f].as<uint32_t>();
        }
    }
    NC(nccl().GroupStart());
    for (uint32_t f = 0; f < nf; ++f)
        if (maxc[f]) NC(nccl().AllGather(V.b_vis[f].p, V.b_gath_pad[f].p, maxc[f], ncclUint32, comm, s));
    NC(nccl().GroupEnd());
    for (uint32_t f = 0; f < nf; ++f) {
        launch_compact_gathered(s, V.b_gath_pad[f].as<uint32_t>(), maxc[f], c->b_counts_all.as<uint32_t>(), R, (int)f,
                                V.b_gath[f].as<uint32_t>());
        c->launches += maxc[f] ? 1 : 0;
    }
    CU(cudaGetLastError());
    V.gathered = true;
    V.gathered_on_host = false;
    return FYX_OK;
}

extern "C" int32_t fyx_allgather_visible(fyx_ctx *c)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->comm) return fail(c, FYX_ERR_STATE, "fyx_comm_init has not been called");
    VisSlot &V = c->vs[c->cur];
    if (!V.nf) return FYX_OK;
    CU(cudaSetDevice(c->device));
    // order after any in-frame exchange of the other slot (it ran on the collective stream)
    if (c->comm_stream && c->vs[c->cur ^ 1].gathered) CU(cudaStreamWaitEvent(c->stream, c->vs[c->cur ^ 1].ev_gather, 0));
    int32_t rc = allgather_begin(c, V, c->stream);
    if (rc) return rc;
    rc = allgather_finish(c, V, c->stream);
    if (rc) return rc;
    CU(cudaEventRecord(V.ev_gather, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_get_visible_gathered_device(fyx_ctx *c, uint32_t f, const uint32_t **d_idx, uint32_t *out_count)
{
    if (!c || !d_idx || !out_count) return FYX_ERR_INVALID_ARGUMENT;
    VisSlot &V = c->vs[c->cur];
    if (f >= V.nf || !V.gathered) return fail(c, FYX_ERR_INVALID_ARGUMENT, "frustum %u has no gathered list", f);
    *d_idx = V.b_gath[f].as<uint32_t>();
    *out_count = V.gath_count[f];
    return FYX_OK;
}

extern "C" int32_t fyx_get_visible_gathered(fyx_ctx *c, uint32_t f, const uint32_t **out_idx, uint32_t *out_count)
{
    if (!c || !out_idx || !out_count) return FYX_ERR_INVALID_ARGUMENT;
    VisSlot &V = c->vs[c->readable];
    if (V.pending) return fail(c, FYX_ERR_STATE, "the frame is still in flight: call fyx_frame_wait first");
    if (f >= V.nf || !V.gathered) return fail(c, FYX_ERR_INVALID_ARGUMENT, "frustum %u has no gathered list", f);
    CU(cudaSetDevice(c->device));
    const size_t n = V.gath_count[f];
    if (!V.gathered_on_host) {
        // bring all frusta at once (one synchronisation), after the collective stream has produced them
        CU(cudaStreamWaitEvent(c->stream, V.ev_gather, 0));
        for (uint32_t g = 0; g < V.nf; ++g) {
            const size_t m = V.gath_count[g];
            int32_t rc = host_gath_ensure(c, V, g, m);
            if (rc) return rc;
            if (m) CU(cudaMemcpyAsync(V.h_gath[g], V.b_gath[g].p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
        V.gathered_on_host = true;
    }
    *out_idx = V.h_gath[f];
    *out_count = (uint32_t)n;
    return FYX_OK;
}
