// fyx_comm.inl — multi-GPU exchange of the visible lists (included at the end of fyx_api.cu).
//
// One process per GPU, one fyx_ctx per process.  The node array is sharded by sub-tree
// so transforms / boxes / skinning never cross GPUs; the only exchange step of the path is the
// all-gather of the compacted visible-index lists over NVLink 5 / NVSwitch.  Two device-side forms:
//
//   peer  (default when cudaIpc mapping works on every rank): the ranks store their lists straight into
//         each other's memory at their final offsets — fyx_peer.cu; no host synchronisation, no padding;
//   nccl  (FYX_EXCHANGE=nccl, or the fallback): ncclAllGather of the counts, host wait for them, one grouped
//         ncclAllGather per frustum of fixed max-count slots, a pack kernel per frustum.
//
// and, independent of that choice, the HOST copy of the gathered lists: every rank DMA-copies its OWN lists
// into one node-wide host segment (fyx_hostseg.hpp) at its offsets — N PCIe links in parallel — instead of
// one rank pulling all N parts over its single link (FYX_HOSTSEG=0 restores per-rank copies of the whole
// gathered device list).  NCCL stays the bootstrap channel of both (counts of slots, cudaIpc handles, the
// segment's descriptor travel through ncclAllGather).  The reference has no counterpart (single process).
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

// all-gather `bytes` (a multiple of 4) of host data per rank through NCCL: the bootstrap channel of the exchange set-up
int32_t bootstrap_allgather(fyx_ctx *c, const void *mine, void *all, size_t bytes)
{
    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    DevBuf tmp;
    int32_t rc = dev_ensure(c, tmp, bytes * (size_t)(c->nranks + 1));
    if (rc) return rc;
    char *d = tmp.as<char>();
    CU(cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, c->stream));
    NC(nccl().AllGather(d, d + bytes, bytes / 4, ncclUint32, comm, c->stream));
    CU(cudaMemcpyAsync(all, d + bytes, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dev_free(tmp);
    return FYX_OK;
}

void exchange_teardown(fyx_ctx *c)
{
    PeerState &P = c->peer;
    for (int r = 0; r < kPeerMaxRanks; ++r) {
        if (P.mapped[r]) cudaIpcCloseMemHandle(P.mapped[r]);
        P.mapped[r] = nullptr;
        P.pp.base[r] = nullptr;
    }
    if (P.local) cudaFree(P.local);
    P.local = nullptr;
    if (P.cta_done) cudaFree(P.cta_done);
    P.cta_done = nullptr;
    if (P.h_counts) cudaFreeHost(P.h_counts);
    P.h_counts = nullptr;
    P.ready = false;
    if (c->hostseg.base && c->hostseg_registered) cudaHostUnregister(c->hostseg.base);
    c->hostseg_registered = false;
    c->hostseg.close();
    c->hostseg_ready = false;
    c->exchange_built = false;
}

// Collective: build the peer mapping and the host segment for lists of up to `nf` frusta.  Every rank calls it at the
// same point (the first gathered frame after fyx_comm_init / fyx_set_topology / a larger frustum count).  Failures on
// any rank switch the feature off on EVERY rank (the outcome is all-gathered), never half-way.
int32_t exchange_setup(fyx_ctx *c, uint32_t nf)
{
    exchange_teardown(c);
    const int R = c->nranks;
    int32_t rc;
    // 1. list capacity = every rank's slots together (worst case: everything visible)
    uint32_t mine[2] = {c->n_slots, nf}, all[2 * kPeerMaxRanks];
    if ((rc = bootstrap_allgather(c, mine, all, sizeof mine))) return rc;
    uint64_t total = 0;
    uint32_t nf_cap = nf;
    for (int r = 0; r < R; ++r) {
        total += all[2 * r];
        nf_cap = std::max(nf_cap, all[2 * r + 1]);
    }
    total = (total + 63) & ~uint64_t(63); // 256-byte aligned regions
    c->exch_total_cap = total;
    c->exch_nf_cap = nf_cap;

    // 2. peer mapping
    PeerState &P = c->peer;
    uint32_t ok = (c->want_peer && R <= kPeerMaxRanks) ? 1u : 0u;
    cudaIpcMemHandle_t hmine;
    memset(&hmine, 0, sizeof hmine);
    if (ok) {
        const size_t bytes = kPeerCtrlBytes + (size_t)2 * nf_cap * total * sizeof(uint32_t);
        if (cudaMalloc(&P.local, bytes) != cudaSuccess || cudaMemset(P.local, 0, kPeerCtrlBytes) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void **>(&P.cta_done), 256) != cudaSuccess || cudaMemset(P.cta_done, 0, 256) != cudaSuccess ||
            cudaHostAlloc(reinterpret_cast<void **>(&P.h_counts), sizeof(uint32_t) * 2 * 2 * kPeerMaxRanks * FYX_MAX_FRUSTA, cudaHostAllocDefault) != cudaSuccess ||
            cudaIpcGetMemHandle(&hmine, P.local) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    struct Card { uint32_t ok; uint32_t pad; unsigned char handle[64]; };
    Card cm, call[kPeerMaxRanks];
    cm.ok = ok;
    cm.pad = 0;
    memcpy(cm.handle, &hmine, 64);
    if ((rc = bootstrap_allgather(c, &cm, call, sizeof cm))) return rc;
    for (int r = 0; r < R; ++r) ok &= call[r].ok;
    if (ok) {
        for (int r = 0; r < R && ok; ++r) {
            if (r == c->rank) {
                P.pp.base[r] = static_cast<unsigned char *>(P.local);
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, call[r].handle, 64);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            P.mapped[r] = p;
            P.pp.base[r] = static_cast<unsigned char *>(p);
        }
    }
    // everybody must have mapped everybody
    uint32_t okall[kPeerMaxRanks];
    if ((rc = bootstrap_allgather(c, &ok, okall, sizeof ok))) return rc;
    for (int r = 0; r < R; ++r) ok &= okall[r];
    if (ok) {
        P.pp.nranks = R;
        P.pp.rank = c->rank;
        P.pp.total_cap = total;
        P.pp.nf_cap = nf_cap;
        P.pp.cta_done = P.cta_done;
        P.pp.d_err = c->d_err;
        P.ready = true;
    } else {
        for (int r = 0; r < kPeerMaxRanks; ++r) {
            if (P.mapped[r]) cudaIpcCloseMemHandle(P.mapped[r]);
            P.mapped[r] = nullptr;
        }
        if (P.local) cudaFree(P.local);
        P.local = nullptr;
        P.ready = false;
    }

    // 3. host segment (rank 0 creates, the others open its descriptor)
    uint32_t hs_ok = c->want_hostseg ? 1u : 0u;
    int64_t pid_fd[2] = {0, 0};
    if (hs_ok && c->rank == 0 && !c->hostseg.create(total, nf_cap, R, 0, pid_fd)) hs_ok = 0;
    struct Desc { int64_t pid_fd[2]; uint32_t ok, pad; } dm, dall[kPeerMaxRanks];
    dm.pid_fd[0] = pid_fd[0];
    dm.pid_fd[1] = pid_fd[1];
    dm.ok = hs_ok;
    dm.pad = 0;
    if ((rc = bootstrap_allgather(c, &dm, dall, sizeof dm))) return rc;
    hs_ok = dall[0].ok && c->want_hostseg;
    if (hs_ok && c->rank != 0 && !c->hostseg.open_from(dall[0].pid_fd, total, nf_cap, R, c->rank)) hs_ok = 0;
    if (hs_ok) {
        // page-lock the mapping so that the D2H copies into it are plain DMA (not fatal if the driver refuses: the copies
        // then go through its staging buffer)
        c->hostseg_registered = cudaHostRegister(c->hostseg.base, c->hostseg.bytes, cudaHostRegisterPortable) == cudaSuccess;
        if (!c->hostseg_registered) cudaGetLastError();
    }
    uint32_t hsall[kPeerMaxRanks];
    if ((rc = bootstrap_allgather(c, &hs_ok, hsall, sizeof hs_ok))) return rc;
    for (int r = 0; r < R; ++r) hs_ok &= hsall[r];
    if (!hs_ok) {
        if (c->hostseg.base && c->hostseg_registered) cudaHostUnregister(c->hostseg.base);
        c->hostseg_registered = false;
        c->hostseg.close();
    }
    c->hostseg_ready = hs_ok != 0;
    c->exchange_built = true;
    c->exchange_slots = c->n_slots;
    return FYX_OK;
}

int32_t exchange_ensure(fyx_ctx *c, uint32_t nf)
{
    if (c->exchange_built && nf <= c->exch_nf_cap && c->n_slots <= c->exchange_slots) return FYX_OK;
    // (re)building is COLLECTIVE: every rank gets here in the same gathered frame — true for the first one and for a
    // larger frustum count; a shard that outgrew the size it advertised must call fyx_comm_init again on every rank
    if (c->exchange_built && c->n_slots > c->exchange_slots)
        return fail(c, FYX_ERR_STATE, "this rank's shard grew from %u to %u nodes after the exchange was built: call fyx_comm_init again on every rank",
                    c->exchange_slots, c->n_slots);
    CU(cudaDeviceSynchronize()); // nothing of an older mapping may still be in flight
    return exchange_setup(c, nf);
}

} // namespace

static void fyx_comm_destroy_internal(fyx_ctx *c)
{
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    exchange_teardown(c);
    if (c->comm_stream) {
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
    c->gather_epoch = 0;
    // FYX_EXCHANGE=nccl keeps the collective on NCCL; FYX_HOSTSEG=0 keeps per-rank host copies of the gathered lists
    const char *ex = getenv("FYX_EXCHANGE"), *hs = getenv("FYX_HOSTSEG");
    c->want_peer = !(ex && strcmp(ex, "nccl") == 0);
    c->want_hostseg = !(hs && strcmp(hs, "0") == 0);
    const char *pc = getenv("FYX_PEER_CTAS");
    c->peer_push_ctas = (pc && atoi(pc) > 0) ? (unsigned)atoi(pc) : 96u;
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

extern "C" uint32_t fyx_comm_mode(const fyx_ctx *c)
{
    if (!c || !c->comm) return 0u;
    return FYX_COMM_NCCL | (c->exchange_built && c->peer.ready ? FYX_COMM_PEER_STORES : 0u) | (c->exchange_built && c->hostseg_ready ? FYX_COMM_HOST_SEGMENT : 0u) |
           (c->exchange_built ? 0u : FYX_COMM_UNDECIDED);
}

// ---- NCCL form -------------------------------------------------------------------------------------------------
// Step 1, enqueued on stream `s`: pack the counters, all-gather them, start their copy to the host.
static int32_t nccl_allgather_begin(fyx_ctx *c, VisSlot &V, cudaStream_t s)
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
static int32_t nccl_allgather_finish(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    const uint32_t nf = V.nf;
    const int R = c->nranks;
    CU(cudaStreamSynchronize(s));
    for (int r = 0; r < R; ++r)
        for (uint32_t f = 0; f < nf; ++f) V.counts_all[r][f] = c->h_counts_all[r * FYX_MAX_FRUSTA + f];
    V.counts_all_known = true;
    for (uint32_t f = 0; f < nf; ++f) V.h_counts[f] = V.counts_all[c->rank][f];
    V.counts_on_host = true; // own counts are now known on the host as well
    uint32_t maxc[FYX_MAX_FRUSTA] = {};
    int32_t rc;
    for (uint32_t f = 0; f < nf; ++f) {
        uint64_t total = 0;
        for (int r = 0; r < R; ++r) {
            const uint32_t n = V.counts_all[r][f];
            maxc[f] = std::max(maxc[f], n);
            total += n;
        }
        V.gath_count[f] = (uint32_t)total;
        if ((rc = dev_ensure(c, V.b_gath_pad[f], sizeof(uint32_t) * std::max<size_t>((size_t)maxc[f] * R, 1)))) return rc;
        if ((rc = dev_ensure(c, V.b_gath[f], sizeof(uint32_t) * std::max<size_t>(total, 1)))) return rc;
        V.gath_ptr[f] = V.b_gath[f].as<uint32_t>();
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
    return FYX_OK;
}

// ---- the frame's exchange, either form ---------------------------------------------------------------------------
// Called once per gathered frame, after the cull has been enqueued (V.ev_cull recorded on the frame's stream when `s`
// is another stream).  On return the exchange is enqueued on `s`; V.ev_gather is recorded by the caller.
static int32_t allgather_enqueue(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    int32_t rc = exchange_ensure(c, V.nf);
    if (rc) return rc;
    V.epoch = ++c->gather_epoch;
    V.gathered = true;
    V.gathered_on_host = false;
    V.seg_published = false;
    V.counts_all_known = false;
    V.via_peer = c->peer.ready;
    if (c->hostseg_ready) c->hostseg.begin(V.epoch); // this rank no longer reads the host lists of epoch - 2
    CU(cudaEventRecord(c->ev_x0[V.epoch & 1], s));
    c->x_epoch[V.epoch & 1] = V.epoch;
    c->x_slot[V.epoch & 1] = (int)(&V - c->vs);
    if (c->gath_read_valid[V.epoch & 1]) { // a private D2H copy of the list buffers this epoch reuses must have drained
        CU(cudaStreamWaitEvent(s, c->ev_gath_read[V.epoch & 1], 0));
        c->gath_read_valid[V.epoch & 1] = false;
    }
    if (!V.via_peer) {
        if ((rc = nccl_allgather_begin(c, V, s))) return rc;
        return FYX_OK; // nccl_allgather_finish follows once the rest of the frame is enqueued (it waits on the host)
    }
    PeerState &P = c->peer;
    PeerParams pp = P.pp;
    pp.nf = (int)V.nf;
    pp.epoch = (uint32_t)V.epoch;
    pp.own_counts = V.d_counts;
    for (uint32_t f = 0; f < V.nf; ++f) {
        pp.own_list[f] = V.b_vis[f].as<uint32_t>();
        V.gath_ptr[f] = peer_list(pp, c->rank, (uint32_t)(V.epoch & 1), f);
    }
    launch_peer_counts(s, pp);
    // the counts of every rank are on this device now: start their (small) copy to the host right away — the host segment
    // needs the offsets, the getters the totals — while the lists are still travelling
    uint32_t *h = P.h_counts + (size_t)(V.epoch & 1) * 2 * kPeerMaxRanks * FYX_MAX_FRUSTA;
    const PeerCtrl *ctrl = peer_ctrl(pp, c->rank);
    CU(cudaMemcpyAsync(h, ctrl->counts[V.epoch & 1], sizeof(uint32_t) * kPeerMaxRanks * FYX_MAX_FRUSTA, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(V.ev_counts_all, s));
    // enough CTAs to keep the NVLink egress busy, few enough to slip in beside the skinning kernel
    launch_peer_push(s, pp, c->peer_push_ctas);
    c->launches += 3;
    CU(cudaGetLastError());
    return FYX_OK;
}

static int32_t allgather_finish(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    if (V.via_peer) return FYX_OK;
    return nccl_allgather_finish(c, V, s);
}

// counts of every rank for this frame on the host (peer form: after the count kernel; NCCL form: already there)
static int32_t resolve_counts(fyx_ctx *c, VisSlot &V)
{
    if (V.counts_all_known) return FYX_OK;
    CU(cudaEventSynchronize(V.ev_counts_all));
    const uint32_t *h = c->peer.h_counts + (size_t)(V.epoch & 1) * 2 * kPeerMaxRanks * FYX_MAX_FRUSTA;
    for (int r = 0; r < c->nranks; ++r)
        for (uint32_t f = 0; f < V.nf; ++f) V.counts_all[r][f] = h[r * FYX_MAX_FRUSTA + f];
    for (uint32_t f = 0; f < V.nf; ++f) {
        uint64_t total = 0;
        for (int r = 0; r < c->nranks; ++r) total += V.counts_all[r][f];
        V.gath_count[f] = (uint32_t)total;
        V.h_counts[f] = V.counts_all[c->rank][f];
    }
    V.counts_on_host = true;
    V.counts_all_known = true;
    return FYX_OK;
}

// This rank's own lists into the node-wide host segment at their offsets (stream `s`, which must already be ordered
// after the cull), then the publication the readers wait for.  Synchronises `s`.
static int32_t hostseg_publish(fyx_ctx *c, VisSlot &V, cudaStream_t s)
{
    if (V.seg_published) return FYX_OK;
    int32_t rc = resolve_counts(c, V);
    if (rc) return rc;
    HostSeg &H = c->hostseg;
    if (!H.wait_writable(V.epoch)) return fail(c, FYX_ERR_NCCL, "%s", H.err.c_str());
    for (uint32_t f = 0; f < V.nf; ++f) {
        uint64_t off = 0;
        for (int q = 0; q < c->rank; ++q) off += V.counts_all[q][f];
        const size_t n = V.counts_all[c->rank][f];
        V.seg_own[f] = H.list(V.epoch, f) + off;
        if (n) CU(cudaMemcpyAsync(V.seg_own[f], V.b_vis[f].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    H.publish(V.epoch);
    V.seg_published = true;
    V.lists_on_host = true; // the own lists are readable in place (fyx_get_visible)
    V.own_in_seg = true;
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
    V.host_copy_private = true; // stand-alone exchange: a rank that wants the lists on the host copies them itself
    int32_t rc = allgather_enqueue(c, V, c->stream);
    if (rc) return rc;
    rc = allgather_finish(c, V, c->stream);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_x1[V.epoch & 1], c->stream));
    CU(cudaEventRecord(V.ev_gather, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_comm_get_stats(fyx_ctx *c, fyx_comm_stats *out)
{
    if (!c || !out) return FYX_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof *out);
    if (!c->comm || !c->gather_epoch) return fail(c, FYX_ERR_STATE, "no gathered frame has run yet");
    CU(cudaSetDevice(c->device));
    const int p = (int)(c->gather_epoch & 1);
    VisSlot &V = c->vs[c->x_slot[p]];
    if (V.epoch != c->x_epoch[p] || !V.gathered) return fail(c, FYX_ERR_STATE, "the last exchange's frame has been overwritten");
    CU(cudaEventSynchronize(c->ev_x1[p]));
    CU(cudaEventElapsedTime(&out->device_ms, c->ev_x0[p], c->ev_x1[p]));
    int32_t rc = resolve_counts(c, V);
    if (rc) return rc;
    out->epoch = V.epoch;
    for (uint32_t f = 0; f < V.nf; ++f) {
        out->entries_own += V.counts_all[c->rank][f];
        out->entries_total += V.gath_count[f];
    }
    out->egress_bytes = V.via_peer ? 4ull * out->entries_own * (uint64_t)(c->nranks - 1) : 0ull;
    out->mode = fyx_comm_mode(c);
    return FYX_OK;
}

extern "C" int32_t fyx_get_visible_gathered_device(fyx_ctx *c, uint32_t f, const uint32_t **d_idx, uint32_t *out_count)
{
    if (!c || !d_idx || !out_count) return FYX_ERR_INVALID_ARGUMENT;
    VisSlot &V = c->vs[c->cur];
    if (f >= V.nf || !V.gathered) return fail(c, FYX_ERR_INVALID_ARGUMENT, "frustum %u has no gathered list", f);
    CU(cudaSetDevice(c->device));
    int32_t rc = resolve_counts(c, V);
    if (rc) return rc;
    *d_idx = V.gath_ptr[f];
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
    int32_t rc = resolve_counts(c, V);
    if (rc) return rc;
    if (c->hostseg_ready && !V.host_copy_private) {
        // the whole list is assembled in the node-wide segment: publish this rank's part if a synchronous frame has not
        // done so yet, then wait until every rank's part of this epoch is there
        if (!V.seg_published) {
            CU(cudaStreamWaitEvent(c->stream, V.ev_gather, 0));
            if ((rc = hostseg_publish(c, V, c->stream))) return rc;
        }
        if (!V.seg_complete) {
            if (!c->hostseg.wait_complete(V.epoch)) return fail(c, FYX_ERR_NCCL, "%s", c->hostseg.err.c_str());
            V.seg_complete = true;
        }
        *out_idx = c->hostseg.list(V.epoch, f);
        *out_count = V.gath_count[f];
        return FYX_OK;
    }
    if (!V.gathered_on_host) {
        // bring all frusta at once (one synchronisation), after the collective stream has produced them
        CU(cudaStreamWaitEvent(c->stream, V.ev_gather, 0));
        for (uint32_t g = 0; g < V.nf; ++g) {
            const size_t m = V.gath_count[g];
            rc = host_gath_ensure(c, V, g, m);
            if (rc) return rc;
            if (m) CU(cudaMemcpyAsync(V.h_gath[g], V.gath_ptr[g], m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
        V.gathered_on_host = true;
    }
    *out_idx = V.h_gath[f];
    *out_count = V.gath_count[f];
    return FYX_OK;
}
