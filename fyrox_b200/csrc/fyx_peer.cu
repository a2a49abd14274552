// fyx_peer.cu — the all-gather of the visible lists as direct stores into peer GPUs' memory over NVLink 5 / NVSwitch.
//
// The exchange is a pure placement problem: rank r's visible entries of frustum f belong at
//   [ sum_{q<r} count[q][f],  ... + count[r][f] )   of every rank's list of f.
// With every rank's list buffers (one cudaMalloc per rank, cudaIpc-mapped by all the others) addressable from
// the kernels, the ranks do it themselves, without the host in the loop:
//   1. k_peer_counts   writes this rank's F counters into the count table of every rank, then a flag per rank
//                      (st.release.sys) carrying the frame's epoch; then waits (ld.acquire.sys) until every
//                      rank's flag shows the epoch: all counts are here, offsets are computable;
//   2. k_peer_push     copies the local compacted list of every frustum to its final place in all N buffers
//                      (128-bit stores, destination-aligned; 4*N bytes of NVLink egress per entry), then the
//                      last CTA raises this rank's "done" flag at every rank;
//   3. k_peer_wait     waits for every rank's done flag: this rank's gathered lists are complete; writes the
//                      per-frustum totals where the host (and device consumers) can read them.
// All three run on the context's high-priority collective stream beside the palette / skinning kernels of the
// frame.  No padding, no pack kernel, no host synchronisation (the NCCL path in fyx_comm.inl needs the counts
// on the host before it can size its fixed-slot payload).  Buffers alternate with the epoch's parity; a rank
// raises its count flag for epoch e only after its own consumers of epoch e-2 are done (stream order), and
// nobody pushes epoch e before having seen every count flag of e, so a buffer is never overwritten while its
// owner still reads it.  Spins are bounded (~4 s of %globaltimer): a rank that never arrives sets E_PEER_TIMEOUT
// instead of hanging the GPU.  The reference has no counterpart (single process).
#include "fyx_internal.h"

namespace fyx {

__device__ __forceinline__ void st_release_sys(uint32_t *p, const uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag >= epoch (epochs only grow); false on timeout
__device__ __forceinline__ bool spin_until(const uint32_t *flag, const uint32_t epoch)
{
    if (ld_acquire_sys(flag) >= epoch) return true;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > 4000000000ull) return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) k_peer_counts(const PeerParams pp)
{
    const int t = threadIdx.x, R = pp.nranks, nf = pp.nf, s = pp.epoch & 1u;
    // counts[s][me][f] at every rank
    if (t < R * nf) {
        const int r = t / nf, f = t % nf;
        peer_ctrl(pp, r)->counts[s][pp.rank][f] = pp.own_counts[f * kCountStride];
    }
    __threadfence_system();
    __syncthreads();
    if (t < R) st_release_sys(&peer_ctrl(pp, t)->cnt_flag[s][pp.rank], pp.epoch);
    // every rank's counts have arrived here
    bool ok = true;
    if (t < R) ok = spin_until(&peer_ctrl(pp, pp.rank)->cnt_flag[s][t], pp.epoch);
    if (!ok) atomicOr(pp.d_err, E_PEER_TIMEOUT);
}

__global__ void __launch_bounds__(256) k_peer_push(const PeerParams pp)
{
    const int R = pp.nranks, nf = pp.nf, s = pp.epoch & 1u;
    const PeerCtrl *mine = peer_ctrl(pp, pp.rank);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    for (int f = 0; f < nf; ++f) {
        uint32_t off = 0;
        for (int q = 0; q < pp.rank; ++q) off += mine->counts[s][q][f];
        const uint32_t n = mine->counts[s][pp.rank][f];
        if (!n) continue;
        const uint32_t *src = pp.own_list[f];
        // destination-aligned 128-bit body: entries [head, head + 4*groups) of the local list land at 16-byte-aligned
        // addresses of every destination (the buffers are 256-byte aligned and the offset is the same everywhere)
        const uint32_t head = min(n, (4u - (off & 3u)) & 3u);
        const uint32_t groups = (n - head) >> 2;
        const uint32_t tail0 = head + 4u * groups;
        for (uint32_t g = tid; g < groups; g += nthreads) {
            const uint32_t i = head + 4u * g;
            const uint4 v = make_uint4(src[i], src[i + 1], src[i + 2], src[i + 3]);
            for (int r = 0; r < R; ++r) {
                const int dst = (pp.rank + 1 + r) % R; // start with the neighbour: the ranks do not all hit rank 0 first
                *reinterpret_cast<uint4 *>(peer_list(pp, dst, s, f) + off + i) = v;
            }
        }
        // ragged ends (at most 3 + 3 entries)
        if (tid < head + (n - tail0)) {
            const uint32_t i = tid < head ? tid : tail0 + (tid - head);
            const uint32_t v = src[i];
            for (int r = 0; r < R; ++r) peer_list(pp, (pp.rank + 1 + r) % R, s, f)[off + i] = v;
        }
    }
    // this rank's part has landed everywhere once every CTA got here: the last one tells the others
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(pp.cta_done, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            *pp.cta_done = 0u;
            for (int r = 0; r < R; ++r) st_release_sys(&peer_ctrl(pp, r)->done_flag[s][pp.rank], pp.epoch);
        }
    }
}

__global__ void __launch_bounds__(32) k_peer_wait(const PeerParams pp)
{
    const int t = threadIdx.x, R = pp.nranks, s = pp.epoch & 1u;
    PeerCtrl *mine = peer_ctrl(pp, pp.rank);
    bool ok = true;
    if (t < R) ok = spin_until(&mine->done_flag[s][t], pp.epoch);
    if (!ok) atomicOr(pp.d_err, E_PEER_TIMEOUT);
    __syncwarp();
    if (t < pp.nf) {
        uint32_t total = 0;
        for (int q = 0; q < R; ++q) total += mine->counts[s][q][t];
        mine->totals[s][t] = total;
    }
}

void launch_peer_counts(cudaStream_t s, const PeerParams &pp) { k_peer_counts<<<1, 256, 0, s>>>(pp); }

void launch_peer_push(cudaStream_t s, const PeerParams &pp, unsigned push_ctas)
{
    k_peer_push<<<push_ctas, 256, 0, s>>>(pp);
    k_peer_wait<<<1, 32, 0, s>>>(pp);
}

} // namespace fyx
