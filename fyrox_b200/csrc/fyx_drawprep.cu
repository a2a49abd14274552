// fyx_drawprep.cu — N3: the step right after the cull.
//
// The reference, for every node the DFS of RenderDataBundleStorage::from_graph accepts, runs
// Mesh::collect_render_data (scene/mesh/mod.rs:691-805): sort index of the node's global position
// (RenderContext::calculate_sorting_index, renderer/bundle.rs:118-127), then one push() per surface into the
// bundle keyed by (material, surface data, render path) (bundle.rs:1248-1278); RenderDataBundle::write_uniforms
// later writes `world` and `view_projection * world` per instance (bundle.rs:483-487).  Here the visible list of
// one frustum becomes, in three launches, instance arrays grouped by bundle (a counting sort keyed by the
// host-assigned bundle id) plus the bundle table (first, count, sort index of the first-pushed instance).
// HBM-bound gather/scatter: per instance 4+4 B of list + 48 B G + 4 B flags read, 128+8+4 B written.
#include "fyx_internal.h"

namespace fyx {

namespace {

// RenderContext::calculate_sorting_index: z of view_matrix.transform_point(global_position) (nalgebra: divided
// by n = row3·p + m33 unless n == 0), * 1000, `as i64` (saturating, NaN -> 0), u64::MAX/2 saturating_add_signed.
__device__ __forceinline__ uint64_t sorting_index(const float *view, const float px, const float py, const float pz)
{
    const float n = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(view[3], px), FYX_MUL(view[7], py)), FYX_MUL(view[11], pz)), view[15]);
    float z = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(view[2], px), FYX_MUL(view[6], py)), FYX_MUL(view[10], pz)), view[14]);
    if (n != 0.0f) z = __fdiv_rn(z, n);
    const float zf = FYX_MUL(z, 1000.0f);
    const long long d = __float2ll_rz(zf); // cvt.rzi.s64.f32: saturates, NaN -> 0, like Rust's `as i64`
    const uint64_t center = 0xFFFFFFFFFFFFFFFFull / 2ull;
    if (d >= 0) {
        const uint64_t ud = (uint64_t)d;
        return (center > 0xFFFFFFFFFFFFFFFFull - ud) ? 0xFFFFFFFFFFFFFFFFull : center + ud;
    }
    const uint64_t ud = (uint64_t)(-(d + 1)) + 1ull;
    return (ud > center) ? 0ull : center - ud;
}

// surfaces of the visible entry's node: (first, count) — count 0 = the default single surface
__device__ __forceinline__ uint2 surfaces_of(const InstParams &ip, const uint32_t slot)
{
    return ip.ms_range ? ip.ms_range[slot] : make_uint2(0u, 0u);
}
__device__ __forceinline__ uint32_t warp_max_u32(const uint32_t v) { return __reduce_max_sync(0xFFFFFFFFu, v); }

// pass 1: sort index per visible entry, bundle histogram over its surfaces (Mesh::collect_render_data pushes one instance per
// surface, all with the node's sort index, scene/mesh/mod.rs:726-805), first-pushed instance per bundle
__global__ void __launch_bounds__(kBlock) k_inst_keys(const NodeArrays a, const InstParams ip)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const bool valid = i < ip.n;
    uint32_t slot = 0u, nsurf = 0u;
    uint2 sr = make_uint2(0u, 0u);
    unsigned long long key = ~0ull;
    if (valid) {
        slot = ip.vis_slot[i];
        // Base::global_position() = translation column of the global transform
        const float px = a.G[0][slot].w, py = a.G[1][slot].w, pz = a.G[2][slot].w;
        ip.tmp_sort[i] = sorting_index(ip.view, px, py, pz);
        const uint32_t rank = ip.rank_of_slot ? ip.rank_of_slot[slot] : ip.vis_node[i];
        key = ((unsigned long long)rank << 32) | i;
        sr = surfaces_of(ip, slot);
        nsurf = sr.y ? sr.y : 1u;
    }
    const uint32_t rounds = warp_max_u32(nsurf);
    const int lane = threadIdx.x & 31;
    for (uint32_t k = 0; k < rounds; ++k) {
        const bool act = k < nsurf;
        uint32_t b = 0xFFFFFFFFu;
        if (act) b = sr.y ? ip.ms_bundle[sr.x + k] : (ip.bundle_of_slot ? ip.bundle_of_slot[slot] : 0u);
        // warp-aggregated histogram: one atomic per distinct bundle in the warp
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
        if (act) {
            if ((peers & ((1u << lane) - 1u)) == 0u) atomicAdd(ip.hist + b, (uint32_t)__popc(peers));
            // the minimum only ever decreases: a plain read is a safe filter in front of the atomic
            if (key < *reinterpret_cast<volatile unsigned long long *>(ip.first_key + b)) atomicMin(ip.first_key + b, key);
        }
    }
}

// pass 2 (one CTA): exclusive scan of the histogram -> first instance of every bundle, table of the non-empty
// bundles; the histogram is zeroed again (pass 3 uses it as the scatter cursor)
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) k_inst_scan(const InstParams ip)
{
    __shared__ uint32_t s_wsum[2][kScanBlock / 32];
    __shared__ uint32_t s_carry[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < ip.n_bundle_ids; base += kScanBlock) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t c = (b < ip.n_bundle_ids) ? ip.hist[b] : 0u;
        const uint32_t ne = c ? 1u : 0u;
        uint32_t xc = c, xn = ne; // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tc = __shfl_up_sync(0xFFFFFFFFu, xc, d), tn = __shfl_up_sync(0xFFFFFFFFu, xn, d);
            if (lane >= d) { xc += tc; xn += tn; }
        }
        if (lane == 31) { s_wsum[0][warp] = xc; s_wsum[1][warp] = xn; }
        __syncthreads();
        if (warp == 0) {
            uint32_t wc = s_wsum[0][lane], wn = s_wsum[1][lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tc = __shfl_up_sync(0xFFFFFFFFu, wc, d), tn = __shfl_up_sync(0xFFFFFFFFu, wn, d);
                if (lane >= d) { wc += tc; wn += tn; }
            }
            s_wsum[0][lane] = wc; // inclusive over warps
            s_wsum[1][lane] = wn;
        }
        __syncthreads();
        const uint32_t off = s_carry[0] + (warp ? s_wsum[0][warp - 1] : 0u) + xc - c;
        const uint32_t k = s_carry[1] + (warp ? s_wsum[1][warp - 1] : 0u) + xn - ne;
        if (b < ip.n_bundle_ids) {
            ip.offset[b] = off;
            ip.hist[b] = 0u;
            if (c) {
                fyx_bundle r;
                r.id = b;
                r.first = off;
                r.count = c;
                r.reserved = 0u;
                r.sort_index = ip.tmp_sort[(uint32_t)(ip.first_key[b] & 0xFFFFFFFFull)];
                ip.o_bundles[k] = r;
            }
        }
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) {
            s_carry[0] += s_wsum[0][kScanBlock / 32 - 1];
            s_carry[1] += s_wsum[1][kScanBlock / 32 - 1];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ip.o_n_bundles[0] = s_carry[1];
        ip.o_n_bundles[2] = s_carry[0]; // instances = sum of the histogram
    }
}

// pass 3: scatter into bundle order; matrices are written by 8 lanes per instance (one float4 column each), so a
// warp store covers four whole 128-byte instance records
__global__ void __launch_bounds__(kBlock) k_inst_scatter(const NodeArrays a, const InstParams ip)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < ip.n;
    uint32_t slot = 0u, nsurf = 0u, node_flags = 0u;
    uint2 sr = make_uint2(0u, 0u);
    if (valid) {
        slot = ip.vis_slot[i];
        sr = surfaces_of(ip, slot);
        nsurf = sr.y ? sr.y : 1u;
        node_flags = a.flags[slot];
    }
    const uint32_t rounds = warp_max_u32(nsurf);
    for (uint32_t k = 0; k < rounds; ++k) {
        const bool act = k < nsurf;
        uint32_t b = 0xFFFFFFFFu, pos = 0u, skin = FYX_NONE;
        if (act) {
            b = sr.y ? ip.ms_bundle[sr.x + k] : (ip.bundle_of_slot ? ip.bundle_of_slot[slot] : 0u);
            skin = sr.y ? ip.ms_skin[sr.x + k] : ip.surf_of_slot[slot];
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
        uint32_t base = 0u;
        const int leader = __ffs(peers) - 1;
        if (act && lane == leader) base = ip.offset[b] + atomicAdd(ip.hist + b, (uint32_t)__popc(peers));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (act) {
            pos = base + __popc(peers & ((1u << lane) - 1u));
            ip.o_node[pos] = ip.vis_node[i];
            ip.o_sort[pos] = ip.tmp_sort[i];
            ip.o_surf[pos] = k;
            ip.o_skin[pos] = skin;
        }
        // the instance's world matrix: identity for a skinned surface (its vertices are placed by the bone palette,
        // scene/mesh/mod.rs:733-737) and for a static batch (:716), else the node's global transform
        const bool ident = act && ((skin != FYX_NONE) || (node_flags & FYX_NODE_STATIC_BATCH));
        const int j = lane & 7, c = j & 3;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = q * 4 + (lane >> 3);
            const uint32_t pe = __shfl_sync(0xFFFFFFFFu, pos, e);
            const uint32_t se = __shfl_sync(0xFFFFFFFFu, slot, e);
            const bool ve = __shfl_sync(0xFFFFFFFFu, (int)act, e) != 0;
            const bool ie = __shfl_sync(0xFFFFFFFFu, (int)ident, e) != 0;
            if (!ve) continue;
            float w0, w1, w2, w3 = (c == 3) ? 1.0f : 0.0f;
            if (ie) {
                w0 = (c == 0) ? 1.0f : 0.0f;
                w1 = (c == 1) ? 1.0f : 0.0f;
                w2 = (c == 2) ? 1.0f : 0.0f;
            } else {
                const float4 r0 = a.G[0][se], r1 = a.G[1][se], r2 = a.G[2][se];
                w0 = (c == 0) ? r0.x : (c == 1) ? r0.y : (c == 2) ? r0.z : r0.w;
                w1 = (c == 0) ? r1.x : (c == 1) ? r1.y : (c == 2) ? r1.z : r1.w;
                w2 = (c == 0) ? r2.x : (c == 1) ? r2.y : (c == 2) ? r2.z : r2.w;
            }
            float4 o;
            if (j < 4) {
                o = make_float4(w0, w1, w2, w3);
            } else {
                // (view_projection * world)[r, c], nalgebra order (Appendix A1), full 4x4: VP is projective
                const float *m = ip.vp;
                o.x = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(m[0], w0), FYX_MUL(m[4], w1)), FYX_MUL(m[8], w2)), FYX_MUL(m[12], w3));
                o.y = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(m[1], w0), FYX_MUL(m[5], w1)), FYX_MUL(m[9], w2)), FYX_MUL(m[13], w3));
                o.z = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(m[2], w0), FYX_MUL(m[6], w1)), FYX_MUL(m[10], w2)), FYX_MUL(m[14], w3));
                o.w = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(m[3], w0), FYX_MUL(m[7], w1)), FYX_MUL(m[11], w2)), FYX_MUL(m[15], w3));
            }
            ip.o_mats[(size_t)pe * 8u + j] = o;
        }
    }
}

} // namespace

// N4 LOD filter, one hierarchy level: bit f of lodp[slot] = "hidden for observer f by the node's own LOD range or by an
// ancestor's" (renderer/bundle.rs:898-916: normalised distance outside [begin, end]; :995: the DFS does not descend).
__global__ void __launch_bounds__(kBlock) k_lod_level(const NodeArrays a, const uint32_t lo, const uint32_t hi, const float2 *range,
                                                      uint32_t *lodp, const LodParams lp)
{
    const uint32_t slot = lo + blockIdx.x * kBlock + threadIdx.x;
    if (slot >= hi) return;
    const uint32_t p = a.parent[slot];
    uint32_t hidden = (p != FYX_NONE) ? lodp[p] : 0u;
    const float2 r = range[slot];
    if (r.x == r.x) { // the node is an object of a LOD level
        const float gx = a.G[0][slot].w, gy = a.G[1][slot].w, gz = a.G[2][slot].w; // global_position()
        for (int f = 0; f < lp.nf; ++f) {
            // metric_distance: |observer - position|, nalgebra's left-to-right dot, then sqrt
            const float dx = FYX_ADD(lp.ox[f], -gx), dy = FYX_ADD(lp.oy[f], -gy), dz = FYX_ADD(lp.oz[f], -gz);
            const float dist = __fsqrt_rn(FYX_ADD(FYX_ADD(FYX_MUL(dx, dx), FYX_MUL(dy, dy)), FYX_MUL(dz, dz)));
            const float normalized = __fdiv_rn(FYX_ADD(dist, -lp.zn[f]), lp.zr[f]);
            const bool visible = (normalized >= r.x) && (normalized <= r.y);
            if (!visible) hidden |= 1u << f;
        }
    }
    lodp[slot] = hidden;
}

void launch_lod_level(cudaStream_t s, const NodeArrays &a, uint32_t lo, uint32_t hi, const float2 *range, uint32_t *lodp, const LodParams &lp)
{
    if (hi <= lo) return;
    k_lod_level<<<(hi - lo + kBlock - 1) / kBlock, kBlock, 0, s>>>(a, lo, hi, range, lodp, lp);
}

// N3, bone matrices of the packed instances — RenderDataBundle::write_uniforms (renderer/bundle.rs:484-496): a skinned instance
// gets a block of MAX_BONE_MATRICES (255) mat4: its bone_matrices, then all-zero matrices; an unskinned one gets none.
// pass 1: which instance gets which block (order of the blocks is unspecified); inst_skin = the fyx surface id that skins the instance
__global__ void __launch_bounds__(kBlock) k_bone_block_index(const uint32_t n, const uint32_t *inst_skin, uint32_t *block_of_inst, uint32_t *counter)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    block_of_inst[i] = (inst_skin[i] != FYX_NONE) ? atomicAdd(counter, 1u) : FYX_NONE;
}
// pass 2: one CTA per instance copies the surface's palette and pads with zeros (1020 float4 per block)
__global__ void __launch_bounds__(kBlock) k_bone_blocks(const uint32_t *inst_skin, const uint2 *surf_bones /* (first palette entry, n_bones) */,
                                                        const float4 *palette, const uint32_t *block_of_inst, float4 *blocks)
{
    const uint32_t i = blockIdx.x;
    const uint32_t blk = block_of_inst[i];
    if (blk == FYX_NONE) return;
    const uint2 sb = surf_bones[inst_skin[i]];
    float4 *dst = blocks + (size_t)blk * (FYX_MAX_BONES * 4);
    const float4 *src = palette + (size_t)sb.x * 4;
    for (uint32_t e = threadIdx.x; e < FYX_MAX_BONES * 4; e += kBlock) dst[e] = (e < sb.y * 4u) ? src[e] : make_float4(0.f, 0.f, 0.f, 0.f);
}

void launch_bone_block_index(cudaStream_t s, uint32_t n, const uint32_t *inst_skin, uint32_t *block_of_inst, uint32_t *counter)
{
    if (n) k_bone_block_index<<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(n, inst_skin, block_of_inst, counter);
}
void launch_bone_blocks(cudaStream_t s, uint32_t n, const uint32_t *inst_skin, const uint2 *surf_bones, const float *palette, const uint32_t *block_of_inst,
                        float *blocks)
{
    if (n) k_bone_blocks<<<n, kBlock, 0, s>>>(inst_skin, surf_bones, reinterpret_cast<const float4 *>(palette), block_of_inst, reinterpret_cast<float4 *>(blocks));
}

void launch_inst_count(cudaStream_t s, const NodeArrays &a, const InstParams &ip)
{
    if (ip.n) k_inst_keys<<<(ip.n + kBlock - 1) / kBlock, kBlock, 0, s>>>(a, ip);
    k_inst_scan<<<1, kScanBlock, 0, s>>>(ip);
}

void launch_inst_scatter(cudaStream_t s, const NodeArrays &a, const InstParams &ip)
{
    if (ip.n) k_inst_scatter<<<(ip.n + kBlock - 1) / kBlock, kBlock, 0, s>>>(a, ip);
}

} // namespace fyx
