// fyx_drawprep.inl — N3: instances of a visible list grouped by bundle (included at the end of fyx_api.cu).
// Kernels: fyx_drawprep.cu.  Reference: renderer/bundle.rs:118-127 (sort index), :483-487 (instance block),
// :1248-1278 (push into the bundle keyed by material / surface data / render path).

namespace {

int32_t inst_host_ensure(fyx_ctx *c, void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return FYX_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    const size_t nb = std::max<size_t>(4096, bytes + bytes / 2);
    CU(cudaHostAlloc(p, nb, cudaHostAllocDefault));
    *cap = nb;
    return FYX_OK;
}

void inst_free(fyx_ctx *c)
{
    for (auto &o : c->inst) {
        dev_free(o.b_node);
        dev_free(o.b_sort);
        dev_free(o.b_mats);
        dev_free(o.b_bundles);
        dev_free(o.b_block_of);
        dev_free(o.b_blocks);
        dev_free(o.b_surf);
        dev_free(o.b_skin);
        if (o.h_surf) cudaFreeHost(o.h_surf);
        o.h_surf = nullptr;
        for (int k = 0; k < 4; ++k) {
            if (o.h[k]) cudaFreeHost(o.h[k]);
            o.h[k] = nullptr;
            o.h_cap[k] = 0;
        }
        o.valid = false;
    }
    for (auto &V : c->vs)
        for (auto &b : V.b_vis_slot) dev_free(b);
    dev_free(c->b_lod_range);
    dev_free(c->b_lodp);
    for (auto &b : c->b_light) dev_free(b);
    dev_free(c->b_light_ptrs);
    dev_free(c->b_light_counts);
    if (c->h_light_counts) cudaFreeHost(c->h_light_counts);
    c->h_light_counts = nullptr;
    dev_free(c->b_bundle);
    dev_free(c->b_rank_slot);
    dev_free(c->b_inst_hist);
    dev_free(c->b_inst_first);
    dev_free(c->b_inst_offset);
    dev_free(c->b_inst_tmp);
    dev_free(c->b_inst_nb);
    if (c->h_inst_nb) cudaFreeHost(c->h_inst_nb);
    c->h_inst_nb = nullptr;
}

} // namespace

extern "C" int32_t fyx_set_bundle_ids(fyx_ctx *c, uint32_t count, const uint32_t *idx, const uint32_t *ids)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    if (!count) return FYX_OK;
    if (!ids) return fail(c, FYX_ERR_INVALID_ARGUMENT, "bundle_ids is NULL");
    CU(cudaSetDevice(c->device));
    uint32_t mx = 0;
    for (uint32_t i = 0; i < count; ++i) mx = std::max(mx, ids[i]);
    if (mx >= (1u << 24)) return fail(c, FYX_ERR_INVALID_ARGUMENT, "bundle id %u too large: ids must be dense (< 2^24)", mx);
    if (!c->have_bundles) {
        int32_t rc = dev_ensure(c, c->b_bundle, std::max<size_t>(c->n_slots, 1) * 4);
        if (rc) return rc;
        CU(cudaMemsetAsync(c->b_bundle.p, 0, std::max<size_t>(c->n_slots, 1) * 4, c->stream));
        c->have_bundles = true;
        c->n_bundle_ids = 1;
    }
    c->n_bundle_ids = std::max(c->n_bundle_ids, mx + 1);
    void *d_v = nullptr, *d_i = nullptr;
    int32_t rc = stage_to_device(c, ids, (size_t)count * 4, idx, idx ? (size_t)count * 4 : 0, false, &d_v, &d_i);
    if (rc) return rc;
    launch_scatter_u32(c->stream, c->b_bundle.as<uint32_t>(), count, static_cast<const uint32_t *>(d_i),
                       static_cast<const uint32_t *>(d_v), c->b_slot_of_node.as<uint32_t>(), c->n_nodes, 1);
    c->launches++;
    CU(cudaGetLastError());
    return FYX_OK;
}

// Mesh::surfaces of the listed nodes (scene/mesh/mod.rs:726-805: one SurfaceInstanceData per surface)
extern "C" int32_t fyx_set_node_surfaces(fyx_ctx *c, uint32_t count, const uint32_t *idx, const uint32_t *first, const uint32_t *bundle_ids,
                                         const uint32_t *skin_surface)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    if (!count) return FYX_OK;
    if (!idx || !first) return fail(c, FYX_ERR_INVALID_ARGUMENT, "idx / first are NULL");
    const uint32_t total = first[count];
    if (total && !bundle_ids) return fail(c, FYX_ERR_INVALID_ARGUMENT, "bundle_ids is NULL");
    for (uint32_t i = 0; i < count; ++i) {
        if (first[i + 1] < first[i]) return fail(c, FYX_ERR_INVALID_ARGUMENT, "first[] must not decrease");
        if (idx[i] >= c->n_nodes) return fail(c, FYX_ERR_INVALID_ARGUMENT, "node %u out of range", idx[i]);
    }
    uint32_t mx = 0;
    for (uint32_t k = 0; k < total; ++k) {
        mx = std::max(mx, bundle_ids[k]);
        if (skin_surface && skin_surface[k] != FYX_NONE && skin_surface[k] >= c->surfaces.size())
            return fail(c, FYX_ERR_INVALID_ARGUMENT, "skin surface id %u out of range", skin_surface[k]);
    }
    if (mx >= (1u << 24)) return fail(c, FYX_ERR_INVALID_ARGUMENT, "bundle id %u too large: ids must be dense (< 2^24)", mx);
    if (c->ms_of_node.size() < c->n_nodes) c->ms_of_node.resize(c->n_nodes, make_uint2(0u, 0u));
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t n = first[i + 1] - first[i];
        c->ms_of_node[idx[i]] = make_uint2((uint32_t)c->ms_bundle_h.size(), n); // appended; an older range of the node is abandoned
        for (uint32_t k = first[i]; k < first[i + 1]; ++k) {
            c->ms_bundle_h.push_back(bundle_ids[k]);
            c->ms_skin_h.push_back(skin_surface ? skin_surface[k] : FYX_NONE);
        }
    }
    if (total) c->n_bundle_ids = std::max(c->n_bundle_ids, mx + 1);
    c->have_ms = true;
    c->ms_dirty = true;
    return FYX_OK;
}

static int32_t ms_flush(fyx_ctx *c)
{
    if (!c->have_ms || !c->ms_dirty) return FYX_OK;
    std::vector<uint2> by_slot(std::max<uint32_t>(c->n_slots, 1), make_uint2(0u, 0u));
    for (uint32_t sl = 0; sl < c->n_slots; ++sl) {
        const uint32_t node = c->node_of_slot[sl];
        if (node < c->ms_of_node.size()) by_slot[sl] = c->ms_of_node[node];
    }
    int32_t rc;
    if ((rc = dev_ensure(c, c->b_ms_range, by_slot.size() * sizeof(uint2)))) return rc;
    if ((rc = dev_ensure(c, c->b_ms_bundle, std::max<size_t>(c->ms_bundle_h.size(), 1) * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_ms_skin, std::max<size_t>(c->ms_skin_h.size(), 1) * 4))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(c->b_ms_range.p, by_slot.data(), by_slot.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    if (!c->ms_bundle_h.empty()) {
        CU(cudaMemcpy(c->b_ms_bundle.p, c->ms_bundle_h.data(), c->ms_bundle_h.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->b_ms_skin.p, c->ms_skin_h.data(), c->ms_skin_h.size() * 4, cudaMemcpyHostToDevice));
    }
    c->ms_dirty = false;
    return FYX_OK;
}

extern "C" int32_t fyx_enable_instances(fyx_ctx *c, uint32_t enable)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    c->instances_enabled = enable != 0;
    return FYX_OK;
}

extern "C" int32_t fyx_pack_instances(fyx_ctx *c, uint32_t f, const float *view, const float *vp)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!view || !vp) return fail(c, FYX_ERR_INVALID_ARGUMENT, "view / view_projection matrix is NULL");
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    VisSlot &V = c->vs[c->cur];
    if (V.pending) return fail(c, FYX_ERR_STATE, "the frame is still in flight: call fyx_frame_wait first");
    if (f >= V.nf) return fail(c, FYX_ERR_INVALID_ARGUMENT, "frustum %u was not part of the last cull (%u frusta)", f, V.nf);
    if (!V.have_slots) return fail(c, FYX_ERR_STATE, "the last cull was made without fyx_enable_instances");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (!V.counts_on_host) {
        CU(cudaMemcpy2DAsync(V.h_counts, sizeof(uint32_t), V.d_counts, sizeof(uint32_t) * kCountStride, sizeof(uint32_t), V.nf,
                             cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        V.counts_on_host = true;
    }
    const uint32_t n = V.h_counts[f];
    const uint32_t nb = (c->have_bundles || c->have_ms) ? c->n_bundle_ids : 1u;
    InstOut &o = c->inst[f];
    o.valid = false;
    int32_t rc;
    if ((rc = commit_surfaces(c))) return rc; // surf_of_slot: which nodes are skinned
    if ((rc = ms_flush(c))) return rc;
    const size_t n1 = std::max<uint32_t>(n, 1);
    if ((rc = dev_ensure(c, o.b_bundles, (size_t)nb * sizeof(fyx_bundle)))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_hist, (size_t)nb * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_first, (size_t)nb * 8))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_offset, (size_t)nb * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_tmp, n1 * 8))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_nb, 256))) return rc;
    if (!c->h_inst_nb) CU(cudaHostAlloc(reinterpret_cast<void **>(&c->h_inst_nb), 64, cudaHostAllocDefault));
    // DFS rank per slot (bundle sort index = that of the instance the reference's DFS pushes first)
    const bool have_rank = c->dfs_rank.size() == c->n_nodes && c->n_nodes > 0;
    if (have_rank && !c->rank_on_device) {
        std::vector<uint32_t> r(c->n_slots);
        for (uint32_t sl = 0; sl < c->n_slots; ++sl) r[sl] = c->dfs_rank[c->node_of_slot[sl]];
        if ((rc = dev_ensure(c, c->b_rank_slot, std::max<size_t>(c->n_slots, 1) * 4))) return rc;
        CU(cudaStreamSynchronize(s));
        if (c->n_slots) CU(cudaMemcpy(c->b_rank_slot.p, r.data(), (size_t)c->n_slots * 4, cudaMemcpyHostToDevice));
        c->rank_on_device = true;
    }
    CU(cudaMemsetAsync(c->b_inst_hist.p, 0, (size_t)nb * 4, s));
    CU(cudaMemsetAsync(c->b_inst_first.p, 0xFF, (size_t)nb * 8, s));

    InstParams ip{};
    ip.n = n;
    ip.vis_node = V.b_vis[f].as<uint32_t>();
    ip.vis_slot = V.b_vis_slot[f].as<uint32_t>();
    ip.bundle_of_slot = c->have_bundles ? c->b_bundle.as<uint32_t>() : nullptr;
    ip.rank_of_slot = have_rank ? c->b_rank_slot.as<uint32_t>() : nullptr;
    ip.ms_range = c->have_ms ? c->b_ms_range.as<uint2>() : nullptr;
    ip.ms_bundle = c->b_ms_bundle.as<uint32_t>();
    ip.ms_skin = c->b_ms_skin.as<uint32_t>();
    ip.surf_of_slot = c->b_surf_of_slot.as<uint32_t>();
    memcpy(ip.view, view, 64);
    memcpy(ip.vp, vp, 64);
    ip.n_bundle_ids = nb;
    ip.hist = c->b_inst_hist.as<uint32_t>();
    ip.first_key = c->b_inst_first.as<unsigned long long>();
    ip.offset = c->b_inst_offset.as<uint32_t>();
    ip.tmp_sort = c->b_inst_tmp.as<uint64_t>();
    ip.o_bundles = o.b_bundles.as<fyx_bundle>();
    ip.o_n_bundles = c->b_inst_nb.as<uint32_t>();
    // phase 1: sort indices, histogram over (visible node, surface), scan -> number of bundles and of instances
    launch_inst_count(s, c->a, ip);
    c->launches += n ? 2 : 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_inst_nb, c->b_inst_nb.p, 12, cudaMemcpyDeviceToHost, s));
    rc = sync_and_check(c);
    if (rc) return rc;
    const uint32_t n_inst = c->h_inst_nb[2];
    const size_t m1 = std::max<uint32_t>(n_inst, 1);
    if ((rc = dev_ensure(c, o.b_node, m1 * 4))) return rc;
    if ((rc = dev_ensure(c, o.b_sort, m1 * 8))) return rc;
    if ((rc = dev_ensure(c, o.b_surf, m1 * 4))) return rc;
    if ((rc = dev_ensure(c, o.b_skin, m1 * 4))) return rc;
    if ((rc = dev_ensure(c, o.b_mats, m1 * 128))) return rc;
    ip.o_node = o.b_node.as<uint32_t>();
    ip.o_sort = o.b_sort.as<uint64_t>();
    ip.o_surf = o.b_surf.as<uint32_t>();
    ip.o_skin = o.b_skin.as<uint32_t>();
    ip.o_mats = o.b_mats.as<float4>();
    // phase 2: scatter into bundle order
    launch_inst_scatter(s, c->a, ip);
    c->launches += n ? 1 : 0;
    CU(cudaGetLastError());
    rc = sync_and_check(c);
    if (rc) return rc;
    o.count = n_inst;
    o.n_visible = n;
    o.n_bundles = c->h_inst_nb[0];
    o.on_host = false;
    o.valid = true;
    o.blocks_valid = false;
    return FYX_OK;
}

// N3: the bone-matrix blocks write_uniforms builds per skinned instance (renderer/bundle.rs:484-496), from the palettes of
// the last fyx_build_palettes / fyx_render_prep
extern "C" int32_t fyx_pack_bone_matrices(fyx_ctx *c, uint32_t f)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid) return fail(c, FYX_ERR_STATE, "fyx_pack_instances has not been called for frustum %u", f);
    CU(cudaSetDevice(c->device));
    int32_t rc = commit_surfaces(c);
    if (rc) return rc;
    InstOut &o = c->inst[f];
    cudaStream_t s = c->stream;
    if ((rc = dev_ensure(c, o.b_block_of, std::max<size_t>(o.count, 1) * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_inst_nb, 256))) return rc;
    uint32_t *counter = c->b_inst_nb.as<uint32_t>() + 1;
    CU(cudaMemsetAsync(counter, 0, 4, s));
    launch_bone_block_index(s, o.count, o.b_skin.as<uint32_t>(), o.b_block_of.as<uint32_t>(), counter);
    CU(cudaMemcpyAsync(c->h_inst_nb + 1, counter, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    o.n_blocks = c->h_inst_nb[1];
    if ((rc = dev_ensure(c, o.b_blocks, std::max<size_t>(o.n_blocks, 1) * FYX_MAX_BONES * 64))) return rc;
    launch_bone_blocks(s, o.count, o.b_skin.as<uint32_t>(), c->b_surf_bones.as<uint2>(), c->b_palette.as<float>(), o.b_block_of.as<uint32_t>(),
                       o.b_blocks.as<float>());
    c->launches += o.count ? 2 : 0;
    CU(cudaGetLastError());
    rc = sync_and_check(c);
    if (rc) return rc;
    o.blocks_valid = true;
    return FYX_OK;
}

extern "C" int32_t fyx_get_bone_matrix_blocks_device(fyx_ctx *c, uint32_t f, fyx_bone_blocks *out)
{
    if (!c || !out) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid || !c->inst[f].blocks_valid) return fail(c, FYX_ERR_STATE, "fyx_pack_bone_matrices has not been called for frustum %u", f);
    const InstOut &o = c->inst[f];
    out->count = o.count;
    out->n_blocks = o.n_blocks;
    out->block_of_instance = o.b_block_of.as<uint32_t>();
    out->blocks = o.b_blocks.as<float>();
    return FYX_OK;
}

extern "C" int32_t fyx_get_bone_matrix_block(fyx_ctx *c, uint32_t f, uint32_t instance, float *out_255x16, uint32_t *out_has_block)
{
    if (!c || !out_has_block) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid || !c->inst[f].blocks_valid) return fail(c, FYX_ERR_STATE, "fyx_pack_bone_matrices has not been called for frustum %u", f);
    const InstOut &o = c->inst[f];
    if (instance >= o.count) return fail(c, FYX_ERR_INVALID_ARGUMENT, "instance %u out of range (%u)", instance, o.count);
    CU(cudaSetDevice(c->device));
    uint32_t blk = FYX_NONE;
    CU(cudaMemcpy(&blk, o.b_block_of.as<uint32_t>() + instance, 4, cudaMemcpyDeviceToHost));
    *out_has_block = blk != FYX_NONE;
    if (blk != FYX_NONE && out_255x16) CU(cudaMemcpy(out_255x16, o.b_blocks.as<float>() + (size_t)blk * FYX_MAX_BONES * 16, (size_t)FYX_MAX_BONES * 64, cudaMemcpyDeviceToHost));
    return FYX_OK;
}

extern "C" int32_t fyx_get_instances_device(fyx_ctx *c, uint32_t f, fyx_instances *out)
{
    if (!c || !out) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid) return fail(c, FYX_ERR_STATE, "fyx_pack_instances has not been called for frustum %u", f);
    const InstOut &o = c->inst[f];
    out->count = o.count;
    out->n_bundles = o.n_bundles;
    out->node = o.b_node.as<uint32_t>();
    out->sort_index = o.b_sort.as<uint64_t>();
    out->matrices = o.b_mats.as<float>();
    out->bundles = o.b_bundles.as<fyx_bundle>();
    return FYX_OK;
}

extern "C" int32_t fyx_get_instances(fyx_ctx *c, uint32_t f, fyx_instances *out)
{
    if (!c || !out) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid) return fail(c, FYX_ERR_STATE, "fyx_pack_instances has not been called for frustum %u", f);
    InstOut &o = c->inst[f];
    CU(cudaSetDevice(c->device));
    if (!o.on_host) {
        const size_t bytes[4] = {(size_t)o.count * 4, (size_t)o.count * 8, (size_t)o.count * 128, (size_t)o.n_bundles * sizeof(fyx_bundle)};
        const void *src[4] = {o.b_node.p, o.b_sort.p, o.b_mats.p, o.b_bundles.p};
        for (int k = 0; k < 4; ++k) {
            int32_t rc = inst_host_ensure(c, &o.h[k], &o.h_cap[k], std::max<size_t>(bytes[k], 1));
            if (rc) return rc;
            if (bytes[k]) CU(cudaMemcpyAsync(o.h[k], src[k], bytes[k], cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
        o.on_host = true;
    }
    out->count = o.count;
    out->n_bundles = o.n_bundles;
    out->node = static_cast<const uint32_t *>(o.h[0]);
    out->sort_index = static_cast<const uint64_t *>(o.h[1]);
    out->matrices = static_cast<const float *>(o.h[2]);
    out->bundles = static_cast<const fyx_bundle *>(o.h[3]);
    return FYX_OK;
}

extern "C" int32_t fyx_get_instance_surfaces(fyx_ctx *c, uint32_t f, const uint32_t **out_ordinal)
{
    if (!c || !out_ordinal) return FYX_ERR_INVALID_ARGUMENT;
    if (f >= FYX_MAX_FRUSTA || !c->inst[f].valid) return fail(c, FYX_ERR_STATE, "fyx_pack_instances has not been called for frustum %u", f);
    InstOut &o = c->inst[f];
    CU(cudaSetDevice(c->device));
    int32_t rc = inst_host_ensure(c, &o.h_surf, &o.h_surf_cap, std::max<size_t>((size_t)o.count * 4, 1));
    if (rc) return rc;
    if (o.count) CU(cudaMemcpyAsync(o.h_surf, o.b_surf.p, (size_t)o.count * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out_ordinal = static_cast<const uint32_t *>(o.h_surf);
    return FYX_OK;
}

// ---- N4 (light list): renderer/bundle.rs:926-974 --------------------------------------------------------------------
extern "C" int32_t fyx_cull_lights(fyx_ctx *c)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    const uint32_t nf = (uint32_t)c->cp.nf;
    if (!nf) return fail(c, FYX_ERR_STATE, "no cull has been made yet: the light lists use its frusta");
    if (c->vs[c->cur].pending) return fail(c, FYX_ERR_STATE, "the frame is still in flight: call fyx_frame_wait first");
    CU(cudaSetDevice(c->device));
    int32_t rc;
    uint32_t *ptrs[FYX_MAX_FRUSTA] = {};
    for (uint32_t f = 0; f < nf; ++f) {
        if ((rc = dev_ensure(c, c->b_light[f], std::max<size_t>(c->n_slots, 1) * 4))) return rc;
        ptrs[f] = c->b_light[f].as<uint32_t>();
    }
    if ((rc = dev_ensure(c, c->b_light_ptrs, sizeof ptrs))) return rc;
    if ((rc = dev_ensure(c, c->b_light_counts, sizeof(uint32_t) * kCountStride * FYX_MAX_FRUSTA))) return rc;
    if (!c->h_light_counts) CU(cudaHostAlloc(reinterpret_cast<void **>(&c->h_light_counts), sizeof(uint32_t) * FYX_MAX_FRUSTA, cudaHostAllocDefault));
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->b_light_ptrs.p, ptrs, sizeof ptrs, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(c->b_light_counts.p, 0, sizeof(uint32_t) * kCountStride * FYX_MAX_FRUSTA, s));
    launch_cull_lights(s, c->a, c->cp, c->b_light_ptrs.as<uint32_t *>(), c->b_light_counts.as<uint32_t>());
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy2DAsync(c->h_light_counts, sizeof(uint32_t), c->b_light_counts.p, sizeof(uint32_t) * kCountStride, sizeof(uint32_t), nf,
                         cudaMemcpyDeviceToHost, s));
    rc = sync_and_check(c);
    if (rc) return rc;
    for (uint32_t f = 0; f < nf; ++f) {
        c->h_light[f].resize(c->h_light_counts[f]);
        if (c->h_light_counts[f])
            CU(cudaMemcpy(c->h_light[f].data(), c->b_light[f].p, (size_t)c->h_light_counts[f] * 4, cudaMemcpyDeviceToHost));
        std::sort(c->h_light[f].begin(), c->h_light[f].end()); // pool order, like the reference's pair_iter
    }
    c->light_nf = nf;
    c->lights_valid = true;
    return FYX_OK;
}

extern "C" int32_t fyx_select_reflection_probes(fyx_ctx *c, uint32_t count, uint32_t *out)
{
    if (!c || !out) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    if (count != c->observers.size()) return fail(c, FYX_ERR_INVALID_ARGUMENT, "%u results asked for, %zu observers set (fyx_set_observers)", count, c->observers.size());
    if (!count) return FYX_OK;
    CU(cudaSetDevice(c->device));
    LodParams lp{};
    lp.nf = (int)count;
    for (uint32_t f = 0; f < count; ++f) {
        lp.ox[f] = c->observers[f].translation[0];
        lp.oy[f] = c->observers[f].translation[1];
        lp.oz[f] = c->observers[f].translation[2];
    }
    int32_t rc = dev_ensure(c, c->b_light_counts, sizeof(uint32_t) * kCountStride * FYX_MAX_FRUSTA);
    if (rc) return rc;
    uint32_t *best = c->b_light_counts.as<uint32_t>();
    CU(cudaMemsetAsync(best, 0, sizeof(uint32_t) * FYX_MAX_FRUSTA, c->stream));
    launch_select_probes(c->stream, c->a, lp, best);
    c->launches++;
    CU(cudaGetLastError());
    uint32_t h[FYX_MAX_FRUSTA];
    CU(cudaMemcpyAsync(h, best, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, c->stream));
    rc = sync_and_check(c);
    if (rc) return rc;
    for (uint32_t f = 0; f < count; ++f) out[f] = h[f] ? h[f] - 1u : FYX_NONE;
    c->lights_valid = false; // the scratch counters were reused
    return FYX_OK;
}

extern "C" int32_t fyx_get_visible_lights(fyx_ctx *c, uint32_t f, const uint32_t **out_idx, uint32_t *out_count)
{
    if (!c || !out_idx || !out_count) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->lights_valid) return fail(c, FYX_ERR_STATE, "fyx_cull_lights has not been called for the last cull");
    if (f >= c->light_nf) return fail(c, FYX_ERR_INVALID_ARGUMENT, "frustum %u was not part of the last cull (%u frusta)", f, c->light_nf);
    *out_idx = c->h_light[f].data();
    *out_count = (uint32_t)c->h_light[f].size();
    return FYX_OK;
}

// ---- N4 (LOD filter): renderer/bundle.rs:898-916, 988-1004 -------------------------------------------------------------
static bool lod_active(const fyx_ctx *c, uint32_t nf) { return c->have_lod && nf && c->observers.size() == nf; }

static bool unfused_cull(const fyx_ctx *c, uint32_t nf) { return nf && (lod_active(c, nf) || c->maybe_static_batch); }

static int32_t lod_pass(fyx_ctx *c);

// The cull as its own pass over the finished boxes: LOD filter bits first (if any), then either one launch over all slots
// or — with static batches in the graph — one launch per hierarchy level, parents first, carrying the pruned frusta down.
static int32_t cull_unfused(fyx_ctx *c, uint32_t nf)
{
    const bool lod = lod_active(c, nf);
    int32_t rc;
    if (lod && (rc = lod_pass(c))) return rc;
    const uint32_t *lodp = lod ? c->b_lodp.as<uint32_t>() : nullptr;
    if (!c->maybe_static_batch) {
        launch_cull(c->stream, c->a, c->cp, lodp);
        c->launches++;
    } else {
        if ((rc = dev_ensure(c, c->b_prune, std::max<size_t>(c->n_slots, 1) * 4))) return rc;
        const size_t nl = c->level_off.size() ? c->level_off.size() - 1 : 0;
        for (size_t l = 0; l < nl; ++l) {
            launch_cull_range(c->stream, c->a, c->cp, lodp, c->level_off[l], c->level_off[l + 1], c->b_prune.as<uint32_t>());
            c->launches += (c->level_off[l + 1] > c->level_off[l]);
        }
    }
    CU(cudaGetLastError());
    return FYX_OK;
}

// hidden-frusta bits of every node, one small launch per hierarchy level (parents first)
static int32_t lod_pass(fyx_ctx *c)
{
    LodParams lp{};
    lp.nf = (int)c->observers.size();
    for (int f = 0; f < lp.nf; ++f) {
        const fyx_observer &o = c->observers[f];
        lp.ox[f] = o.translation[0];
        lp.oy[f] = o.translation[1];
        lp.oz[f] = o.translation[2];
        lp.zn[f] = o.z_near;
        lp.zr[f] = o.z_far - o.z_near; // z_range, as the reference computes it
    }
    int32_t rc = dev_ensure(c, c->b_lodp, std::max<size_t>(c->n_slots, 1) * 4);
    if (rc) return rc;
    const size_t nl = c->level_off.size() ? c->level_off.size() - 1 : 0;
    for (size_t l = 0; l < nl; ++l) {
        launch_lod_level(c->stream, c->a, c->level_off[l], c->level_off[l + 1], c->b_lod_range.as<float2>(), c->b_lodp.as<uint32_t>(), lp);
        c->launches += (c->level_off[l + 1] > c->level_off[l]);
    }
    CU(cudaGetLastError());
    return FYX_OK;
}

extern "C" int32_t fyx_set_lod_ranges(fyx_ctx *c, uint32_t count, const uint32_t *idx, const float *begin_end)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    if (!count) return FYX_OK;
    if (!begin_end) return fail(c, FYX_ERR_INVALID_ARGUMENT, "begin_end is NULL");
    CU(cudaSetDevice(c->device));
    // host-side scatter through the slot map into a staging image of the column (LOD objects are few, calls are rare)
    const size_t n = std::max<size_t>(c->n_slots, 1);
    std::vector<float2> col(n);
    int32_t rc = dev_ensure(c, c->b_lod_range, n * sizeof(float2));
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    if (c->have_lod) {
        CU(cudaMemcpy(col.data(), c->b_lod_range.p, n * sizeof(float2), cudaMemcpyDeviceToHost));
    } else {
        const float qnan = std::nanf("");
        std::fill(col.begin(), col.end(), make_float2(qnan, qnan));
    }
    for (uint32_t e = 0; e < count; ++e) {
        const uint32_t node = idx ? idx[e] : e;
        if (node >= c->n_nodes) continue;
        const uint32_t slot = c->slot_of_node[node];
        if (slot == FYX_NONE) continue;
        col[slot] = make_float2(begin_end[2 * (size_t)e], begin_end[2 * (size_t)e + 1]);
    }
    CU(cudaMemcpy(c->b_lod_range.p, col.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    c->have_lod = true;
    return FYX_OK;
}

extern "C" int32_t fyx_set_observers(fyx_ctx *c, uint32_t count, const fyx_observer *obs)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (count > FYX_MAX_FRUSTA) return fail(c, FYX_ERR_INVALID_ARGUMENT, "more than FYX_MAX_FRUSTA observers");
    if (count && !obs) return fail(c, FYX_ERR_INVALID_ARGUMENT, "observers is NULL");
    c->observers.assign(obs, obs + count);
    return FYX_OK;
}
