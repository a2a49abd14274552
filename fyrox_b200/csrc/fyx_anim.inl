// fyx_anim.inl — N2: animation sampling on the device (included at the end of fyx_api.cu).
// Kernels: fyx_anim.cu.  Reference: fyrox-animation/src/lib.rs (Animation), container.rs / track.rs (tracks),
// fyrox-math/src/curve.rs (curves), fyrox-impl/src/scene/animation/mod.rs (update_animations, pose application).

namespace {

// fyrox-math/src/lib.rs:179-203 on the host (this file is compiled with -ffp-contract=off)
float host_wrapf(float n, float min_limit, float max_limit)
{
    if (n >= min_limit && n <= max_limit) return n;
    if (max_limit == 0.0f && min_limit == 0.0f) return 0.0f;
    max_limit -= min_limit;
    const float offset = min_limit;
    min_limit = 0.0f;
    n -= offset;
    const float num_of_max = std::floor(std::fabs(n / max_limit));
    if (n >= max_limit) n -= num_of_max * max_limit;
    else if (n < min_limit) n += (num_of_max + 1.0f) * max_limit;
    return n + offset;
}

// Animation::set_time_position (fyrox-animation/src/lib.rs:432-440)
float host_time_position(const AnimStateDev &s, float t)
{
    if (s.looped) return host_wrapf(t, s.slice_start, s.slice_end);
    if (t < s.slice_start) t = s.slice_start;
    if (t > s.slice_end) t = s.slice_end;
    return t;
}

void anim_free(fyx_ctx *c)
{
    DevBuf *bufs[] = {&c->b_anim_keys, &c->b_anim_tracks, &c->b_anim_state, &c->b_anim_hints, &c->b_anim_values, &c->b_anim_ok,
                      &c->b_anim_bk, &c->b_anim_node_slot, &c->b_anim_node_begin, &c->b_anim_node_tracks};
    for (DevBuf *b : bufs) dev_free(*b);
    c->anims.clear();
    c->anim_tracks.clear();
    c->pend_keys.clear();
    c->pend_tracks.clear();
    c->pend_bk.clear();
    c->pend_state.clear();
    c->n_anim_keys = 0;
    c->n_blend_groups = 0;
    c->blend_groups.clear();
    c->an = AnimArrays{};
    c->anim_csr_dirty = true;
}

void anim_rebuild_arrays(fyx_ctx *c)
{
    AnimArrays &an = c->an;
    an.n_tracks = (uint32_t)c->anim_tracks.size();
    an.n_anims = (uint32_t)c->anims.size();
    an.keys = c->b_anim_keys.as<fyx_curve_key>();
    an.tracks = c->b_anim_tracks.as<AnimTrackDev>();
    an.state = c->b_anim_state.as<AnimStateDev>();
    an.hints = c->b_anim_hints.as<uint4>();
    an.values = c->b_anim_values.as<float4>();
    an.value_ok = c->b_anim_ok.as<uint32_t>();
    an.track_bind_kind = c->b_anim_bk.as<uint32_t>();
    an.node_slot = c->b_anim_node_slot.as<uint32_t>();
    an.node_begin = c->b_anim_node_begin.as<uint32_t>();
    an.node_tracks = c->b_anim_node_tracks.as<uint32_t>();
}

// animated nodes (distinct live targets, by slot) and their tracks in (animation, track) order
int32_t anim_build_csr(fyx_ctx *c)
{
    const uint32_t nt = (uint32_t)c->anim_tracks.size();
    // per node: directly applied animations first (animation order), then blend group after blend group with the
    // sources in the group's order; tracks in their own order inside an animation
    std::vector<uint32_t> track_anim(nt);
    for (uint32_t a = 0; a < c->anims.size(); ++a)
        for (uint32_t i = 0; i < c->anims[a].n_tracks; ++i) track_anim[c->anims[a].first_track + i] = a;
    struct Ent { uint32_t slot, group, pos, track; };
    std::vector<Ent> st;
    st.reserve(nt);
    for (uint32_t i = 0; i < nt; ++i) {
        const uint32_t node = c->anim_tracks[i].target_node;
        if (node >= c->n_nodes) continue; // invalid handle: logged and skipped by the reference
        const uint32_t slot = c->slot_of_node[node];
        if (slot == FYX_NONE) continue;
        const AnimHost &h = c->anims[track_anim[i]];
        st.push_back(Ent{slot, h.st.group, h.st.group ? h.pos_in_group : track_anim[i], i});
    }
    std::stable_sort(st.begin(), st.end(), [](const Ent &x, const Ent &y) {
        if (x.slot != y.slot) return x.slot < y.slot;
        if (x.group != y.group) return x.group < y.group; // 0 = directly applied: first
        return x.pos < y.pos;                             // tracks of one animation keep their order (stable sort)
    });
    std::vector<uint32_t> node_slot, node_begin, node_tracks(st.size());
    for (size_t k = 0; k < st.size(); ++k) {
        if (k == 0 || st[k].slot != st[k - 1].slot) {
            node_slot.push_back(st[k].slot);
            node_begin.push_back((uint32_t)k);
        }
        node_tracks[k] = st[k].track;
    }
    node_begin.push_back((uint32_t)st.size());
    int32_t rc;
    if ((rc = dev_ensure(c, c->b_anim_node_slot, std::max<size_t>(node_slot.size(), 1) * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_node_begin, node_begin.size() * 4))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_node_tracks, std::max<size_t>(node_tracks.size(), 1) * 4))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    if (!node_slot.empty()) CU(cudaMemcpy(c->b_anim_node_slot.p, node_slot.data(), node_slot.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->b_anim_node_begin.p, node_begin.data(), node_begin.size() * 4, cudaMemcpyHostToDevice));
    if (!node_tracks.empty()) CU(cudaMemcpy(c->b_anim_node_tracks.p, node_tracks.data(), node_tracks.size() * 4, cudaMemcpyHostToDevice));
    anim_rebuild_arrays(c);
    c->an.n_nodes = (uint32_t)node_slot.size();
    c->anim_csr_dirty = false;
    return FYX_OK;
}

} // namespace

namespace {

// upload what fyx_anim_add queued since the last flush (device buffers grow, keeping hints / time of the older animations)
int32_t anim_flush(fyx_ctx *c)
{
    if (c->pend_state.empty()) return FYX_OK;
    const uint32_t nt = (uint32_t)c->anim_tracks.size(), nk = c->n_anim_keys, na = (uint32_t)c->anims.size();
    const uint32_t t0 = nt - (uint32_t)c->pend_tracks.size(), k0 = nk - (uint32_t)c->pend_keys.size(), a0 = na - (uint32_t)c->pend_state.size();
    int32_t rc;
    if ((rc = dev_ensure(c, c->b_anim_keys, std::max<size_t>(nk, 1) * sizeof(fyx_curve_key), true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_tracks, std::max<size_t>(nt, 1) * sizeof(AnimTrackDev), true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_bk, std::max<size_t>(nt, 1) * 4, true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_hints, std::max<size_t>(nt, 1) * sizeof(uint4), true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_values, std::max<size_t>(nt, 1) * sizeof(float4), true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_ok, std::max<size_t>(nt, 1) * 4, true))) return rc;
    if ((rc = dev_ensure(c, c->b_anim_state, std::max<size_t>(na, 1) * sizeof(AnimStateDev), true))) return rc;
    CU(cudaStreamSynchronize(c->stream)); // the grown buffers' copies have landed; the sources below are pageable
    if (!c->pend_keys.empty())
        CU(cudaMemcpy(c->b_anim_keys.as<fyx_curve_key>() + k0, c->pend_keys.data(), c->pend_keys.size() * sizeof(fyx_curve_key), cudaMemcpyHostToDevice));
    if (!c->pend_tracks.empty()) {
        CU(cudaMemcpy(c->b_anim_tracks.as<AnimTrackDev>() + t0, c->pend_tracks.data(), c->pend_tracks.size() * sizeof(AnimTrackDev), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->b_anim_bk.as<uint32_t>() + t0, c->pend_bk.data(), c->pend_bk.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemset(c->b_anim_hints.as<uint4>() + t0, 0, c->pend_tracks.size() * sizeof(uint4))); // HintContainer::default()
        CU(cudaMemset(c->b_anim_ok.as<uint32_t>() + t0, 0, c->pend_tracks.size() * 4));            // AnimationPose::default(): no values
    }
    CU(cudaMemcpy(c->b_anim_state.as<AnimStateDev>() + a0, c->pend_state.data(), c->pend_state.size() * sizeof(AnimStateDev), cudaMemcpyHostToDevice));
    std::vector<fyx_curve_key>().swap(c->pend_keys);
    std::vector<AnimTrackDev>().swap(c->pend_tracks);
    std::vector<uint32_t>().swap(c->pend_bk);
    std::vector<AnimStateDev>().swap(c->pend_state);
    anim_rebuild_arrays(c);
    return FYX_OK;
}

} // namespace

extern "C" int32_t fyx_anim_clear(fyx_ctx *c)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    anim_free(c);
    return FYX_OK;
}

extern "C" int32_t fyx_anim_add(fyx_ctx *c, const fyx_animation_desc *d, uint32_t *out_id)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!d || d->struct_size < sizeof(fyx_animation_desc)) return fail(c, FYX_ERR_INVALID_ARGUMENT, "fyx_animation_desc is NULL or struct_size too small");
    if ((d->n_tracks && !d->tracks) || (d->n_keys && !d->keys)) return fail(c, FYX_ERR_INVALID_ARGUMENT, "tracks / keys is NULL");
    if (!(d->time_slice_start <= d->time_slice_end)) return fail(c, FYX_ERR_INVALID_ARGUMENT, "time slice start > end"); // lib.rs:446 asserts
    for (uint32_t i = 0; i < d->n_tracks; ++i) {
        const fyx_anim_track &t = d->tracks[i];
        if (t.binding > FYX_BIND_ROTATION) return fail(c, FYX_ERR_UNSUPPORTED, "track %u: property bindings are not supported", i);
        if (t.value_kind > FYX_TV_QUAT || t.n_curves > 4) return fail(c, FYX_ERR_INVALID_ARGUMENT, "track %u: bad value kind / curve count", i);
        for (uint32_t k = 0; k < t.n_curves; ++k)
            if ((uint64_t)t.first_key[k] + t.n_keys[k] > d->n_keys) return fail(c, FYX_ERR_INVALID_ARGUMENT, "track %u curve %u: keys out of range", i, k);
    }
    for (uint32_t k = 0; k < d->n_keys; ++k)
        if (d->keys[k].kind > FYX_KEY_CUBIC) return fail(c, FYX_ERR_INVALID_ARGUMENT, "key %u: bad kind", k);
    const uint32_t id = (uint32_t)c->anims.size();
    const uint32_t t0 = (uint32_t)c->anim_tracks.size(), k0 = c->n_anim_keys;
    // uploads are deferred to the next call that needs the device tables (anim_flush): thousands of animations
    // (one per skeleton) arrive back to back at load time
    for (uint32_t i = 0; i < d->n_tracks; ++i) {
        const fyx_anim_track &t = d->tracks[i];
        AnimTrackDev o{};
        o.anim = id;
        o.value_kind = t.value_kind;
        o.enabled = t.enabled ? 1u : 0u;
        o.n_curves = t.n_curves;
        for (int k = 0; k < 4; ++k) {
            o.first_key[k] = (k < (int)t.n_curves) ? k0 + t.first_key[k] : 0u;
            o.n_keys[k] = (k < (int)t.n_curves) ? t.n_keys[k] : 0u;
            if (o.n_keys[k]) {
                o.first_loc[k] = d->keys[t.first_key[k]].location;
                o.last_loc[k] = d->keys[t.first_key[k] + t.n_keys[k] - 1].location;
            }
        }
        c->pend_tracks.push_back(o);
        c->pend_bk.push_back(t.binding | (t.value_kind << 8));
    }
    c->pend_keys.insert(c->pend_keys.end(), d->keys, d->keys + d->n_keys);
    AnimStateDev st{};
    st.speed = d->speed;
    st.slice_start = d->time_slice_start;
    st.slice_end = d->time_slice_end;
    st.looped = d->looped ? 1u : 0u;
    st.enabled = d->enabled ? 1u : 0u;
    st.time = host_time_position(st, d->time_position);
    c->pend_state.push_back(st);
    AnimHost h;
    h.first_track = t0;
    h.n_tracks = d->n_tracks;
    h.st = st;
    c->anims.push_back(h);
    c->anim_tracks.insert(c->anim_tracks.end(), d->tracks, d->tracks + d->n_tracks);
    c->n_anim_keys = k0 + d->n_keys;
    c->anim_csr_dirty = true;
    if (out_id) *out_id = id;
    return FYX_OK;
}

#define ANIM_CHECK(id)                                                                          \
    if (!c) return FYX_ERR_INVALID_ARGUMENT;                                                    \
    if ((id) >= c->anims.size()) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation %u does not exist", (id)); \
    CU(cudaSetDevice(c->device));                                                               \
    { int32_t rc__ = anim_flush(c); if (rc__) return rc__; }

extern "C" int32_t fyx_anim_set_enabled(fyx_ctx *c, uint32_t anim, uint32_t enabled)
{
    ANIM_CHECK(anim)
    const uint32_t v = enabled ? 1u : 0u;
    c->anims[anim].st.enabled = v;
    CU(cudaMemcpyAsync(&c->b_anim_state.as<AnimStateDev>()[anim].enabled, &v, 4, cudaMemcpyHostToDevice, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_anim_set_track_enabled(fyx_ctx *c, uint32_t anim, uint32_t track, uint32_t enabled)
{
    ANIM_CHECK(anim)
    if (track >= c->anims[anim].n_tracks) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation %u has no track %u", anim, track);
    const uint32_t v = enabled ? 1u : 0u, gi = c->anims[anim].first_track + track;
    c->anim_tracks[gi].enabled = v;
    CU(cudaMemcpyAsync(&c->b_anim_tracks.as<AnimTrackDev>()[gi].enabled, &v, 4, cudaMemcpyHostToDevice, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_anim_set_speed(fyx_ctx *c, uint32_t anim, float speed)
{
    ANIM_CHECK(anim)
    c->anims[anim].st.speed = speed;
    CU(cudaMemcpyAsync(&c->b_anim_state.as<AnimStateDev>()[anim].speed, &speed, 4, cudaMemcpyHostToDevice, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_anim_set_time_position(fyx_ctx *c, uint32_t anim, float time)
{
    ANIM_CHECK(anim)
    const float t = host_time_position(c->anims[anim].st, time);
    CU(cudaMemcpyAsync(&c->b_anim_state.as<AnimStateDev>()[anim].time, &t, 4, cudaMemcpyHostToDevice, c->stream));
    return FYX_OK;
}

extern "C" int32_t fyx_anim_get_time_positions(fyx_ctx *c, uint32_t first, uint32_t count, float *out)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!count) return FYX_OK;
    if (!out || (uint64_t)first + count > c->anims.size()) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation range out of bounds");
    CU(cudaSetDevice(c->device));
    int32_t rc = anim_flush(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy2D(out, sizeof(float), &c->b_anim_state.as<AnimStateDev>()[first].time, sizeof(AnimStateDev), sizeof(float), count,
                    cudaMemcpyDeviceToHost));
    return FYX_OK;
}

static int32_t animate_enqueue(fyx_ctx *c, float dt)
{
    if (c->anims.empty()) return FYX_OK;
    int32_t rc = ensure_trs_store(c);
    if (rc) return rc;
    if ((rc = anim_flush(c))) return rc;
    if (c->anim_csr_dirty && (rc = anim_build_csr(c))) return rc;
    launch_animate(c->stream, c->a, c->an, c->b_trs.as<fyx_trs>(), c->have_statics ? c->b_statics.as<fyx_transform_statics>() : nullptr, dt,
                   c->d_err);
    c->launches += (c->an.n_tracks ? 1 : 0) + (c->an.n_nodes ? 1 : 0) + (c->an.n_anims ? 1 : 0);
    CU(cudaGetLastError());
    return FYX_OK;
}

extern "C" int32_t fyx_animate(fyx_ctx *c, float dt)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!c->have_topology) return fail(c, FYX_ERR_STATE, "fyx_set_topology has not been called");
    CU(cudaSetDevice(c->device));
    return animate_enqueue(c, dt);
}

// A blend group: the animations stop being applied directly (AnimationPlayer::auto_apply = false) and become the
// PlayAnimation sources of one BlendAnimations pose node with constant weights (machine/node/blend.rs:136-166).
extern "C" int32_t fyx_anim_blend_group(fyx_ctx *c, uint32_t n, const uint32_t *anims, const float *weights, uint32_t *out_group)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!n || !anims || !weights) return fail(c, FYX_ERR_INVALID_ARGUMENT, "empty blend group");
    for (uint32_t i = 0; i < n; ++i) {
        if (anims[i] >= c->anims.size()) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation %u does not exist", anims[i]);
        if (c->anims[anims[i]].st.group) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation %u already belongs to blend group %u", anims[i], c->anims[anims[i]].st.group);
        for (uint32_t j = 0; j < i; ++j)
            if (anims[j] == anims[i]) return fail(c, FYX_ERR_INVALID_ARGUMENT, "animation %u listed twice", anims[i]);
        // inside a group every (node, property) has one track per animation, of a Vector3 / UnitQuaternion kind
        const AnimHost &h = c->anims[anims[i]];
        std::vector<uint64_t> seen;
        seen.reserve(h.n_tracks);
        for (uint32_t t = 0; t < h.n_tracks; ++t) {
            const fyx_anim_track &tr = c->anim_tracks[h.first_track + t];
            if (tr.value_kind != FYX_TV_VECTOR3 && tr.value_kind != FYX_TV_QUAT && tr.value_kind != FYX_TV_QUAT_EULER)
                return fail(c, FYX_ERR_UNSUPPORTED, "animation %u track %u: only Vector3 / UnitQuaternion tracks can be blended", anims[i], t);
            seen.push_back(((uint64_t)tr.target_node << 2) | tr.binding);
        }
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end())
            return fail(c, FYX_ERR_UNSUPPORTED, "animation %u animates one property of one node with two tracks", anims[i]);
    }
    CU(cudaSetDevice(c->device));
    int32_t rc = anim_flush(c);
    if (rc) return rc;
    const uint32_t g = ++c->n_blend_groups;
    for (uint32_t i = 0; i < n; ++i) {
        AnimHost &h = c->anims[anims[i]];
        h.st.group = g;
        h.st.weight = weights[i];
        h.pos_in_group = i;
        struct { uint32_t group; float weight; } gw = {g, weights[i]};
        CU(cudaMemcpyAsync(&c->b_anim_state.as<AnimStateDev>()[anims[i]].group, &gw, 8, cudaMemcpyHostToDevice, c->stream));
    }
    c->blend_groups.emplace_back(anims, anims + n);
    c->anim_csr_dirty = true;
    if (out_group) *out_group = g;
    return FYX_OK;
}

// PoseWeight::Constant values of a group's sources
extern "C" int32_t fyx_anim_set_blend_weights(fyx_ctx *c, uint32_t group, uint32_t n, const float *weights)
{
    if (!c) return FYX_ERR_INVALID_ARGUMENT;
    if (!group || group > c->blend_groups.size()) return fail(c, FYX_ERR_INVALID_ARGUMENT, "blend group %u does not exist", group);
    const std::vector<uint32_t> &src = c->blend_groups[group - 1];
    if (n != src.size() || !weights) return fail(c, FYX_ERR_INVALID_ARGUMENT, "blend group %u has %zu sources", group, src.size());
    CU(cudaSetDevice(c->device));
    for (uint32_t i = 0; i < n; ++i) {
        c->anims[src[i]].st.weight = weights[i];
        CU(cudaMemcpyAsync(&c->b_anim_state.as<AnimStateDev>()[src[i]].weight, &weights[i], 4, cudaMemcpyHostToDevice, c->stream));
    }
    return FYX_OK;
}
