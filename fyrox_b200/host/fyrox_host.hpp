// fyrox_host.hpp — C++ host-side mirror of the reference's interface for the render-prep path, on top of
// the C ABI (include/fyrox_b200.h).  The reference is compiled code (Rust); this image has no Rust
// toolchain, so the host layer a Rust shim would be is written in C++ with the reference's names and
// argument meaning (header-only, C++17):
//
//   Handle                         fyrox-core/src/pool/handle.rs:38-47
//   Transform / TransformBuilder   fyrox-impl/src/scene/transform.rs:79-127,421-550
//   Node (Base + Mesh bits)        fyrox-impl/src/scene/base.rs:389-483, scene/mesh/mod.rs:328-377
//   Graph::{add_node,link_nodes,remove_node,update,update_hierarchical_data,global_scale}
//                                  fyrox-impl/src/scene/graph/mod.rs:408-424,1272-1292,1459-1504,1835-1845,2044-2160
//   ObserverPosition               fyrox-impl/src/renderer/observer.rs:47-60
//   RenderDataBundleStorage::from_graph   fyrox-impl/src/renderer/bundle.rs:873-1009
//
// No arithmetic of the hot path happens here: Graph::update scatters what changed through
// fyx_set_local_trs / fyx_set_flags (the device evaluates Transform::calculate_local_transform) and runs
// the sm_100a kernels; queries read back.  Errors throw fyrox::Error carrying fyx_last_error().
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <limits>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fyrox_b200.h"

namespace fyrox {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct Handle {
    uint32_t index = 0, generation = 0;
    bool is_none() const { return generation == 0; }
    bool is_some() const { return generation != 0; }
    bool operator==(const Handle &o) const { return index == o.index && generation == o.generation; }
    bool operator!=(const Handle &o) const { return !(*this == o); }
};
static const Handle HANDLE_NONE{};

using Vec3 = std::array<float, 3>;
using Mat4 = std::array<float, 16>; // column-major, nalgebra's Matrix4<f32> layout

struct AxisAlignedBoundingBox {
    Vec3 min{{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()}};
    Vec3 max{{-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()}};
    static AxisAlignedBoundingBox unit() { return from_min_max({{-0.5f, -0.5f, -0.5f}}, {{0.5f, 0.5f, 0.5f}}); }
    static AxisAlignedBoundingBox from_min_max(Vec3 a, Vec3 b)
    {
        AxisAlignedBoundingBox r;
        r.min = a;
        r.max = b;
        return r;
    }
};

// Transform (scene/transform.rs:79-127): what animation / game code writes; the matrix is computed on the device.
struct Transform {
    Vec3 local_position{{0, 0, 0}};
    std::array<float, 4> local_rotation{{0, 0, 0, 1}}; // i, j, k, w
    Vec3 local_scale{{1, 1, 1}};
    std::array<float, 4> pre_rotation{{0, 0, 0, 1}};
    std::array<float, 9> post_rotation_matrix{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    Vec3 rotation_offset{{0, 0, 0}}, rotation_pivot{{0, 0, 0}}, scaling_offset{{0, 0, 0}}, scaling_pivot{{0, 0, 0}};
    bool has_default_statics() const
    {
        const Transform d;
        return pre_rotation == d.pre_rotation && post_rotation_matrix == d.post_rotation_matrix && rotation_offset == d.rotation_offset &&
               rotation_pivot == d.rotation_pivot && scaling_offset == d.scaling_offset && scaling_pivot == d.scaling_pivot;
    }
};

struct TransformBuilder {
    Transform t;
    TransformBuilder &with_local_position(Vec3 v) { t.local_position = v; return *this; }
    TransformBuilder &with_local_rotation(std::array<float, 4> q) { t.local_rotation = q; return *this; }
    TransformBuilder &with_local_scale(Vec3 v) { t.local_scale = v; return *this; }
    Transform build() const { return t; }
};

// BlendShape / BlendShapesContainer (scene/mesh/surface.rs:71-218): weight in 0..100 and the texels of the volume texture
// from_lists builds (n_shapes layers x layer_stride records of 9 binary16 values: position, normal, tangent offsets)
struct BlendShape {
    float weight = 100.0f;
    std::string name;
};
struct BlendShapesContainer {
    std::vector<BlendShape> blend_shapes;
    std::vector<uint16_t> blend_shape_storage;
    uint32_t layer_stride = 0;
};

struct Surface {
    std::vector<Handle> bones;
    std::vector<uint8_t> vertex_buffer; // AnimatedVertex records (68 B)
    int64_t surface_id = -1;
    BlendShapesContainer blend_shapes_container; // SurfaceData::blend_shapes_container
    bool shapes_uploaded = false;
};

enum class NodeKind { Pivot, Mesh };
enum class BatchingMode { None, Static, Dynamic }; // scene/mesh/mod.rs:120-140

class Graph;

class Node {
  public:
    NodeKind kind = NodeKind::Pivot;
    bool frustum_culling = true, cast_shadows = true;
    uint32_t render_mask = 0xFFFFFFFFu;
    Handle parent;
    std::vector<Handle> children;
    AxisAlignedBoundingBox local_bounding_box = AxisAlignedBoundingBox::unit();
    Mat4 inv_bind_pose_transform{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
    std::vector<Surface> surfaces;

    // Mesh::batching_mode / set_batching_mode (scene/mesh/mod.rs:372, 612-617): a Static mesh that is rendered returns
    // RdcControlFlow::Break and the DFS of from_graph does not visit its children (:701-725)
    BatchingMode batching_mode() const { return batching_mode_; }
    void set_batching_mode(BatchingMode m);
    // Mesh::blend_shapes_mut (scene/mesh/mod.rs:453-456): the weights (0..100) of the first surface that has shapes
    std::vector<BlendShape> &blend_shapes_mut();

    const Transform &local_transform() const { return transform_; }
    Transform &local_transform_mut();          // marks TransformChanged (scene/base.rs:343-352)
    bool visibility() const { return visibility_; }
    bool is_enabled() const { return enabled_; }
    void set_visibility(bool v);
    void set_enabled(bool v);

    Vec3 global_position() const;
    Mat4 global_transform() const;
    bool global_visibility() const;
    bool is_globally_enabled() const;
    AxisAlignedBoundingBox world_bounding_box() const;

  private:
    friend class Graph;
    friend struct BaseBuilder;
    Transform transform_;
    bool visibility_ = true, enabled_ = true;
    BatchingMode batching_mode_ = BatchingMode::None;
    Graph *graph_ = nullptr;
    Handle self_;
    uint32_t flags_word() const
    {
        return FYX_NODE_ALIVE | (visibility_ ? FYX_NODE_VISIBILITY : 0u) | (enabled_ ? FYX_NODE_ENABLED : 0u) |
               (frustum_culling ? FYX_NODE_FRUSTUM_CULLING : 0u) | (cast_shadows ? FYX_NODE_CAST_SHADOWS : 0u) |
               (kind == NodeKind::Mesh ? FYX_NODE_RENDERABLE : 0u) |
               ((kind == NodeKind::Mesh && batching_mode_ == BatchingMode::Static) ? FYX_NODE_STATIC_BATCH : 0u);
    }
};

struct ObserverPosition {
    Vec3 translation{{0, 0, 0}};
    float z_near = 0.1f, z_far = 100.0f;
    Mat4 view_matrix, projection_matrix;
};

inline bool is_shadow_pass(const std::string &name) { return name == "DirectionalShadow" || name == "SpotShadow" || name == "PointShadow"; }

class Graph {
  public:
    explicit Graph(int device = -1)
    {
        fyx_config cfg{};
        cfg.struct_size = sizeof cfg;
        cfg.device = device;
        const int32_t rc = fyx_create(&cfg, &ctx_);
        if (rc) throw Error(rc, fyx_last_error(nullptr));
        Node root;
        root_ = add_node(std::move(root)); // Graph::new, graph/mod.rs:408-424
    }
    ~Graph() { fyx_destroy(ctx_); }
    Graph(const Graph &) = delete;
    Graph &operator=(const Graph &) = delete;

    Handle root() const { return root_; }
    fyx_ctx *context() { return ctx_; }
    uint32_t capacity() const { return (uint32_t)records_.size(); }
    bool is_valid_handle(Handle h) const { return h.is_some() && h.index < records_.size() && alive_[h.index] && generation_[h.index] == h.generation; }
    Node *try_get_node(Handle h) { return is_valid_handle(h) ? &records_[h.index] : nullptr; }
    Node &operator[](Handle h)
    {
        if (!is_valid_handle(h)) throw Error(FYX_ERR_INVALID_ARGUMENT, "invalid handle");
        return records_[h.index];
    }

    Handle add_node(Node node) // graph/mod.rs:2044-2088
    {
        uint32_t i;
        if (!free_.empty()) {
            i = free_.back();
            free_.pop_back();
            generation_[i] += 1;
            records_[i] = std::move(node);
            alive_[i] = true;
        } else {
            i = (uint32_t)records_.size();
            records_.push_back(std::move(node));
            generation_.push_back(1);
            alive_.push_back(true);
        }
        Handle h{i, generation_[i]};
        const std::vector<Handle> kids = records_[i].children;
        records_[i].children.clear();
        records_[i].parent = HANDLE_NONE;
        records_[i].graph_ = this;
        records_[i].self_ = h;
        topology_dirty_ = true;
        if (root_.is_some()) link_nodes(h, root_);
        for (Handle c : kids) link_nodes(c, h);
        return h;
    }

    void link_nodes(Handle child, Handle parent) // graph/mod.rs:2114-2131
    {
        isolate(child);
        (*this)[child].parent = parent;
        (*this)[parent].children.push_back(child);
        topology_dirty_ = true;
    }

    void remove_node(Handle h) // graph/mod.rs:2091-2111
    {
        isolate(h);
        std::vector<Handle> stack{h};
        while (!stack.empty()) {
            Handle x = stack.back();
            stack.pop_back();
            Node *n = try_get_node(x);
            if (!n) continue;
            for (Handle c : n->children) stack.push_back(c);
            alive_[x.index] = false;
            free_.push_back(x.index);
        }
        topology_dirty_ = true;
    }

    // Graph::update reduced to its hierarchical part (process_node_messages, graph/mod.rs:1303-1399,1459-1473)
    void update()
    {
        sync();
        check(fyx_update_transforms(ctx_, FYX_UPDATE_INCREMENTAL));
    }
    void update_hierarchical_data() // graph/mod.rs:1272-1292
    {
        sync();
        check(fyx_update_transforms(ctx_, FYX_UPDATE_ALL));
    }

    Vec3 global_scale(Handle h) // graph/mod.rs:1835-1845
    {
        Vec3 s{{1, 1, 1}};
        Node *n = try_get_node(h);
        while (n) {
            for (int k = 0; k < 3; ++k) s[k] = s[k] * n->local_transform().local_scale[k];
            n = try_get_node(n->parent);
        }
        return s;
    }

    Mat4 global_transform(Handle h)
    {
        Mat4 m;
        const uint32_t i = h.index;
        check(fyx_get_global_matrices(ctx_, 1, &i, m.data()));
        return m;
    }
    uint32_t global_flags(Handle h)
    {
        uint32_t f = 0;
        const uint32_t i = h.index;
        check(fyx_get_global_flags(ctx_, 1, &i, &f));
        return f;
    }
    AxisAlignedBoundingBox world_bounding_box(Handle h)
    {
        float b[6];
        const uint32_t i = h.index;
        check(fyx_get_world_aabbs(ctx_, 1, &i, b));
        return AxisAlignedBoundingBox::from_min_max({{b[0], b[1], b[2]}}, {{b[3], b[4], b[5]}});
    }

    void check(int32_t rc)
    {
        if (rc) throw Error(rc, fyx_last_error(ctx_));
    }
    Handle handle_from_index(uint32_t i) const { return (i < records_.size() && alive_[i]) ? Handle{i, generation_[i]} : HANDLE_NONE; }

  private:
    friend class Node;
    void isolate(Handle h) // graph/mod.rs:2143-2160
    {
        Node &n = (*this)[h];
        Node *p = try_get_node(n.parent);
        n.parent = HANDLE_NONE;
        if (p)
            for (size_t k = 0; k < p->children.size(); ++k)
                if (p->children[k] == h) {
                    p->children.erase(p->children.begin() + (long)k);
                    break;
                }
    }
    static fyx_trs to_trs(const Transform &t)
    {
        fyx_trs r;
        memcpy(r.position, t.local_position.data(), 12);
        memcpy(r.rotation, t.local_rotation.data(), 16);
        memcpy(r.scale, t.local_scale.data(), 12);
        return r;
    }
    static fyx_transform_statics to_statics(const Transform &t)
    {
        fyx_transform_statics s;
        memcpy(s.pre_rotation, t.pre_rotation.data(), 16);
        memcpy(s.post_rotation_matrix, t.post_rotation_matrix.data(), 36);
        memcpy(s.rotation_offset, t.rotation_offset.data(), 12);
        memcpy(s.rotation_pivot, t.rotation_pivot.data(), 12);
        memcpy(s.scaling_offset, t.scaling_offset.data(), 12);
        memcpy(s.scaling_pivot, t.scaling_pivot.data(), 12);
        return s;
    }
    void sync()
    {
        const uint32_t cap = capacity();
        if (topology_dirty_) {
            std::vector<uint32_t> parent(cap, FYX_NONE), flags(cap, 0), mask(cap, 0), sidx, tidx;
            std::vector<float> aabb((size_t)cap * 6, 0.f);
            std::vector<fyx_trs> trs;
            std::vector<fyx_transform_statics> statics;
            for (uint32_t i = 0; i < cap; ++i) {
                if (!alive_[i]) continue;
                Node &n = records_[i];
                parent[i] = is_valid_handle(n.parent) ? n.parent.index : FYX_NONE;
                flags[i] = n.flags_word();
                mask[i] = n.render_mask;
                memcpy(&aabb[(size_t)i * 6], n.local_bounding_box.min.data(), 12);
                memcpy(&aabb[(size_t)i * 6 + 3], n.local_bounding_box.max.data(), 12);
                tidx.push_back(i);
                trs.push_back(to_trs(n.transform_));
                if (!n.transform_.has_default_statics()) {
                    sidx.push_back(i);
                    statics.push_back(to_statics(n.transform_));
                }
            }
            check(fyx_set_topology(ctx_, cap, root_.index, parent.data(), flags.data(), mask.data(), aabb.data(), nullptr));
            if (!sidx.empty()) check(fyx_set_transform_statics(ctx_, (uint32_t)sidx.size(), sidx.data(), statics.data()));
            check(fyx_set_local_trs(ctx_, (uint32_t)tidx.size(), tidx.data(), trs.data()));
            upload_surfaces();
            topology_dirty_ = false;
            dirty_transform_.clear();
            dirty_flags_.clear();
            return;
        }
        if (!dirty_transform_.empty()) {
            std::vector<uint32_t> idx, sidx;
            std::vector<fyx_trs> trs;
            std::vector<fyx_transform_statics> statics;
            for (uint32_t i : dirty_transform_) {
                if (i >= cap || !alive_[i]) continue;
                idx.push_back(i);
                trs.push_back(to_trs(records_[i].transform_));
                if (!records_[i].transform_.has_default_statics()) {
                    sidx.push_back(i);
                    statics.push_back(to_statics(records_[i].transform_));
                }
            }
            if (!sidx.empty()) check(fyx_set_transform_statics(ctx_, (uint32_t)sidx.size(), sidx.data(), statics.data()));
            check(fyx_set_local_trs(ctx_, (uint32_t)idx.size(), idx.data(), trs.data()));
            dirty_transform_.clear();
        }
        if (!dirty_flags_.empty()) {
            std::vector<uint32_t> idx, fl;
            for (uint32_t i : dirty_flags_) {
                if (i >= cap || !alive_[i]) continue;
                idx.push_back(i);
                fl.push_back(records_[i].flags_word());
            }
            check(fyx_set_flags(ctx_, (uint32_t)idx.size(), idx.data(), fl.data()));
            dirty_flags_.clear();
        }
        upload_surfaces();
    }
    void upload_surfaces()
    {
        static const fyx_vertex_layout animated{68, 0, 20, 48, 64}; // scene/mesh/vertex.rs:140-210
        for (uint32_t i = 0; i < capacity(); ++i) {
            if (!alive_[i]) continue;
            for (Surface &s : records_[i].surfaces) {
                if (s.surface_id >= 0 || s.bones.empty()) continue;
                std::vector<uint32_t> bones;
                std::vector<float> ib;
                for (Handle b : s.bones) {
                    bones.push_back(is_valid_handle(b) ? b.index : FYX_NONE);
                    const Mat4 &m = is_valid_handle(b) ? records_[b.index].inv_bind_pose_transform : Node().inv_bind_pose_transform;
                    ib.insert(ib.end(), m.begin(), m.end());
                }
                uint32_t sid = 0;
                const uint32_t nv = (uint32_t)(s.vertex_buffer.size() / 68);
                check(fyx_add_skinned_surface(ctx_, i, (uint32_t)bones.size(), bones.data(), ib.data(), nv, nv ? s.vertex_buffer.data() : nullptr,
                                              nv ? &animated : nullptr, &sid));
                s.surface_id = sid;
            }
            // SurfaceData::blend_shapes_container -> fyx_set_blend_shapes once, Mesh::blend_shapes weights when they were touched
            for (Surface &s : records_[i].surfaces) {
                BlendShapesContainer &bc = s.blend_shapes_container;
                if (s.surface_id < 0 || bc.blend_shapes.empty() || bc.blend_shape_storage.empty()) continue;
                std::vector<float> w;
                for (const BlendShape &b : bc.blend_shapes) w.push_back(b.weight);
                if (!s.shapes_uploaded) {
                    check(fyx_set_blend_shapes(ctx_, (uint32_t)s.surface_id, (uint32_t)w.size(), bc.blend_shape_storage.data(), bc.layer_stride, w.data()));
                    s.shapes_uploaded = true;
                } else if (dirty_shapes_.count(i)) {
                    check(fyx_set_blend_shape_weights(ctx_, (uint32_t)s.surface_id, (uint32_t)w.size(), w.data()));
                }
            }
        }
        dirty_shapes_.clear();
    }

    fyx_ctx *ctx_ = nullptr;
    std::vector<Node> records_;
    std::vector<uint32_t> generation_;
    std::vector<bool> alive_;
    std::vector<uint32_t> free_;
    Handle root_;
    bool topology_dirty_ = true;
    std::set<uint32_t> dirty_transform_, dirty_flags_, dirty_shapes_;
};

inline Transform &Node::local_transform_mut()
{
    if (graph_) graph_->dirty_transform_.insert(self_.index);
    return transform_;
}
inline void Node::set_visibility(bool v)
{
    visibility_ = v;
    if (graph_) graph_->dirty_flags_.insert(self_.index);
}
inline void Node::set_enabled(bool v)
{
    enabled_ = v;
    if (graph_) graph_->dirty_flags_.insert(self_.index);
}
inline void Node::set_batching_mode(BatchingMode m)
{
    batching_mode_ = m;
    if (graph_) graph_->dirty_flags_.insert(self_.index);
}
inline std::vector<BlendShape> &Node::blend_shapes_mut()
{
    if (graph_) graph_->dirty_shapes_.insert(self_.index);
    for (Surface &s : surfaces)
        if (!s.blend_shapes_container.blend_shapes.empty()) return s.blend_shapes_container.blend_shapes;
    static std::vector<BlendShape> none;
    return none;
}
inline Mat4 Node::global_transform() const { return graph_->global_transform(self_); }
inline Vec3 Node::global_position() const
{
    const Mat4 m = global_transform();
    return {{m[12], m[13], m[14]}};
}
inline bool Node::global_visibility() const { return (graph_->global_flags(self_) & FYX_NODE_GLOBAL_VISIBILITY) != 0; }
inline bool Node::is_globally_enabled() const { return (graph_->global_flags(self_) & FYX_NODE_GLOBAL_ENABLED) != 0; }
inline AxisAlignedBoundingBox Node::world_bounding_box() const { return graph_->world_bounding_box(self_); }

// BaseBuilder / PivotBuilder / MeshBuilder (scene/base.rs:1255-1400)
struct BaseBuilder {
    Node n;
    std::vector<Handle> kids;
    BaseBuilder &with_local_transform(const Transform &t) { n.transform_ = t; return *this; }
    BaseBuilder &with_visibility(bool v) { n.visibility_ = v; return *this; }
    BaseBuilder &with_enabled(bool v) { n.enabled_ = v; return *this; }
    BaseBuilder &with_cast_shadows(bool v) { n.cast_shadows = v; return *this; }
    BaseBuilder &with_frustum_culling(bool v) { n.frustum_culling = v; return *this; }
    BaseBuilder &with_render_mask(uint32_t m) { n.render_mask = m; return *this; }
    BaseBuilder &with_local_bounding_box(const AxisAlignedBoundingBox &b) { n.local_bounding_box = b; return *this; }
    BaseBuilder &with_child(Handle h) { kids.push_back(h); return *this; }
    Handle build_pivot(Graph &g)
    {
        n.kind = NodeKind::Pivot;
        n.children = kids;
        return g.add_node(std::move(n));
    }
    Handle build_mesh(Graph &g)
    {
        n.kind = NodeKind::Mesh;
        n.children = kids;
        return g.add_node(std::move(n));
    }
};

// The visible-node part of RenderDataBundleStorage (renderer/bundle.rs:873-1009)
struct RenderDataBundleStorage {
    std::vector<Handle> visible_handles;
    static RenderDataBundleStorage from_graph(Graph &graph, uint32_t render_mask, float /*elapsed_time*/, const ObserverPosition &op,
                                              const std::string &render_pass_name)
    {
        Mat4 vp;
        fyx_mat4_mul(op.projection_matrix.data(), op.view_matrix.data(), vp.data()); // bundle.rs:894
        fyx_frustum f;
        if (fyx_frustum_from_view_projection_matrix(vp.data(), &f) != FYX_OK) fyx_frustum_default(&f); // unwrap_or_default, :896
        const uint32_t pass = is_shadow_pass(render_pass_name) ? FYX_PASS_SHADOW : 0u;
        graph.check(fyx_cull(graph.context(), 1, &f, &render_mask, &pass));
        const uint32_t *idx = nullptr;
        uint32_t n = 0;
        graph.check(fyx_get_visible(graph.context(), 0, &idx, &n));
        RenderDataBundleStorage s;
        for (uint32_t k = 0; k < n; ++k) s.visible_handles.push_back(graph.handle_from_index(idx[k]));
        return s;
    }
};

} // namespace fyrox
