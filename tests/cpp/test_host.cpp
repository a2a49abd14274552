// The reference's own hierarchy tests (fyrox-impl/src/scene/graph/mod.rs:2602-2739) and a from_graph cull,
// transcribed onto the C++ host mirror (fyrox_b200/host/fyrox_host.hpp).  Prints "OK" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>

#include "../../fyrox_b200/host/fyrox_host.hpp"

using namespace fyrox;

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                        \
        }                                                                        \
    } while (0)

static bool eq(const Vec3 &a, float x, float y, float z) { return a[0] == x && a[1] == y && a[2] == z; }

static Mat4 look_at_rh_neg_z() // camera at the origin looking down -Z: the identity view
{
    return Mat4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
}
static Mat4 perspective(float aspect, float fovy, float znear, float zfar) // nalgebra Perspective3::new
{
    Mat4 m{};
    const float m22 = 1.0f / std::tan(fovy / 2.0f);
    m[5] = m22;
    m[0] = m22 / aspect;
    m[10] = (zfar + znear) / (znear - zfar);
    m[14] = zfar * znear * 2.0f / (znear - zfar);
    m[11] = -1.0f;
    return m;
}

int main()
{
    try {
        { // test_hierarchy_changes_propagation, graph/mod.rs:2646-2739
            Graph graph;
            Handle c = BaseBuilder().with_local_transform(TransformBuilder().with_local_position({{0, 0, 1}}).build()).build_pivot(graph);
            Handle b = BaseBuilder().with_visibility(false).with_enabled(false)
                           .with_local_transform(TransformBuilder().with_local_position({{0, 1, 0}}).build()).with_child(c).build_pivot(graph);
            Handle d = BaseBuilder().with_local_transform(TransformBuilder().with_local_position({{1, 1, 1}}).build()).build_pivot(graph);
            Handle a = BaseBuilder().with_local_transform(TransformBuilder().with_local_position({{1, 0, 0}}).build()).with_child(b).with_child(d).build_pivot(graph);
            CHECK(graph.root().index == 0 && graph.root().generation == 1 && c.index == 1 && c.generation == 1); // K10
            graph.update();
            CHECK(eq(graph[a].global_position(), 1, 0, 0));
            CHECK(eq(graph[b].global_position(), 1, 1, 0));
            CHECK(eq(graph[c].global_position(), 1, 1, 1));
            CHECK(eq(graph[d].global_position(), 2, 1, 1));
            CHECK(graph[a].global_visibility() && !graph[b].global_visibility() && !graph[c].global_visibility() && graph[d].global_visibility());
            CHECK(graph[a].is_globally_enabled() && !graph[b].is_globally_enabled() && !graph[c].is_globally_enabled() && graph[d].is_globally_enabled());
            graph[b].local_transform_mut().local_position = {{0, 2, 0}};
            graph[a].set_enabled(false);
            graph[b].set_visibility(true);
            graph.update();
            CHECK(eq(graph[a].global_position(), 1, 0, 0));
            CHECK(eq(graph[b].global_position(), 1, 2, 0));
            CHECK(eq(graph[c].global_position(), 1, 2, 1));
            CHECK(eq(graph[d].global_position(), 2, 1, 1));
            for (Handle h : {a, b, c, d}) CHECK(graph[h].global_visibility() && !graph[h].is_globally_enabled());
        }
        { // static batching: a rendered BatchingMode::Static mesh returns RdcControlFlow::Break (scene/mesh/mod.rs:701-725)
            Graph graph;
            const AxisAlignedBoundingBox box = AxisAlignedBoundingBox::from_min_max({{-1, -1, -1}}, {{1, 1, 1}});
            auto at = [](float x, float y, float z) { return TransformBuilder().with_local_position({{x, y, z}}).build(); };
            Handle child_a = BaseBuilder().with_local_bounding_box(box).with_local_transform(at(1, 0, 0)).build_mesh(graph);
            Handle parent_a = BaseBuilder().with_local_bounding_box(box).with_local_transform(at(0, 0, -10)).with_child(child_a).build_mesh(graph);
            Handle child_b = BaseBuilder().with_local_bounding_box(box).with_local_transform(at(0, 0, -70)).build_mesh(graph);
            Handle parent_b = BaseBuilder().with_local_bounding_box(box).with_local_transform(at(0, 0, 50)).with_child(child_b).build_mesh(graph);
            graph.update();
            ObserverPosition op;
            op.view_matrix = look_at_rh_neg_z();
            op.projection_matrix = perspective(16.0f / 9.0f, 60.0f * 3.14159265358979f / 180.0f, 0.1f, 150.0f);
            auto vis = [&]() {
                std::set<uint32_t> v;
                for (Handle h : RenderDataBundleStorage::from_graph(graph, 0xFFFFFFFFu, 0.0f, op, "GBuffer").visible_handles) v.insert(h.index);
                return v;
            };
            CHECK((vis() == std::set<uint32_t>{parent_a.index, child_a.index, child_b.index}));
            graph[parent_a].set_batching_mode(BatchingMode::Static);
            graph[parent_b].set_batching_mode(BatchingMode::Static);
            graph.update();
            CHECK((vis() == std::set<uint32_t>{parent_a.index, child_b.index})); // a rendered batch hides its child, a culled one does not
            graph[parent_a].set_batching_mode(BatchingMode::None);
            graph.update();
            CHECK((vis() == std::set<uint32_t>{parent_a.index, child_a.index, child_b.index}));
        }
        { // test_global_scale, graph/mod.rs:2602-2644 + a cull through from_graph
            Graph graph;
            Handle c = BaseBuilder().with_local_transform(TransformBuilder().with_local_scale({{1, 2, 3}}).build()).build_pivot(graph);
            Handle b = BaseBuilder().with_local_transform(TransformBuilder().with_local_scale({{3, 2, 1}}).build()).with_child(c).build_pivot(graph);
            Handle a = BaseBuilder().with_local_transform(TransformBuilder().with_local_scale({{1, 1, 2}}).build()).with_child(b).build_pivot(graph);
            CHECK(eq(graph.global_scale(a), 1, 1, 2) && eq(graph.global_scale(b), 3, 2, 2) && eq(graph.global_scale(c), 3, 4, 6));
            const AxisAlignedBoundingBox box = AxisAlignedBoundingBox::from_min_max({{-1, -1, -1}}, {{1, 1, 1}});
            Handle near_ = BaseBuilder().with_local_bounding_box(box).with_local_transform(TransformBuilder().with_local_position({{0, 0, -10}}).build()).build_mesh(graph);
            Handle behind = BaseBuilder().with_local_bounding_box(box).with_local_transform(TransformBuilder().with_local_position({{0, 0, 50}}).build()).build_mesh(graph);
            Handle noshadow = BaseBuilder().with_cast_shadows(false).with_local_bounding_box(box)
                                  .with_local_transform(TransformBuilder().with_local_position({{2, 0, -10}}).build()).build_mesh(graph);
            graph.update();
            ObserverPosition op;
            op.view_matrix = look_at_rh_neg_z();
            op.projection_matrix = perspective(16.0f / 9.0f, 60.0f * 3.14159265358979f / 180.0f, 0.1f, 150.0f);
            auto idx_set = [](const RenderDataBundleStorage &s) {
                std::set<uint32_t> r;
                for (Handle h : s.visible_handles) r.insert(h.index);
                return r;
            };
            const auto main_pass = idx_set(RenderDataBundleStorage::from_graph(graph, 0xFFFFFFFFu, 0.0f, op, "GBuffer"));
            CHECK((main_pass == std::set<uint32_t>{near_.index, noshadow.index}));
            const auto shadow = idx_set(RenderDataBundleStorage::from_graph(graph, 0xFFFFFFFFu, 0.0f, op, "SpotShadow"));
            CHECK((shadow == std::set<uint32_t>{near_.index}));
            CHECK(!main_pass.count(behind.index));
            // the world box of a Base node is the unit box moved by its global transform (scene/base.rs:741-750)
            const AxisAlignedBoundingBox wb = graph[near_].world_bounding_box();
            CHECK(wb.min[2] == -11.0f && wb.max[2] == -9.0f);
            graph.remove_node(noshadow);
            graph.update();
            const auto after = idx_set(RenderDataBundleStorage::from_graph(graph, 0xFFFFFFFFu, 0.0f, op, "GBuffer"));
            CHECK((after == std::set<uint32_t>{near_.index}));
        }
        std::puts("OK");
        return 0;
    } catch (const Error &e) {
        std::fprintf(stderr, "fyrox::Error %d: %s\n", e.code, e.what());
        return 2;
    }
}
