"""Pins the oracle (oracle/fyrox_oracle.c) against the reference's own unit-test vectors K1..K10
(tests/golden/reference_kats.json, transcribed from the cited Rust tests), and the product's
host-side math (fyx_frustum_from_view_projection_matrix / fyx_mat4_mul) against the same vectors.
CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
from oracle_binding import Aabb, Frustum, Plane, fp, vec

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
L = ob.lib()


def col_major(row_major16):
    return np.array(row_major16, dtype=np.float32).reshape(4, 4).T.reshape(16).copy()


def identity_frustum() -> Frustum:
    f = ob.frustum_from_vp(col_major(KATS["K3_frustum_from_identity"]["vp_row_major"]))
    assert f is not None
    return f


def test_k1_aabb_transform():
    k = KATS["K1_aabb_transform"]
    m = ob.mat4_mul(ob.translation(*k["translation"]), ob.scaling(*k["scaling"]))
    a = Aabb.make(k["aabb"][:3], k["aabb"][3:])
    out = Aabb()
    L.orc_aabb_transform(C.byref(a), fp(m), C.byref(out))
    assert out.to_np().tolist() == k["expected"]


def test_k2_aabb_basics():
    k = KATS["K2_aabb_basics"]
    a = Aabb()
    L.orc_aabb_default(C.byref(a))
    assert a.to_np().tolist() == np.array(k["default"], np.float32).tolist()
    assert not L.orc_aabb_is_valid(C.byref(a))
    L.orc_aabb_add_point(C.byref(a), fp(vec(1, 1, 1)))
    assert L.orc_aabb_is_valid(C.byref(a))
    L.orc_aabb_add_point(C.byref(a), fp(vec(-1, -1, -1)))
    assert L.orc_aabb_is_valid(C.byref(a))

    u = Aabb()
    L.orc_aabb_unit(C.byref(u))
    assert u.to_np().tolist() == k["unit"]
    assert not L.orc_aabb_is_degenerate(C.byref(u))
    col = Aabb.make([0, 0, 0], [0, 0, 0])
    assert L.orc_aabb_is_degenerate(C.byref(col))

    b = Aabb()
    L.orc_aabb_default(C.byref(b))
    for p in k["add_point"]["points"]:
        L.orc_aabb_add_point(C.byref(b), fp(vec(*p)))
    assert b.to_np().tolist() == k["add_point"]["expected"]

    c = Aabb.make(k["add_box"]["start"][:3], k["add_box"]["start"][3:])
    bx = Aabb.make(k["add_box"]["box"][:3], k["add_box"]["box"][3:])
    L.orc_aabb_add_box(C.byref(c), C.byref(bx))
    assert c.to_np().tolist() == k["add_box"]["expected"]

    r1 = Aabb.make([-1, -1, -1], [1, 1, 1])
    corners = np.empty((8, 3), np.float32)
    L.orc_aabb_corners(C.byref(r1), fp(corners.reshape(-1)))
    assert corners.tolist() == k["corners_of_radius_1"]
    # is_contains_point is inclusive on all corners of the unit box (aabb.rs:511-519)
    L.orc_aabb_corners(C.byref(u), fp(corners.reshape(-1)))
    assert L.orc_aabb_is_contains_point(C.byref(u), fp(vec(0, 0, 0)))
    for p in corners:
        assert L.orc_aabb_is_contains_point(C.byref(u), fp(np.ascontiguousarray(p)))


def test_k3_frustum_from_identity():
    k = KATS["K3_frustum_from_identity"]
    f = identity_frustum()
    planes, corners = ob.frustum_planes_corners(f)
    for i, (a, b, c, d) in enumerate(k["planes_abcd"]):
        p = Plane()
        assert L.orc_plane_from_abcd(a, b, c, d, C.byref(p))
        assert planes[i].tolist() == [p.n[0], p.n[1], p.n[2], p.d]
        assert planes[i].tolist() == [a, b, c, d]  # already unit length
    assert (corners == np.array(k["corners"], np.float32)).all()


def test_k3_product_host_math_matches():
    """fyx_frustum_from_view_projection_matrix is the product's copy of the same extraction."""
    import fyrox_b200 as fb

    k = KATS["K3_frustum_from_identity"]
    f = fb.frustum_from_view_projection_matrix(col_major(k["vp_row_major"]))
    planes, corners = fb.frustum_to_numpy(f)
    assert planes.tolist() == k["planes_abcd"]
    assert (corners == np.array(k["corners"], np.float32)).all()
    # and on random view-projection matrices it equals the oracle bit for bit
    rng = np.random.default_rng(7)
    for _ in range(200):
        eye = rng.uniform(-50, 50, 3).astype(np.float32)
        tgt = eye + rng.normal(size=3).astype(np.float32)
        view = ob.look_at_rh(eye, tgt, (0, 1, 0))
        proj = ob.perspective(float(rng.uniform(0.5, 2.5)), float(rng.uniform(0.3, 2.0)), 0.1, float(rng.uniform(50, 500)))
        vp_o = ob.mat4_mul(proj, view)
        vp_p = fb.mat4_mul(proj, view)
        assert vp_o.tobytes() == vp_p.tobytes()
        fo = ob.frustum_from_vp(vp_o)
        fpz = fb.frustum_from_view_projection_matrix(vp_p)
        po, co = ob.frustum_planes_corners(fo)
        pp, cp = fb.frustum_to_numpy(fpz)
        assert po.tobytes() == pp.tobytes() and co.tobytes() == cp.tobytes()
    assert fb.frustum_from_view_projection_matrix(np.zeros(16, np.float32)) is None
    dflt_p, dflt_c = fb.frustum_to_numpy(fb.frustum_default())
    fo = Frustum()
    L.orc_frustum_default(C.byref(fo))
    po, co = ob.frustum_planes_corners(fo)
    assert po.tobytes() == dflt_p.tobytes() and co.tobytes() == dflt_c.tobytes()


def test_k4_frustum_queries():
    k = KATS["K4_frustum_queries"]
    f = identity_frustum()
    pts = np.array(k["cloud_true"], np.float32)
    assert L.orc_frustum_is_intersects_point_cloud(C.byref(f), fp(pts.reshape(-1)), len(pts))
    pts = np.array(k["cloud_false"], np.float32)
    assert not L.orc_frustum_is_intersects_point_cloud(C.byref(f), fp(pts.reshape(-1)), len(pts))
    t = Aabb.make(k["aabb_true"][:3], k["aabb_true"][3:])
    assert L.orc_frustum_is_intersects_aabb(C.byref(f), C.byref(t))
    fa = Aabb.make(k["aabb_false"][:3], k["aabb_false"][3:])
    assert not L.orc_frustum_is_intersects_aabb(C.byref(f), C.byref(fa))
    assert L.orc_frustum_is_intersects_aabb_offset(C.byref(f), C.byref(t), fp(vec(*k["offset_true"])))
    assert not L.orc_frustum_is_intersects_aabb_offset(C.byref(f), C.byref(t), fp(vec(*k["offset_false"])))
    assert L.orc_frustum_is_contains_point(C.byref(f), fp(vec(*k["contains_true"])))
    assert not L.orc_frustum_is_contains_point(C.byref(f), fp(vec(*k["contains_false"])))


def test_k5_plane():
    k = KATS["K5_plane"]
    p = Plane()
    assert L.orc_plane_from_abcd(1.0, 1.0, 1.0, 0.0, C.byref(p))
    assert [p.n[0], p.n[1], p.n[2]] == np.array(k["from_abcd_1110"]["normal"], np.float32).tolist()
    assert p.d == k["from_abcd_1110"]["d"]
    assert not L.orc_plane_from_abcd(0.0, 0.0, 0.0, 0.0, C.byref(p))
    d = k["dot"]
    q = Plane()
    for i in range(3):
        q.n[i] = d["normal"][i]
    q.d = d["d"]
    assert L.orc_plane_dot(C.byref(q), fp(vec(*d["point"]))) == d["expected"]
    planes = []
    for n in ((0, 0, 1), (0, 1, 0), (1, 0, 0)):
        pl = Plane()
        for i in range(3):
            pl.n[i] = n[i]
        pl.d = 0.0
        planes.append(pl)
    out = np.empty(3, np.float32)
    L.orc_plane_intersection_point(C.byref(planes[0]), C.byref(planes[1]), C.byref(planes[2]), fp(out))
    assert (out == np.array(k["intersection_of_axis_planes"], np.float32)).all()  # -0.0 == 0.0, as assert_eq! on f32


def build_k6_graph():
    """Same construction order as the reference test: c, b, d are built before a (children first)."""
    k = KATS["K6_hierarchy_propagation"]
    g = ob.Graph()
    c = g.add_node()
    g.set_local_matrix(c, ob.translation(*k["local_positions"]["c"]))
    b = g.add_node()
    g.set_local_matrix(b, ob.translation(*k["local_positions"]["b"]))
    g.set_visibility(b, k["b_visibility"])
    g.set_enabled(b, k["b_enabled"])
    g.link_nodes(c, b)
    d = g.add_node()
    g.set_local_matrix(d, ob.translation(*k["local_positions"]["d"]))
    a = g.add_node()
    g.set_local_matrix(a, ob.translation(*k["local_positions"]["a"]))
    g.link_nodes(b, a)
    g.link_nodes(d, a)
    return g, dict(a=a, b=b, c=c, d=d)


def check_k6(g, h, exp):
    for n, pos in exp["global_positions"].items():
        assert g.global_position(h[n]).tolist() == pos, n
    for n, v in exp["global_visibility"].items():
        assert g.global_visibility(h[n]) == v, n
    for n, v in exp["global_enabled"].items():
        assert g.is_globally_enabled(h[n]) == v, n


def test_k6_hierarchy_changes_propagation():
    k = KATS["K6_hierarchy_propagation"]
    g, h = build_k6_graph()
    assert h["c"] == 1 and h["a"] == 4  # K10 numbering: first added node is index 1
    g.update()
    check_k6(g, h, k["first"])
    g.set_local_matrix(h["b"], ob.translation(0.0, 2.0, 0.0))
    g.set_enabled(h["a"], False)
    g.set_visibility(h["b"], True)
    g.update()
    check_k6(g, h, k["second"])
    # a full recompute gives the same state
    g.update_hierarchical_data()
    check_k6(g, h, k["second"])


def test_k7_global_scale():
    k = KATS["K7_global_scale"]
    g = ob.Graph()
    c = g.add_node()
    b = g.add_node()
    g.link_nodes(c, b)
    a = g.add_node()
    g.link_nodes(b, a)
    h = dict(a=a, b=b, c=c)
    ls = np.ones((g.capacity, 3), np.float32)
    for n, s in k["local_scales"].items():
        ls[h[n]] = s
    for n, e in k["expected"].items():
        out = np.empty(3, np.float32)
        L.orc_graph_global_scale(g.h, h[n], fp(ls.reshape(-1)), fp(out))
        assert out.tolist() == e


def test_k8_matrix_layout():
    """Matrix4Ext accessors are plain linear (column-major) indices; orc matrices use the same layout."""
    k = KATS["K8_matrix4_ext"]
    m = np.empty(16, np.float32)
    L.orc_mat4_identity(fp(m))
    for name in ("side", "up", "look", "position"):
        assert m[k["linear_indices"][name]].tolist() == k[name]
    t = ob.translation(7, 8, 9)
    assert t[k["linear_indices"]["position"]].tolist() == [7, 8, 9]


def test_k9_sorting_index():
    k = KATS["K9_sorting_index"]
    view = np.empty(16, np.float32)
    L.orc_mat4_identity(fp(view))
    for case in k["cases"]:
        got = L.orc_calculate_sorting_index(fp(view), fp(vec(0, 0, case["z"])))
        assert got == k["range_center"] + case["delta"]


def test_k10_handle_numbering():
    k = KATS["K10_handle_numbering"]
    g = ob.Graph()
    assert L.orc_graph_root(g.h) == k["root"][0]
    assert g.add_node() == k["first_added"][0]


# ---- unpinned parts: self-consistency of the restatement -------------------------------------------
def test_mat4_mul_matches_numpy_within_rounding_and_order_is_left_to_right():
    rng = np.random.default_rng(1)
    a = rng.normal(size=16).astype(np.float32)
    b = rng.normal(size=16).astype(np.float32)
    c = ob.mat4_mul(a, b)
    A, B = a.reshape(4, 4).T, b.reshape(4, 4).T
    ref = (A.astype(np.float64) @ B.astype(np.float64)).T.reshape(16)
    assert np.allclose(c, ref, rtol=1e-5, atol=1e-5)
    # exact left-to-right accumulation, one rounding per op
    f = np.float32
    for j in range(4):
        for i in range(4):
            y = f(A[i, 0] * B[0, j])
            y = f(y + f(A[i, 1] * B[1, j]))
            y = f(y + f(A[i, 2] * B[2, j]))
            y = f(y + f(A[i, 3] * B[3, j]))
            assert c[j * 4 + i] == y


def test_calculate_local_transform_reduces_to_trs():
    t = ob.Transform()
    L.orc_transform_identity(C.byref(t))
    t.local_position[:] = (1.0, 2.0, 3.0)
    t.local_scale[:] = (2.0, 3.0, 4.0)
    q = np.array([0.1, -0.2, 0.3, 0.9], np.float32)
    q /= np.linalg.norm(q)
    t.local_rotation[:] = q.tolist()
    m = np.empty(16, np.float32)
    L.orc_transform_calculate_local(C.byref(t), fp(m))
    r = np.empty(9, np.float32)
    L.orc_quat_to_rotation_matrix(fp(q), fp(r))
    R = r.reshape(3, 3).T
    M = m.reshape(4, 4).T
    assert np.allclose(M[:3, :3], R * np.array([2.0, 3.0, 4.0], np.float32)[None, :], atol=1e-6)
    assert M[:3, 3].tolist() == [1.0, 2.0, 3.0]
    assert M[3].tolist() == [0.0, 0.0, 0.0, 1.0]
    # the Python host mirror computes the same matrix bit for bit
    from fyrox_b200.scene import TransformBuilder

    tm = TransformBuilder().with_local_position((1, 2, 3)).with_local_scale((2, 3, 4)).with_local_rotation(q).build().matrix()
    assert tm.tobytes() == m.tobytes()


def test_skinned_mesh_world_aabb_reads_current_bone_positions():
    """The DFS-order quirk: bones visited before the mesh contribute their NEW position."""
    g = ob.Graph()
    bone = g.add_node()
    mesh = g.add_node(ob.KIND_MESH)
    g.set_local_aabb(mesh, [-1, -1, -1, 1, 1, 1])
    g.add_surface(mesh, [bone])
    g.set_local_matrix(bone, ob.translation(10, 0, 0))
    g.update()
    assert g.world_bounding_box(mesh).tolist() == [-1, -1, -1, 10, 1, 1]
    # the bone moves, the mesh does not: the cached box stays (Mesh world AABB is only refreshed when the
    # mesh's own global transform changes, mesh/mod.rs:667-689)
    g.set_local_matrix(bone, ob.translation(20, 0, 0))
    g.update()
    assert g.world_bounding_box(mesh).tolist() == [-1, -1, -1, 10, 1, 1]
    g.set_local_matrix(mesh, ob.translation(0, 0, 0))
    g.update()
    assert g.world_bounding_box(mesh).tolist() == [-1, -1, -1, 20, 1, 1]


# ---- N2 (animation sampling): the reference's own curve / wrapf tests ---------------------------------
def _keys(rows, kind=ob.KEY_LINEAR):
    k = np.zeros(len(rows), ob.KEY_DTYPE)
    for i, r in enumerate(rows):
        k[i]["location"], k[i]["value"] = r[0], r[1]
        k[i]["kind"] = r[2] if len(r) > 2 else kind
    return k


def test_k11_wrapf():
    for n, lo, hi, want in KATS["K11_wrapf"]["cases"]:
        assert L.orc_wrapf(n, lo, hi) == want


def test_k12_curve_value_at():
    for step in KATS["K12_curve_value_at"]["steps"]:
        keys = _keys(step["keys"])
        for loc, want in step["queries"]:
            got, _ = ob.curve_value_at(keys, loc, 0)
            assert got == want, (step["keys"], loc)


def test_k13_curve_key_interpolate():
    k = KATS["K13_curve_key_interpolate"]
    keys = {}
    for name, (loc, val, kind) in k["keys"].items():
        keys[name] = ob.CurveKey(loc, val, kind, 0.0, 0.0)
    for a, b, t, want in k["cases"]:
        assert L.orc_key_interpolate(C.byref(keys[a]), C.byref(keys[b]), t) == want, (a, b, t)


def test_curve_hint_is_part_of_the_result():
    """Not a reference test, a property of curve.rs:275-301 the restatement must keep: on a key's location the span
    found through the hint (t = 0 on the right span) and the one found by binary search (t = 1 on the left span) are
    different expressions, so the remembered hint is state."""
    keys = _keys([(0.0, 0.1), (1.0, 0.7), (2.0, 0.3)])
    v_search, h = ob.curve_value_at(keys, 1.0, 0)      # hint 0: binary search -> span [0,1], t = 1: 0.1 + (0.7-0.1)*1
    assert h == 1
    v_hint, h2 = ob.curve_value_at(keys, 1.0, 2)       # hint 2: span [1,2), t = 0: exactly 0.7
    assert h2 == 2 and v_hint == float(np.float32(0.7))
    assert v_search == float(np.float32(0.1) + (np.float32(0.7) - np.float32(0.1)) * np.float32(1.0))


def test_animation_tick_applies_pose_then_advances_time():
    """Animation::tick order (lib.rs:471-496): the pose is sampled at the CURRENT time position, then time advances and
    wraps (wrapf).  One node, one linear position track (0,0,0)@0 -> (2,4,8)@2, dt 0.5 from t = 0.5."""
    keys = _keys([(0, 0), (2, 2), (0, 0), (2, 4), (0, 0), (2, 8)])
    t = np.zeros(1, ob.TRACK_DTYPE)
    t[0]["target_node"], t[0]["binding"], t[0]["value_kind"], t[0]["enabled"], t[0]["n_curves"] = 1, ob.BIND_POSITION, ob.TV_VECTOR3, 1, 3
    t[0]["first_key"][:3] = [0, 2, 4]
    t[0]["n_keys"][:3] = [2, 2, 2]
    a = ob.Animation(t, keys, speed=1.0, looped=True, time_slice=(0.0, 2.0), time_position=0.5)
    parent = np.array([0xFFFFFFFF, 0], np.uint32)
    og = ob.Graph.build(parent, None, None, np.tile(np.eye(4, dtype=np.float32).reshape(16), (2, 1)), None)
    og.update_hierarchical_data()
    tr = (ob.Transform * 2)()
    for i in range(2):
        L.orc_transform_identity(C.byref(tr[i]))
    want = [((0.5, 1.0, 2.0), 1.0), ((1.0, 2.0, 4.0), 1.5), ((1.5, 3.0, 6.0), 2.0), ((2.0, 4.0, 8.0), 0.5), ((0.5, 1.0, 2.0), 1.0)]
    for pos, time_after in want:
        ob.update_animations([a], 0.5, og, tr)
        og.update()
        assert tuple(tr[1].local_position) == pos
        assert tuple(og.global_transform(1)[12:15]) == pos
        assert a.time_position == time_after
    # a clamped (not looped) animation stops at the end of its slice
    b = ob.Animation(t, keys, speed=1.0, looped=False, time_slice=(0.0, 2.0), time_position=1.75)
    ob.update_animations([b], 0.5, og, tr)
    assert b.time_position == 2.0
    # TrackValue::blend_with: lerp for vectors, shortest-way nlerp for rotations (value.rs:201-227,449-454)
    va = np.array([1, 2, 3, 4], np.float32)
    L.orc_track_value_blend(0, fp(va), fp(np.array([3, 2, 1, 0], np.float32)), 0.25)
    assert va.tolist() == [1.5, 2.0, 2.5, 3.0]
    qa = np.array([0, 0, 0, 1], np.float32)
    L.orc_track_value_blend(1, fp(qa), fp(np.array([0, 0, 0, -1], np.float32)), 0.5)
    assert qa.tolist() == [0.0, 0.0, 0.0, -1.0]  # a is negated first (dot < 0), so the blend does not pass through zero


def test_mt_oracle_equals_the_single_threaded_one():
    """The multi-core CPU baseline (fyrox_oracle_mt.c, flat arrays + OpenMP) computes exactly what the reference-shaped
    restatement does: global transforms, flags, mesh boxes (incl. the skinned-mesh bone fold), visible sets for the
    camera and the six cube faces, palettes + skinned streams."""
    from fyrox_b200.scenegen import Scene
    from helpers import cube_frusta, camera_frustum

    sc = Scene(6000, 8, verts_per_unit=300)
    aabb = sc.local_aabb.copy()
    og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, aabb)
    mt = ob.MtGraph(sc.parent, sc.flags, sc.render_mask, sc.local_m16, aabb, threads=4)
    for u in range(sc.n_units):
        mesh, bones, ib = sc.unit_mesh_node(u), sc.unit_bone_nodes(u), sc.unit_inv_bind(u)
        verts, _ = sc.unit_vertices(u)
        for k, b in enumerate(bones):
            og.set_inv_bind(int(b), ib[k])
            mt.set_inv_bind(int(b), ib[k])
        og.add_surface(mesh, bones, verts)
        og.recalc_local_aabb(mesh)
        assert mt.add_surface(mesh, bones, verts) == u
    for frame in range(2):
        idx, m = sc.animate(frame)
        for i in range(idx.size):
            og.set_local_matrix(int(idx[i]), m[i])
        og.L.orc_graph_drop_messages(og.h)
        mt.set_local_matrices(m, idx)
        og.update_hierarchical_data()
        mt.update()
        G = og.global_transforms()
        for i in range(sc.capacity):
            g, wa, fl = mt.get(i)
            assert np.array_equal(g.view(np.uint32), G[i].view(np.uint32)), i
            assert (fl & 1) == og.global_visibility(i) and ((fl >> 1) & 1) == og.is_globally_enabled(i)
            if sc.flags[i] & (1 << 5):
                assert np.array_equal(wa.view(np.uint32), og.world_bounding_box(i).view(np.uint32)), i
        fos, _ = cube_frusta()
        fo, _ = camera_frustum()
        for k, f in enumerate(fos + [fo]):
            shadow = k < 6 and k % 2 == 1
            want = np.sort(og.from_graph(f, 0xFFFF00FF, shadow))
            got = mt.cull(f, 0xFFFF00FF, shadow)
            assert np.array_equal(got, want)
        for u in range(sc.n_units):
            p0, n0 = og.skin(sc.unit_mesh_node(u), 0, sc.verts_per_unit)
            p1, n1 = mt.skin(u, sc.verts_per_unit)
            assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32)) and np.array_equal(n0.view(np.uint32), n1.view(np.uint32))
    mt.skin_all()


def test_blend_group_clones_the_first_source_and_lerps_the_rest():
    """BlendAnimations::eval_pose / AnimationPose::blend_with (machine/node/blend.rs:136-166, pose.rs:41-101): the first
    source with values for a node is cloned (its weight is not used), later sources are lerped in per binding; a binding
    the output does not have is dropped."""
    def const_track(node, binding, vals):
        keys = _keys([(0.0, v) for v in vals], kind=ob.KEY_CONSTANT)
        t = np.zeros(1, ob.TRACK_DTYPE)
        t[0]["target_node"], t[0]["binding"], t[0]["value_kind"], t[0]["enabled"], t[0]["n_curves"] = node, binding, ob.TV_VECTOR3, 1, 3
        t[0]["first_key"][:3] = [0, 1, 2]
        t[0]["n_keys"][:3] = [1, 1, 1]
        return t, keys

    ta, ka = const_track(1, ob.BIND_POSITION, (0.0, 0.0, 0.0))
    tb, kb = const_track(1, ob.BIND_POSITION, (2.0, 4.0, 6.0))
    ts, ks = const_track(1, ob.BIND_SCALE, (3.0, 3.0, 3.0))
    tb2 = np.concatenate([tb, ts])
    tb2[1]["first_key"][:3] = [3, 4, 5]
    kb2 = np.concatenate([kb, ks])
    a = ob.Animation(ta, ka, time_slice=(0.0, 1.0))
    b = ob.Animation(tb2, kb2, time_slice=(0.0, 1.0))
    parent = np.array([0xFFFFFFFF, 0], np.uint32)
    og = ob.Graph.build(parent, None, None, np.tile(np.eye(4, dtype=np.float32).reshape(16), (2, 1)), None)
    tr = (ob.Transform * 2)()
    for i in range(2):
        L.orc_transform_identity(C.byref(tr[i]))
    ob.blend_group_update([a, b], [0.9, 0.25], 0.1, og, tr)
    assert tuple(tr[1].local_position) == (0.5, 1.0, 1.5)   # a*(1-0.25) + b*0.25; a's own weight 0.9 plays no role
    assert tuple(tr[1].local_scale) == (1.0, 1.0, 1.0)      # b's scale has no counterpart in the output: dropped
    ob.blend_group_update([b, a], [0.9, 0.25], 0.1, og, tr)
    assert tuple(tr[1].local_position) == (1.5, 3.0, 4.5) and tuple(tr[1].local_scale) == (3.0, 3.0, 3.0)
    # a disabled source is not ticked but its last pose still blends in
    L.orc_animation_set_enabled(b.h, 0)
    t_before = b.time_position
    ob.blend_group_update([a, b], [1.0, 0.5], 0.1, og, tr)
    assert b.time_position == t_before and tuple(tr[1].local_position) == (1.0, 2.0, 3.0)


def test_lod_filter_prunes_whole_subtrees_and_later_groups_win():
    """from_graph's lod_filter (renderer/bundle.rs:898-916, 988-1004): an object outside its level's normalised-distance
    range hides its whole sub-tree; an object listed twice takes the verdict written last (pool order of the owners)."""
    from helpers import camera_frustum

    NONE_ = 0xFFFFFFFF
    RENDER = ob.lib() and (1 << 5)
    DEF = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4)
    #          0 root   1 owner  2 lvl0   3 lvl1   4 mesh<2  5 mesh<3  6 owner2
    parent = np.array([NONE_, 0, 1, 1, 2, 3, 0], np.uint32)
    flags = np.array([DEF, DEF, DEF, DEF, DEF | RENDER, DEF | RENDER, DEF], np.uint32)
    local = np.tile(np.eye(4, dtype=np.float32).reshape(16), (7, 1))
    local[1, 12:15] = (0, 0, -10)  # the whole group sits 10 units in front of the camera
    aabb = np.tile(np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], np.float32), (7, 1))
    og = ob.Graph.build(parent, flags, None, local, aabb)
    og.update_hierarchical_data()
    og.set_lod_group(1, [(0.0, 0.3, [2]), (0.3, 1.0, [3])])
    fo, _ = camera_frustum(zfar=100.0)
    eye = (0.0, 0.0, 0.0)
    assert og.lod_filter(eye, 0.0, 100.0).tolist() == [True, True, True, False, True, True, True]
    assert sorted(og.from_graph(fo).tolist()) == [4, 5]                       # no LOD: both meshes
    assert og.from_graph_lod(fo, eye, 0.0, 100.0).tolist() == [4]             # level 1 is out of range: its mesh goes with it
    assert og.from_graph_lod(fo, eye, 0.0, 20.0).tolist() == [5]              # normalised 0.5: the other level
    # a second owner (higher index, so visited later) lists node 2 with a range it falls outside of: the later verdict wins
    og.set_lod_group(6, [(0.5, 1.0, [2])])
    assert og.from_graph_lod(fo, eye, 0.0, 100.0).tolist() == []


def test_k15_quat_from_euler_equals_nalgebras_closed_form_bit_for_bit():
    """fyrox-math/src/lib.rs:1460-1477 asserts quat_from_euler((pi, pi, pi), XYZ) == UnitQuaternion::from_euler_angles(pi, pi, pi)
    (UnitQuaternion's PartialEq: all coordinates equal, or all negated).  The oracle's product qz * qy * qx (quat_mul's
    operation order) against nalgebra's closed form restated here in float32, one rounding per operation: the coordinates are
    1 and +-4.37e-8 and their last ulps differ with the order of the sums — they must agree in every bit."""
    k = KATS["K15_quat_from_euler"]
    assert k["euler"] == ["pi", "pi", "pi"] and k["order"] == "XYZ"
    f32 = np.float32
    e = np.full(3, np.pi, f32)
    q = np.zeros(4, f32)
    L.orc_quat_from_euler_xyz(fp(e), fp(q))
    libm = C.CDLL("libm.so.6")
    libm.sinf.restype = libm.cosf.restype = C.c_float
    libm.sinf.argtypes = libm.cosf.argtypes = [C.c_float]
    h = f32(np.pi) * f32(0.5)
    s, c = f32(libm.sinf(float(h))), f32(libm.cosf(float(h)))
    assert s == 1.0 and c != 0.0  # the case is not degenerate: cos(pi_f32 / 2) is -4.37e-8, not 0
    sr = sp = sy = s
    cr = cp = cy = c
    w = cr * cp * cy + sr * sp * sy
    i = sr * cp * cy - cr * sp * sy
    j = cr * sp * cy + sr * cp * sy
    kk = cr * cp * sy - sr * sp * cy
    want = np.array([i, j, kk, w], f32)
    same = (want.view(np.uint32) == q.view(np.uint32)).all()
    negated = (want.view(np.uint32) == (-q).view(np.uint32)).all()
    assert same or negated, (q, want)
    # and the two small coordinates really are different floats (the check can tell summation orders apart)
    assert abs(q[0]) != abs(q[1])


def test_k16_vertex_buffer_attributes_are_read_where_the_reference_puts_them():
    """fyrox-impl/src/scene/mesh/buffer.rs:1687-1828: the reference's interleaved test vertex (76 bytes) and the values its attribute
    views return.  The oracle's vertex reader (the layout the C ABI's fyx_vertex_layout mirrors) must pick position, normal, bone
    weights and the u8 bone indices from the same offsets: skinned with a palette whose bone b translates by (b, 0, 0), every vertex
    comes out at position + 0.25 * (1 + 2 + 3 + 4) on x — exact in f32 for these values — and the normal unchanged."""
    k = KATS["K16_vertex_buffer_attributes"]
    off = k["offsets"]
    n = len(k["vertices"])
    buf = np.zeros(n * k["stride"], np.uint8)
    for i, v in enumerate(k["vertices"]):
        rec = buf[i * k["stride"]:(i + 1) * k["stride"]]
        for name in ("position", "tex_coord", "second_tex_coord", "normal", "tangent", "bone_weights"):
            a = np.asarray(v[name], np.float32)
            rec[off[name]:off[name] + 4 * a.size] = a.view(np.uint8)
        rec[off["bone_indices"]:off["bone_indices"] + 4] = np.asarray(v["bone_indices"], np.uint8)
    sizes = {"position": 12, "tex_coord": 8, "second_tex_coord": 8, "normal": 12, "tangent": 16, "bone_weights": 16, "bone_indices": 4}
    run = 0
    for name in ("position", "tex_coord", "second_tex_coord", "normal", "tangent", "bone_weights", "bone_indices"):
        assert off[name] == run  # running sum of the declared attribute sizes, as VertexBuffer::new lays them out
        run += sizes[name]
    assert run == k["stride"]
    layout = ob.VertexLayout(k["stride"], off["position"], off["normal"], off["bone_weights"], off["bone_indices"])
    pal = np.zeros((6, 16), np.float32)
    for b in range(6):
        pal[b] = ob.translation(float(b), 0.0, 0.0)
    pos = np.zeros((n, 3), np.float32)
    nrm = np.zeros((n, 3), np.float32)
    L.orc_skin_vertices(fp(pal), n, buf.ctypes.data_as(C.c_void_p), C.byref(layout), fp(pos), fp(nrm))
    for i, v in enumerate(k["vertices"]):
        want = np.asarray(v["position"], np.float32) + np.array([2.5, 0.0, 0.0], np.float32)
        assert np.array_equal(pos[i], want), (i, pos[i], want)
        assert np.array_equal(nrm[i], np.asarray(v["normal"], np.float32)), (i, nrm[i])
