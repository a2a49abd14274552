"""Shared helpers of the GPU parity tests: build the same scene in the oracle (CPU restatement) and in
a fyx context (CUDA path through the C ABI), and compare bit for bit."""
from __future__ import annotations

import numpy as np

import fyrox_b200 as fb
import oracle_binding as ob
from fyrox_b200.scenegen import Scene

NONE = 0xFFFFFFFF
UNIT_BOX = np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], np.float32)


def camera_frustum(eye=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), aspect=16 / 9, fovy=np.deg2rad(60.0), znear=0.1, zfar=150.0):
    """F=1 of SURVEY §8d: look_at_rh toward -Z, new_perspective(16/9, 60deg, 0.1, 150).  Returns (oracle Frustum, fyx_frustum)."""
    view = ob.look_at_rh(eye, target, up)
    proj = ob.perspective(float(aspect), float(fovy), float(znear), float(zfar))
    fo = ob.frustum_from_vp(ob.mat4_mul(proj, view))
    ff = fb.frustum_from_view_projection_matrix(fb.mat4_mul(proj, view))
    return fo, ff


# renderer/utils.rs:49-75 — cube-map face look/up vectors of the point-light shadow pass
CUBE_FACES = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, -1)), ((0, 0, 1), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]


def cube_frusta(origin=(0, 0, 0), radius=120.0):
    """F=6 of SURVEY §8d: renderer/shadow/point.rs:162-178 (new_perspective(1, pi/2, 0.01, R), look_at_rh per face)."""
    out = []
    for look, up in CUBE_FACES:
        tgt = tuple(origin[i] + look[i] for i in range(3))
        out.append(camera_frustum(origin, tgt, up, 1.0, np.pi / 2, 0.01, radius))
    return [o for o, _ in out], [f for _, f in out]


def scene_pair(sc: Scene, ctx: fb.Context, with_vertices=True):
    """Load a generated scene into the oracle and into a fyx context.  Returns (oracle graph, surface ids)."""
    aabb = sc.local_aabb.copy()
    og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, aabb)
    surfaces = []
    vert_store = []
    for u in range(sc.n_units):
        mesh = sc.unit_mesh_node(u)
        bones = sc.unit_bone_nodes(u)
        ib = sc.unit_inv_bind(u)
        for k, b in enumerate(bones):
            og.set_inv_bind(int(b), ib[k])
        if with_vertices:
            verts, bb = sc.unit_vertices(u)
            vert_store.append(verts)
            og.add_surface(mesh, bones, verts)
            og.recalc_local_aabb(mesh)  # Mesh::local_bounding_box from the vertex positions
            aabb[mesh] = bb
        else:
            og.add_surface(mesh, bones)
        surfaces.append((mesh, bones, ib, vert_store[-1] if with_vertices else None))
    og.L.orc_graph_drop_messages(og.h)
    ctx.set_topology(sc.parent, sc.flags, sc.render_mask, aabb, root=0)
    ctx.set_local_matrices(sc.local_m16)
    sids = []
    for mesh, bones, ib, verts in surfaces:
        sids.append(ctx.add_skinned_surface(mesh, bones, ib, verts))
    return og, sids


def bits_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Element-wise bit equality of two f32 arrays, except that any NaN equals any NaN (neither Rust nor
    CUDA defines the sign/payload of a generated NaN: x86 makes 0xFFC00000, the GPU 0x7FFFFFFF)."""
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def assert_same_hierarchy(og: ob.Graph, ctx: fb.Context, idx=None):
    n = og.capacity
    idx = np.arange(n, dtype=np.uint32) if idx is None else np.asarray(idx, np.uint32)
    G = ctx.get_global_matrices(idx)
    A = ctx.get_world_aabbs(idx)
    F = ctx.get_global_flags(idx)
    Go = og.global_transforms(idx)
    Ao = og.world_bounding_boxes(idx)
    assert bits_equal(G, Go).all(), f"global transforms differ at {idx[np.nonzero((~bits_equal(G, Go)).any(axis=1))[0][:10]]}"
    assert bits_equal(A, Ao).all(), f"world AABBs differ at {idx[np.nonzero((~bits_equal(A, Ao)).any(axis=1))[0][:10]]}"
    gv = np.array([og.global_visibility(int(i)) for i in idx])
    ge = np.array([og.is_globally_enabled(int(i)) for i in idx])
    assert (((F & fb.NODE_GLOBAL_VISIBILITY) != 0) == gv).all()
    assert (((F & fb.NODE_GLOBAL_ENABLED) != 0) == ge).all()


def assert_same_visible(og: ob.Graph, ctx: fb.Context, oracle_frusta, cam_mask=None, pass_flags=None):
    for f, fo in enumerate(oracle_frusta):
        mask = 0xFFFFFFFF if cam_mask is None else int(cam_mask[f])
        shadow = bool(pass_flags[f] & fb.PASS_SHADOW) if pass_flags is not None else False
        want = np.sort(og.from_graph(fo, mask, shadow))
        got = np.sort(ctx.get_visible(f))
        assert got.size == np.unique(got).size, "duplicate entries in a visible list"
        assert np.array_equal(got, want), f"frustum {f}: {got.size} visible on the GPU, {want.size} in the oracle"
    return want.size


def random_graph(rng, n, p_dead=0.05, p_orphan=0.01, p_mesh=0.6, max_depth_bias=0.0):
    """Random forest with index-shuffled nodes: parents may have larger indices than their children, some
    pool records are free (dead), some alive nodes are orphans (unreachable from the root)."""
    order = rng.permutation(n - 1) + 1  # creation order of nodes 1..n-1; node 0 is the root
    parent = np.full(n, NONE, np.uint32)
    placed = [0]
    flags = np.zeros(n, np.uint32)
    flags[0] = fb.NODE_DEFAULT
    dead = rng.random(n) < p_dead
    dead[0] = False
    for i in order:
        if dead[i]:
            continue
        f = fb.NODE_ALIVE
        f |= fb.NODE_VISIBILITY if rng.random() < 0.93 else 0
        f |= fb.NODE_ENABLED if rng.random() < 0.96 else 0
        f |= fb.NODE_FRUSTUM_CULLING if rng.random() < 0.9 else 0
        f |= fb.NODE_CAST_SHADOWS if rng.random() < 0.7 else 0
        f |= fb.NODE_RENDERABLE if rng.random() < p_mesh else 0
        flags[i] = f
        if rng.random() < p_orphan:
            continue  # stays parentless: an orphan sub-tree root
        if max_depth_bias > 0 and rng.random() < max_depth_bias:
            parent[i] = placed[-1]  # chain ⇒ deep hierarchies
        else:
            parent[i] = placed[rng.integers(len(placed))]
        placed.append(int(i))
    mask = np.where(rng.random(n) < 0.9, 0xFFFFFFFF, (1 << rng.integers(0, 32, n)).astype(np.uint64)).astype(np.uint32)
    local = np.zeros((n, 16), np.float32)
    for i in range(n):
        q = rng.normal(size=4).astype(np.float32)
        q /= np.linalg.norm(q)
        t = ob.Transform()
        ob.lib().orc_transform_identity(t)
        t.local_position[:] = rng.uniform(-20, 20, 3).astype(np.float32).tolist()
        t.local_rotation[:] = q.tolist()
        t.local_scale[:] = rng.uniform(0.5, 1.5, 3).astype(np.float32).tolist()
        ob.lib().orc_transform_calculate_local(t, ob.fp(local[i]))
    local[0] = np.eye(4, dtype=np.float32).reshape(16)
    aabb = np.tile(UNIT_BOX, (n, 1))
    is_mesh = (flags & fb.NODE_RENDERABLE) != 0
    h = rng.uniform(0.1, 3.0, (n, 3)).astype(np.float32)
    aabb[is_mesh, :3] = -h[is_mesh]
    aabb[is_mesh, 3:] = h[is_mesh]
    return parent, flags, mask, local, aabb


def preorder_rank(parent):
    """Pre-order DFS rank from node 0 with children in index order (what ob.Graph.build produces)."""
    n = len(parent)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        if parent[i] != NONE:
            kids[int(parent[i])].append(i)
    rank = np.full(n, NONE, np.uint32)
    stack, r = [0], 0
    while stack:
        x = stack.pop()
        rank[x] = r
        r += 1
        stack.extend(reversed(kids[x]))
    return rank
