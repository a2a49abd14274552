"""The sampler behind the full-size GPU tests and bench.py's parity block (tests/sampled_parity.py) rebuilds reference
values node by node from chain products.  Here it is pinned, on the CPU, against the oracle walking the WHOLE scene
(pointer tree, recursive DFS, from_graph): every node's matrix, box and per-frustum visibility, every unit's palette
and skinned stream must agree bit for bit — including animated bones given as TRS records, skinned-mesh bone folds
and a 2-way sharded scene."""
import numpy as np

import oracle_binding as ob
import sampled_parity as sp
from fyrox_b200.scenegen import Scene
from helpers import cube_frusta


def full_oracle(sc, frame_local):
    aabb = sc.local_aabb.copy()
    og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, aabb)
    mesh_aabb, bones_of = {}, {}
    for u in range(sc.n_units):
        mesh, bones, ib = sc.unit_mesh_node(u), sc.unit_bone_nodes(u), sc.unit_inv_bind(u)
        for k, b in enumerate(bones):
            og.set_inv_bind(int(b), ib[k])
        verts, bb = sc.unit_vertices(u)
        og.add_surface(mesh, bones, verts)
        og.recalc_local_aabb(mesh)
        mesh_aabb[int(mesh)] = bb
        bones_of[int(mesh)] = bones
    idx, loc = frame_local
    for k, i in enumerate(idx):
        og.set_local_matrix(int(i), loc(k) if callable(loc) else loc[k])
    og.L.orc_graph_drop_messages(og.h)
    og.update_hierarchical_data()
    return og, mesh_aabb, bones_of


def run(sc, use_trs):
    if use_trs:
        idx, trs = sc.animate_trs(1)
        loc = sp.trs_bone_local(trs)
    else:
        idx, loc = sc.animate(1)
    og, mesh_aabb, bones_of = full_oracle(sc, (idx, loc))
    truth = sp.SampledTruth(sc, bone_idx=idx, bone_local=loc, mesh_aabb=mesh_aabb)
    fos, _ = cube_frusta()
    want = [set(og.from_graph(fo).tolist()) for fo in fos]
    all_nodes = np.arange(sc.capacity, dtype=np.uint32)
    Go, Ao = og.global_transforms(all_nodes), og.world_bounding_boxes(all_nodes)
    for i in range(sc.capacity):
        g, a, vis = truth.node(i, fos, 0xFFFFFFFF, bones_of.get(i))
        assert sp.bits_equal(g, Go[i]).all(), i
        assert sp.bits_equal(a, Ao[i]).all(), i
        for f in range(len(fos)):
            assert vis[f] == (i in want[f]), (i, f)
    for u in range(sc.n_units):
        verts, _ = sc.unit_vertices(u)
        pal, pos, nrm = truth.unit(u, verts)
        assert sp.bits_equal(pal, og.bone_matrices(sc.unit_mesh_node(u), 0, sc.bones_per_unit)).all()
        po, no = og.skin(sc.unit_mesh_node(u), 0, sc.verts_per_unit)
        assert pos.tobytes() == po.tobytes() and nrm.tobytes() == no.tobytes()
    return want


def test_sampler_equals_whole_scene_oracle_matrices_and_trs():
    run(Scene(3000, n_units=6, verts_per_unit=96), use_trs=False)
    run(Scene(3000, n_units=6, verts_per_unit=96), use_trs=True)


def test_sampler_on_shards_partitions_the_unsharded_visible_sets():
    whole = run(Scene(4000, n_units=8, verts_per_unit=32), use_trs=True)
    parts = [set() for _ in whole]
    for r in range(2):
        sc = Scene(4000, n_units=8, verts_per_unit=32, rank=r, nranks=2)
        idx, trs = sc.animate_trs(1)
        mesh_aabb = {int(sc.unit_mesh_node(u)): sc.unit_vertices(u)[1] for u in range(sc.n_units)}
        bones_of = {int(sc.unit_mesh_node(u)): sc.unit_bone_nodes(u) for u in range(sc.n_units)}
        truth = sp.SampledTruth(sc, bone_idx=idx, bone_local=sp.trs_bone_local(trs), mesh_aabb=mesh_aabb)
        fos, _ = cube_frusta()
        for i in range(1 if r else 0, sc.capacity):  # the replicated root belongs to rank 0's list
            _, _, vis = truth.node(i, fos, 0xFFFFFFFF, bones_of.get(i))
            for f, v in enumerate(vis):
                if v:
                    assert int(sc.global_index[i]) not in parts[f]
                    parts[f].add(int(sc.global_index[i]))
    for f in range(len(whole)):
        assert parts[f] == whole[f]


def test_list_checksum_is_additive_over_disjoint_parts():
    rng = np.random.default_rng(1)
    a = rng.permutation(100000).astype(np.uint32)
    c = [sp.list_checksum(p) for p in (a[:30000], a[30000:70000], a[70000:])]
    n, s, x = sp.list_checksum(a)
    assert n == sum(t[0] for t in c) and s == sum(t[1] for t in c) % (1 << 64) and x == c[0][2] ^ c[1][2] ^ c[2][2]
    assert sp.list_checksum(np.empty(0, np.uint32)) == (0, 0, 0)
