"""Synthetic scene generator: determinism, layout invariants and shard consistency (CPU only)."""
import numpy as np

import fyrox_b200 as fb
from fyrox_b200.scenegen import Scene

NONE = 0xFFFFFFFF


def test_layout_and_determinism():
    a = Scene(20000, n_units=20, verts_per_unit=100)
    b = Scene(20000, n_units=20, verts_per_unit=100)
    assert a.capacity == 20000 == b.capacity
    assert a.n_units == 20
    for k in ("parent", "flags", "render_mask", "global_index"):
        assert (getattr(a, k) == getattr(b, k)).all()
    assert a.local_m16.tobytes() == b.local_m16.tobytes()
    assert a.parent[0] == NONE and (a.parent[1:] != NONE).all()
    # affine local matrices with an exact (+0,+0,+0,1) bottom row
    bottom = a.local_m16[:, [3, 7, 11, 15]]
    assert (bottom.view(np.uint32) == np.array([0, 0, 0, 0x3F800000], np.uint32)).all()
    # every node alive; leaves + skinned meshes renderable
    assert (a.flags & fb.NODE_ALIVE).all()
    assert a.n_renderable == int(((a.flags & fb.NODE_RENDERABLE) != 0).sum())
    # bones precede their mesh node and hang under the same group
    for u in range(a.n_units):
        bones = a.unit_bone_nodes(u)
        mesh = a.unit_mesh_node(u)
        assert (bones < mesh).all() and mesh == bones[-1] + 1
        assert a.parent[mesh] == a.parent[bones[0]]
        assert a.flags[mesh] & fb.NODE_RENDERABLE
        ib = a.unit_inv_bind(u)
        assert (ib[:, [3, 7, 11, 15]].view(np.uint32) == np.array([0, 0, 0, 0x3F800000], np.uint32)).all()
    v1, bb1 = a.unit_vertices(3)
    v2, bb2 = b.unit_vertices(3)
    assert v1.tobytes() == v2.tobytes() and bb1.tolist() == bb2.tolist()
    rec = v1.reshape(-1, 68)
    w = rec[:, 48:64].copy().view(np.float32)
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-5)
    assert (rec[:, 64:68] < 64).all()
    i1, m1 = a.animate(5)
    i2, m2 = b.animate(5)
    assert (i1 == i2).all() and m1.tobytes() == m2.tobytes()
    assert not np.array_equal(a.animate(6)[1], m1)


def test_shards_partition_the_scene():
    whole = Scene(30000, n_units=40, verts_per_unit=10)
    R = 4
    shards = [Scene(30000, n_units=40, verts_per_unit=10, rank=r, nranks=R) for r in range(R)]
    seen = np.zeros(30000, dtype=np.int32)
    for s in shards:
        seen[s.global_index] += 1
        # values are a pure function of the global node id
        assert s.local_m16.tobytes() == whole.local_m16[s.global_index].tobytes()
        assert (s.flags == whole.flags[s.global_index]).all()
        assert (s.render_mask == whole.render_mask[s.global_index]).all()
        # parents stay inside the shard and map to the right global parent
        p = s.parent.copy()
        ok = p != NONE
        assert (s.global_index[p[ok]] == whole.parent[s.global_index[ok]]).all()
    assert seen[0] == R  # the root is replicated
    assert (seen[1:] == 1).all()  # every other node lives on exactly one rank
    assert sum(s.n_units for s in shards) == 40
