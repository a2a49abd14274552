"""Parity at BASELINE.json's full sizes, where the oracle cannot walk the whole scene in seconds:
size-independent properties (fused == separate cull, idempotence, shards partition the visible set) plus
exact oracle checks on random SAMPLES (global matrices by chain products, boxes, the cull predicate,
palettes and skinned vertices)."""
import ctypes as C

import numpy as np
import pytest

import fyrox_b200 as fb
import oracle_binding as ob
from fyrox_b200 import camera
from fyrox_b200.scenegen import Scene
from helpers import bits_equal, camera_frustum, cube_frusta

pytestmark = pytest.mark.gpu
NONE = 0xFFFFFFFF


def load(ctx, sc):
    ctx.set_topology(sc.parent, sc.flags, sc.render_mask, sc.local_aabb, root=0, global_index=sc.global_index)
    ctx.set_local_matrices(sc.local_m16)


def oracle_sample(sc, idx, frusta_o, cam_mask=0xFFFFFFFF):
    """Exact reference values for the sampled nodes: G by chain products, world box, flags, per-frustum visibility."""
    L = ob.lib()
    G = np.empty((len(idx), 16), np.float32)
    A = np.empty((len(idx), 6), np.float32)
    vis = np.zeros((len(idx), len(frusta_o)), bool)
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for k, i in enumerate(idx):
        chain = []
        n = int(i)
        while n != NONE:
            chain.append(n)
            n = int(sc.parent[n])
        g = ident.copy()
        gv = ge = True
        for n in reversed(chain):  # root first: G = parent.G * local
            g = ob.mat4_mul(g, sc.local_m16[n])
            gv = gv and bool(sc.flags[n] & fb.NODE_VISIBILITY)
            ge = ge and bool(sc.flags[n] & fb.NODE_ENABLED)
        G[k] = g
        box = ob.Aabb.make(sc.local_aabb[i, :3], sc.local_aabb[i, 3:])
        out = ob.Aabb()
        L.orc_aabb_transform(C.byref(box), ob.fp(np.ascontiguousarray(g)), C.byref(out))
        A[k] = out.to_np()
        f = int(sc.flags[i])
        base = bool(f & fb.NODE_RENDERABLE) and gv and ge and chain[-1] == 0 and (int(sc.render_mask[i]) & cam_mask) != 0
        for q, fo in enumerate(frusta_o):
            ok = base
            if ok and (f & fb.NODE_FRUSTUM_CULLING):
                ok = bool(L.orc_frustum_is_intersects_aabb(C.byref(fo), C.byref(out)))
            vis[k, q] = ok
    return G, A, vis


@pytest.mark.timeout(900)
def test_c2_ten_million_nodes(ctx):
    """configs[1]: 10 M static nodes, 1 frustum (plus the 6 cube faces)."""
    sc = Scene(10_000_000)
    load(ctx, sc)
    fo1, ff1 = camera_frustum()
    fos6, ffs6 = cube_frusta()
    fos, ffs = [fo1] + fos6, [ff1] + ffs6
    ctx.update_and_cull(ffs, fb.UPDATE_ALL)
    fused = [np.sort(ctx.get_visible(f)) for f in range(len(ffs))]
    assert all(v.size == np.unique(v).size for v in fused)
    # separate update + cull gives the same sets; culling again is idempotent
    ctx.update_transforms(fb.UPDATE_ALL)
    ctx.cull(ffs)
    sep = [np.sort(ctx.get_visible(f)) for f in range(len(ffs))]
    ctx.cull(ffs)
    again = [np.sort(ctx.get_visible(f)) for f in range(len(ffs))]
    for f in range(len(ffs)):
        assert np.array_equal(fused[f], sep[f]) and np.array_equal(sep[f], again[f])
    assert 0 < fused[0].size < sc.n_renderable
    # every renderable, visible-flagged node inside the cube's range shows up in at least one face: spot check via samples
    rng = np.random.default_rng(3)
    sample = np.unique(np.concatenate([rng.integers(0, sc.capacity, 6000), fused[0][:2000], fused[3][:2000]])).astype(np.uint32)
    G, A, vis = oracle_sample(sc, sample, fos)
    assert bits_equal(ctx.get_global_matrices(sample), G).all()
    assert bits_equal(ctx.get_world_aabbs(sample), A).all()
    for f in range(len(ffs)):
        assert np.array_equal(np.isin(sample, fused[f]), vis[:, f]), f"frustum {f}"
    # an incremental update with nothing changed moves nothing
    ctx.update_and_cull(ffs, fb.UPDATE_INCREMENTAL)
    for f in range(len(ffs)):
        assert np.array_equal(np.sort(ctx.get_visible(f)), fused[f])
    # sharded: two contexts, each with half of the sectors; the union of their lists is the whole list
    parts = []
    for r in range(2):
        sh = Scene(10_000_000, rank=r, nranks=2)
        with fb.Context() as c2:
            load(c2, sh)
            c2.update_and_cull(ffs, fb.UPDATE_ALL)
            parts.append([c2.get_visible(f) for f in range(len(ffs))])
        sh.close()
    for f in range(len(ffs)):
        assert np.array_equal(np.sort(np.concatenate([parts[0][f], parts[1][f]])), fused[f])


@pytest.mark.timeout(900)
def test_skinning_at_scale_sampled(ctx):
    """2 000 skinned meshes x 64 bones x 5 000 verts (10 M verts) inside a 1 M-node scene: palettes and skinned
    streams of sampled meshes equal the oracle bit for bit; the rest is covered by a checksum of two runs."""
    sc = Scene(1_000_000, n_units=2000, verts_per_unit=5000)
    load(ctx, sc)
    ctx.reserve_skinning(sc.n_units * 64, sc.n_units * 5000)
    sids = []
    sample_units = [0, 1, 777, 1999]
    kept = {}
    for u in range(sc.n_units):
        verts, bb = sc.unit_vertices(u)
        sids.append(ctx.add_skinned_surface(sc.unit_mesh_node(u), sc.unit_bone_nodes(u), sc.unit_inv_bind(u), verts))
        if u in sample_units:
            kept[u] = verts
    idx, trs = sc.animate_trs(2)
    _, m16 = sc.animate(2)
    ctx.render_prep(update_flags=fb.UPDATE_ALL, changed_m16=m16, changed_idx=idx, frusta=camera.cube_frusta(), readback_visible=True)
    L = ob.lib()
    lay = ob.ANIMATED_VERTEX
    for u in sample_units:
        bones = sc.unit_bone_nodes(u)
        Gb = ctx.get_global_matrices(bones)
        ib = sc.unit_inv_bind(u)
        pal_o = np.stack([ob.mat4_mul(Gb[k], ib[k]) for k in range(len(bones))])
        pal_g = ctx.get_palette(sids[u])
        assert bits_equal(pal_g, pal_o).all()
        pos_o = np.empty((5000, 3), np.float32)
        nrm_o = np.empty((5000, 3), np.float32)
        L.orc_skin_vertices(ob.fp(np.ascontiguousarray(pal_o.reshape(-1))), 5000, kept[u].ctypes.data_as(C.c_void_p), C.byref(lay), ob.fp(pos_o.reshape(-1)), ob.fp(nrm_o.reshape(-1)))
        pos_g, nrm_g = ctx.get_skinned(sids[u])
        assert np.abs(pos_g - pos_o).max() <= 1e-5
        assert pos_g.tobytes() == pos_o.tobytes() and nrm_g.tobytes() == nrm_o.tobytes()
    # determinism: the same frame again (as separate calls) reproduces every stream bit for bit
    chk = [ctx.get_skinned(sids[u])[0].copy() for u in (5, 1000, 1500)]
    ctx.set_local_matrices(m16, idx)
    ctx.update_transforms(fb.UPDATE_INCREMENTAL)
    ctx.build_palettes()
    ctx.skin()
    for u, ref in zip((5, 1000, 1500), chk):
        assert ctx.get_skinned(sids[u])[0].tobytes() == ref.tobytes()


def _full_config(ctx, nodes, units, n_frusta, upload):
    """Load a BASELINE.json config exactly as bench.py does (same generator, seed, loader), run one all-dirty frame and
    check it with the sampled oracle (tests/sampled_parity.py): >= 5 000 nodes + 4 skinned meshes, bit for bit."""
    import sys

    import sampled_parity as sp

    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    import bench

    sc = Scene(nodes, n_units=units, verts_per_unit=5000, bones_per_unit=64, seed=bench.SEED)
    bench.load_scene(ctx, sc, fb, lambda m: None)
    ffs = [camera.camera_frustum()] if n_frusta == 1 else camera.cube_frusta()[:n_frusta]
    fos = bench.oracle_frusta(n_frusta)
    idx, trs = sc.animate_trs(0)
    if upload == "rot":
        ctx.set_local_trs(trs, idx)  # the device keeps position / scale; the frame uploads rotations only
        idx, trs = sc.animate_trs(1)
    out, res = sp.check_frame(ctx, sc, ffs, fos, upload, idx, trs, seed=11)
    assert out["checked_nodes"] >= 5000 and out["checked_units"] >= 3, out
    assert out["ok"], out
    assert out["max_abs_pos_err"] <= 1e-5  # north_star's tolerance; the compare above is bit-exact
    assert 0 < out["visible_entries_own"]
    # the one-call frame and the separate calls agree on every list (size-independent property)
    one = [np.sort(ctx.get_visible(f)) for f in range(n_frusta)]
    ctx.update_and_cull(ffs, fb.UPDATE_ALL)
    for f in range(n_frusta):
        assert np.array_equal(np.sort(ctx.get_visible(f)), one[f])
    sc.close()
    return out


@pytest.mark.timeout(900)
def test_c3_full(ctx):
    """configs[2]: 1 M nodes incl. 10 k skinned meshes x 64 bones x 5 k verts (50 M verts), 1 frustum — at its own size."""
    _full_config(ctx, 1_000_000, 10_000, 1, "trs")


@pytest.mark.timeout(1200)
def test_c4_full(ctx):
    """configs[3] (the BENCH config): 10 M nodes, 50 k skinned meshes (250 M verts), 6 cube-face frusta — at its own size,
    with the bench's default upload format (16-byte rotations)."""
    _full_config(ctx, 10_000_000, 50_000, 6, "rot")
