"""Sampled-oracle parity at sizes the oracle cannot walk whole (BASELINE.json's C2–C5).

Test infrastructure (it executes oracle/): used by the `-m gpu` full-size tests and by bench.py's `parity` block,
which runs AFTER the timed regions on the very context that was timed.  Nothing here is on the product path.

For a random sample of nodes the exact reference values are rebuilt from the scene's inputs with the oracle's own
functions, independent of the GPU's slot order / level kernels:
  * global matrix   = product of the local matrices down the parent chain (Graph::update_global_transform_recursively,
                      scene/graph/mod.rs:1199-1241) through orc_mat4_mul;
  * flags           = AND down the chain (graph/mod.rs:1166-1197);
  * world box       = orc_aabb_transform (fyrox-math/src/aabb.rs:264-287), for skinned meshes followed by
                      orc_aabb_add_point of every bone's global position in bone order (scene/mesh/mod.rs:673-684);
  * visible bit/f   = NodeTrait::should_be_rendered (scene/node/mod.rs:231-256) with orc_frustum_is_intersects_aabb;
  * palette, LBS    = orc_mat4_mul(G_bone, inv_bind) and orc_skin_vertices (scene/mesh/mod.rs:501-522, 781-793).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import fyrox_b200 as fb
import oracle_binding as ob

NONE = 0xFFFFFFFF


def bits_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


class SampledTruth:
    """Exact per-node reference values of one (possibly sharded) generated scene.

    bone_idx / bone_local: the animated nodes of the frame being checked and their local matrices — either an (n,16)
    array or a callable k -> m16 (evaluated lazily, only for the bones a sampled chain touches).
    mesh_aabb: {node: aabb6} overrides of local boxes (Mesh::local_bounding_box of skinned meshes = vertex bounds).
    """

    def __init__(self, sc, bone_idx=None, bone_local=None, mesh_aabb=None, root: int = 0):
        self.sc = sc
        self.L = ob.lib()
        self.root = root
        self.pos_of = None
        if bone_idx is not None and len(bone_idx):
            self.pos_of = np.full(sc.capacity, -1, np.int64)
            self.pos_of[np.asarray(bone_idx, np.int64)] = np.arange(len(bone_idx))
        self.bone_local = bone_local
        self.mesh_aabb = mesh_aabb or {}
        self._g = {}
        self._f = {}

    def local(self, n: int) -> np.ndarray:
        if self.pos_of is not None:
            k = int(self.pos_of[n])
            if k >= 0:
                m = self.bone_local(k) if callable(self.bone_local) else self.bone_local[k]
                return np.ascontiguousarray(m, np.float32).reshape(16)
        return self.sc.local_m16[n]

    def chain(self, n: int):
        """(G, global_visibility, global_enabled, reachable from the root) of node n, memoised."""
        hit = self._g.get(n)
        if hit is not None:
            return hit
        p = int(self.sc.parent[n])
        if p == NONE:
            pg, pv, pe, pr = np.eye(4, dtype=np.float32).reshape(16), True, True, (n == self.root)
        else:
            pg, pv, pe, pr = self.chain(p)
        f = int(self.sc.flags[n])
        out = (ob.mat4_mul(pg, self.local(n)), pv and bool(f & fb.NODE_VISIBILITY), pe and bool(f & fb.NODE_ENABLED), pr)
        self._g[n] = out
        return out

    def node(self, i: int, frusta_o, cam_mask=0xFFFFFFFF, bones=None):
        """(G m16, world box aabb6, [visible in frustum f]) of node i; `bones` = its bone nodes if it is a skinned mesh."""
        g, gv, ge, reach = self.chain(int(i))
        la = self.mesh_aabb.get(int(i), self.sc.local_aabb[i])
        box = ob.Aabb.make(la[:3], la[3:])
        out = ob.Aabb()
        self.L.orc_aabb_transform(C.byref(box), ob.fp(np.ascontiguousarray(g)), C.byref(out))
        if bones is not None:
            for b in bones:
                gb = self.chain(int(b))[0]
                self.L.orc_aabb_add_point(C.byref(out), ob.fp(np.ascontiguousarray(gb[12:15])))
        f = int(self.sc.flags[i])
        base = bool(f & fb.NODE_RENDERABLE) and gv and ge and reach and (int(self.sc.render_mask[i]) & cam_mask) != 0
        vis = []
        for fo in frusta_o:
            ok = base
            if ok and (f & fb.NODE_FRUSTUM_CULLING):
                ok = bool(self.L.orc_frustum_is_intersects_aabb(C.byref(fo), C.byref(out)))
            vis.append(ok)
        return g, out.to_np(), vis

    def unit(self, u: int, verts: np.ndarray):
        """(palette (B,16), skinned positions (V,3), normals (V,3)) of skinned unit u from its AnimatedVertex bytes."""
        sc = self.sc
        bones = sc.unit_bone_nodes(u)
        ib = sc.unit_inv_bind(u)
        pal = np.stack([ob.mat4_mul(self.chain(int(b))[0], ib[k]) for k, b in enumerate(bones)])
        nv = sc.verts_per_unit
        pos = np.empty((nv, 3), np.float32)
        nrm = np.empty((nv, 3), np.float32)
        lay = ob.ANIMATED_VERTEX
        self.L.orc_skin_vertices(ob.fp(np.ascontiguousarray(pal.reshape(-1))), nv, verts.ctypes.data_as(C.c_void_p), C.byref(lay), ob.fp(pos.reshape(-1)),
                                 ob.fp(nrm.reshape(-1)))
        return pal, pos, nrm


def trs_bone_local(trs10: np.ndarray):
    """k -> Transform::calculate_local_transform (scene/transform.rs:421-540, the oracle's restatement) of TRS record k:
    what the device evaluates for `changed_trs` / `changed_rot` uploads."""
    L = ob.lib()

    def f(k: int) -> np.ndarray:
        t = ob.Transform()
        L.orc_transform_identity(t)
        r = trs10[k]
        t.local_position[:] = [float(x) for x in r[0:3]]
        t.local_rotation[:] = [float(x) for x in r[3:7]]
        t.local_scale[:] = [float(x) for x in r[7:10]]
        m = np.empty(16, np.float32)
        L.orc_transform_calculate_local(t, ob.fp(m))
        return m

    return f


def pick_sample(sc, rng, n_random: int, visible_lists_local=(), per_list: int = 300, extra=()):
    """Local node indices to check: random ones + the head/tail of every visible list (so that positives are covered)
    + whatever the caller adds (skinned mesh nodes, bones)."""
    parts = [rng.integers(0, sc.capacity, n_random).astype(np.int64)]
    for v in visible_lists_local:
        if len(v):
            parts.append(np.asarray(v[:per_list], np.int64))
            parts.append(np.asarray(v[-per_list:], np.int64))
    if len(extra):
        parts.append(np.asarray(extra, np.int64))
    return np.unique(np.concatenate(parts)).astype(np.uint32)


class SortedList:
    """One sort of a visible list, then duplicate check and membership of samples by binary search (np.isin / np.unique would
    sort the list again for every query: the lists hold up to 10^8 entries)."""

    def __init__(self, lst):
        self.a = np.sort(np.asarray(lst, dtype=np.uint32))
        self.duplicate_free = bool(self.a.size < 2 or (self.a[1:] != self.a[:-1]).all())

    def contains(self, q) -> np.ndarray:
        q = np.asarray(q, dtype=np.uint32)
        if not self.a.size:
            return np.zeros(q.shape, bool)
        i = np.searchsorted(self.a, q)
        return (i < self.a.size) & (self.a[np.minimum(i, self.a.size - 1)] == q)


def check_nodes(ctx, truth: SampledTruth, sample: np.ndarray, frusta_o, own_lists_gidx, skinned_bones=None, cam_mask=0xFFFFFFFF):
    """Compare the context with the sampled truth.  own_lists_gidx[f] = this context's visible list of frustum f (global
    indices as emitted).  Returns a dict of counters / booleans and the per-sample expected bits (for N>1 union checks)."""
    sc = truth.sc
    skinned_bones = skinned_bones or {}
    G = np.empty((len(sample), 16), np.float32)
    A = np.empty((len(sample), 6), np.float32)
    vis = np.zeros((len(sample), len(frusta_o)), bool)
    for k, i in enumerate(sample):
        G[k], A[k], vis[k] = truth.node(int(i), frusta_o, cam_mask, skinned_bones.get(int(i)))
    Gg = ctx.get_global_matrices(sample)
    Ag = ctx.get_world_aabbs(sample)
    g_ok = bits_equal(Gg, G).all(axis=1)
    a_ok = bits_equal(Ag, A).all(axis=1)
    gid = sc.global_index[sample]
    v_ok = np.ones(len(sample), bool)
    dup_free = True
    for f in range(len(frusta_o)):
        sl = SortedList(own_lists_gidx[f])
        dup_free = dup_free and sl.duplicate_free
        v_ok &= sl.contains(gid) == vis[:, f]
    return {
        "checked_nodes": int(len(sample)),
        "global_matrices_bit_exact": bool(g_ok.all()),
        "world_aabbs_bit_exact": bool(a_ok.all()),
        "visible_set_equal": bool(v_ok.all() and dup_free),
        "bad_nodes": [int(x) for x in sample[~(g_ok & a_ok & v_ok)][:8]],
        "sample_gid": gid,
        "sample_vis": vis,
    }


def check_units(ctx, truth: SampledTruth, units, sids=None):
    """Palettes and skinned streams of the given units (surface id = unit index unless sids says otherwise)."""
    sc = truth.sc
    worst = 0.0
    exact = True
    verts_checked = 0
    for u in units:
        verts, _ = sc.unit_vertices(u)
        pal, pos, nrm = truth.unit(u, verts)
        sid = u if sids is None else sids[u]
        exact &= bool(bits_equal(ctx.get_palette(sid), pal).all())
        pg, ng = ctx.get_skinned(sid)
        worst = max(worst, float(np.abs(pg - pos).max()))
        exact &= pg.tobytes() == pos.tobytes() and ng.tobytes() == nrm.tobytes()
        verts_checked += sc.verts_per_unit
    return {"checked_units": len(list(units)), "checked_verts": int(verts_checked), "skinning_bit_exact": bool(exact), "max_abs_pos_err": worst}


def list_checksum(lst: np.ndarray):
    """(count, sum, xor) of a visible list: a checksum of checksums lets rank 0 verify its gathered list is exactly the
    disjoint union of every rank's own list without moving them again."""
    a = np.asarray(lst, np.uint64)
    return int(a.size), int(a.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(a)) if a.size else 0


def check_frame(ctx, sc, frusta_ff, frusta_o, upload: str, bone_idx, bone_payload, allgather: bool = False, seed: int = 0, n_random: int = 5000,
                n_units_exact: int = 4, n_mesh_nodes: int = 60):
    """Run ONE synchronous all-dirty frame with the given animation frame on `ctx` (already loaded with scene `sc`, surface id
    = unit index) and compare it with the sampled truth: >= n_random nodes (matrix / box bit-exact, per-frustum visibility
    identical) including skinned-mesh nodes with their bone folds, and n_units_exact skinned meshes bit-exact.
    upload: "m16" (bone_payload = (n,16) local matrices), "trs" ((n,10) records) or "rot" ((n,10) records whose rotations are
    uploaded; the device must already hold their positions / scales).  Returns (summary dict, raw check_nodes result)."""
    nb = sc.n_units * sc.bones_per_unit
    kw, truth_kw = {}, {}
    if nb:
        bone_idx = np.ascontiguousarray(bone_idx, np.uint32)
        if upload == "m16":
            kw = dict(changed_m16=np.ascontiguousarray(bone_payload, np.float32), changed_idx=bone_idx)
            truth_kw = dict(bone_idx=bone_idx, bone_local=bone_payload)
        else:
            trs = np.ascontiguousarray(bone_payload, np.float32)
            kw = dict(changed_trs=trs, changed_idx=bone_idx) if upload == "trs" else dict(changed_rot=np.ascontiguousarray(trs[:, 3:7]), changed_idx=bone_idx)
            truth_kw = dict(bone_idx=bone_idx, bone_local=trs_bone_local(trs))
    ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=frusta_ff, readback_visible=True, allgather=allgather, **kw)
    own = [ctx.get_visible(f) for f in range(len(frusta_ff))]
    rng = np.random.default_rng(seed)
    nu = sc.n_units
    units = sorted(set([0, nu - 1] + [int(x) for x in rng.integers(0, nu, max(n_units_exact - 2, 0))])) if nu else []
    mesh_units = sorted(set(units + [int(x) for x in rng.integers(0, nu, n_mesh_nodes)])) if nu else []
    mesh_aabb, bones_of, extra = {}, {}, []
    for u in mesh_units:
        mesh = int(sc.unit_mesh_node(u))
        mesh_aabb[mesh] = sc.unit_vertices(u)[1]
        bones_of[mesh] = sc.unit_bone_nodes(u).copy()
        extra += [mesh] + [int(b) for b in bones_of[mesh][:8]]
    truth = SampledTruth(sc, mesh_aabb=mesh_aabb, **truth_kw)
    # positives: a few entries of every list (global indices) mapped back to local nodes
    gi = sc.global_index
    ident = bool(gi.size) and int(gi[-1]) == gi.size - 1 and int(gi[gi.size // 2]) == gi.size // 2
    order = None if ident else np.argsort(gi, kind="stable")
    pos = []
    for v in own:
        g = np.concatenate([v[:300], v[-300:]]) if v.size else v
        pos.append(g.astype(np.int64) if ident else order[np.searchsorted(gi[order], g)])
    sample = pick_sample(sc, rng, n_random, pos, extra=extra)
    if nu:  # a skinned-mesh node needs its bones for the fold: keep only those whose unit was collected above
        all_mesh = np.fromiter((sc.unit_mesh_node(u) for u in range(nu)), dtype=np.uint32, count=nu)
        known = np.fromiter(bones_of.keys(), dtype=np.uint32, count=len(bones_of))
        sample = sample[~np.isin(sample, all_mesh) | np.isin(sample, known)]
    res = check_nodes(ctx, truth, sample, frusta_o, own, skinned_bones=bones_of)
    ures = check_units(ctx, truth, units) if units else {"checked_units": 0, "checked_verts": 0, "skinning_bit_exact": True, "max_abs_pos_err": 0.0}
    out = {k: v for k, v in res.items() if k not in ("sample_gid", "sample_vis")}
    out.update(ures)
    out["frusta"] = len(frusta_ff)
    out["visible_entries_own"] = int(sum(v.size for v in own))
    out["ok"] = bool(out["global_matrices_bit_exact"] and out["world_aabbs_bit_exact"] and out["visible_set_equal"] and out["skinning_bit_exact"])
    res["own_sums"] = [list_checksum(v) for v in own]
    return out, res
