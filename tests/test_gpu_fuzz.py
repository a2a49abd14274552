"""Randomised differential test: long random sequences of the calls a host makes between frames — transform changes, flag
toggles (visibility, enabled, frustum culling, shadows, static batching), render-mask and box changes, nodes added / removed /
re-parented (fyx_set_topology followed by the NEW nodes' matrices only), incremental and full updates, culls with 1–8 random
frusta, masks and shadow passes through the fused, stand-alone and one-call entry points — against an oracle REBUILT FROM SCRATCH
from the host's arrays after every step.  Whatever state the incremental paths carry (dirty bits, per-node columns moved to
new slots, prune bits, level plans) must never show: matrices / boxes bit-exact, visible sets identical, every step."""
import numpy as np
import pytest

import fyrox_b200 as fb
import oracle_binding as ob
from helpers import NONE, UNIT_BOX, assert_same_hierarchy, random_graph

pytestmark = pytest.mark.gpu


def reachable(parent, flags):
    n = len(parent)
    alive = (flags & fb.NODE_ALIVE) != 0
    reach = np.zeros(n, bool)
    reach[0] = True
    changed = True
    while changed:
        ok = alive & (parent != NONE)
        new = reach.copy()
        new[ok] |= reach[parent[ok]] & alive[parent[ok]]
        changed = bool((new != reach).any())
        reach = new
    return reach


def random_frusta(rng, k):
    fos, ffs = [], []
    while len(fos) < k:
        eye = rng.uniform(-40, 40, 3)
        tgt = eye + rng.normal(size=3)
        up = (0, 1, 0) if abs((tgt - eye)[1]) < 0.9 * np.linalg.norm(tgt - eye) else (1, 0, 0)
        view = ob.look_at_rh(tuple(eye), tuple(tgt), up)
        proj = ob.perspective(float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.4, 2.4)), float(rng.uniform(0.01, 1.0)), float(rng.uniform(20, 400)))
        vp = ob.mat4_mul(proj, view)
        fo = ob.frustum_from_vp(vp)
        ff = fb.frustum_from_view_projection_matrix(vp)
        if fo is None or ff is None:
            continue
        fos.append(fo)
        ffs.append(ff)
    return fos, ffs


def subtree_mask(parent, flags, x):
    """nodes of the sub-tree rooted at x (incl. x)"""
    n = len(parent)
    inside = np.zeros(n, bool)
    inside[x] = True
    changed = True
    alive = (flags & fb.NODE_ALIVE) != 0
    while changed:
        ok = alive & (parent != NONE)
        new = inside.copy()
        new[ok] |= inside[parent[ok]]
        changed = bool((new != inside).any())
        inside = new
    return inside


@pytest.mark.timeout(900)
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_call_sequences_match_an_oracle_rebuilt_from_scratch(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    n = 700
    parent, flags, mask, local, aabb = random_graph(rng, n, p_dead=0.25, p_orphan=0.0, max_depth_bias=0.15 * seed)
    ctx.set_topology(parent, flags, mask, aabb)
    ctx.set_local_matrices(local)
    ctx.update_transforms(fb.UPDATE_ALL)
    settable = fb.NODE_VISIBILITY | fb.NODE_ENABLED | fb.NODE_FRUSTUM_CULLING | fb.NODE_CAST_SHADOWS | fb.NODE_STATIC_BATCH
    for step in range(28):
        alive = np.nonzero((flags & fb.NODE_ALIVE) != 0)[0]
        topo = False
        new_nodes = np.empty(0, np.int64)
        # --- topology edits (every third step or so)
        if rng.random() < 0.35:
            topo = True
            dead = np.nonzero((flags & fb.NODE_ALIVE) == 0)[0]
            if dead.size:  # add nodes under alive parents
                new_nodes = rng.choice(dead, min(dead.size, int(rng.integers(1, 12))), replace=False)
                for i in new_nodes:
                    flags[i] = fb.NODE_DEFAULT | (fb.NODE_RENDERABLE if rng.random() < 0.7 else 0)
                    parent[i] = rng.choice(alive)
                    local[i] = ob.translation(*rng.uniform(-6, 6, 3))
                    h = rng.uniform(0.2, 3.0, 3).astype(np.float32)
                    aabb[i] = np.concatenate([-h, h]) if flags[i] & fb.NODE_RENDERABLE else UNIT_BOX
                    mask[i] = 0xFFFFFFFF if rng.random() < 0.8 else (1 << int(rng.integers(0, 32)))
            # remove whole sub-trees (Graph::remove_node takes the descendants along)
            for _ in range(int(rng.integers(0, 3))):
                cand = alive[alive != 0]
                if not cand.size:
                    break
                x = int(rng.choice(cand))
                sub = subtree_mask(parent, flags, x)
                if sub.sum() > 40 or not (flags[x] & fb.NODE_ALIVE):
                    continue
                flags[sub] = 0
                parent[sub] = NONE
            # re-parent a few nodes (never under their own sub-tree)
            alive = np.nonzero((flags & fb.NODE_ALIVE) != 0)[0]
            for _ in range(int(rng.integers(0, 4))):
                cand = alive[alive != 0]
                if cand.size < 2:
                    break
                x = int(rng.choice(cand))
                sub = subtree_mask(parent, flags, x)
                targets = alive[~sub[alive]]
                if targets.size:
                    parent[x] = int(rng.choice(targets))
        alive = np.nonzero((flags & fb.NODE_ALIVE) != 0)[0]
        # --- transform / flag / mask / box changes on surviving nodes
        moved = rng.choice(alive[alive != 0], min(alive.size - 1, int(rng.integers(0, 40))), replace=False) if alive.size > 1 else np.empty(0, np.int64)
        for i in moved:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            t = ob.Transform()
            ob.lib().orc_transform_identity(t)
            t.local_position[:] = rng.uniform(-15, 15, 3).astype(np.float32).tolist()
            t.local_rotation[:] = q.astype(np.float32).tolist()
            t.local_scale[:] = rng.uniform(0.5, 1.6, 3).astype(np.float32).tolist()
            ob.lib().orc_transform_calculate_local(t, ob.fp(local[i]))
        toggled = rng.choice(alive[alive != 0], min(alive.size - 1, int(rng.integers(0, 25))), replace=False) if alive.size > 1 else np.empty(0, np.int64)
        for i in toggled:
            bit = int(rng.choice([fb.NODE_VISIBILITY, fb.NODE_ENABLED, fb.NODE_FRUSTUM_CULLING, fb.NODE_CAST_SHADOWS, fb.NODE_STATIC_BATCH]))
            if bit == fb.NODE_STATIC_BATCH and not (flags[i] & fb.NODE_RENDERABLE):
                continue
            flags[i] ^= np.uint32(bit)
        remask = rng.choice(alive, min(alive.size, int(rng.integers(0, 10))), replace=False)
        for i in remask:
            mask[i] = int(rng.integers(0, 1 << 32, dtype=np.uint64))
        rebox = rng.choice(alive, min(alive.size, int(rng.integers(0, 10))), replace=False)
        rebox = rebox[(flags[rebox] & fb.NODE_RENDERABLE) != 0]
        for i in rebox:
            h = rng.uniform(0.1, 4.0, 3).astype(np.float32)
            aabb[i] = np.concatenate([-h, h])
        # --- the host's calls
        full = rng.random() < 0.25
        if topo:
            ctx.set_topology(parent, flags, mask, aabb)
            if new_nodes.size:
                ctx.set_local_matrices(local[new_nodes], new_nodes.astype(np.uint32))  # ONLY the new nodes
        live_moved = np.array([i for i in moved if flags[i] & fb.NODE_ALIVE], np.uint32)
        if live_moved.size:
            ctx.set_local_matrices(local[live_moved], live_moved)
        if not topo:
            lt = np.array([i for i in toggled if flags[i] & fb.NODE_ALIVE], np.uint32)
            if lt.size:
                ctx.set_flags(flags[lt], lt)
            if remask.size:
                ctx.set_render_masks(mask[remask], remask.astype(np.uint32))
            if rebox.size:
                ctx.set_local_aabbs(aabb[rebox], rebox.astype(np.uint32))
        k = int(rng.choice([1, 2, 3, 4, 6, 8]))
        fos, ffs = random_frusta(rng, k)
        cam = rng.choice(np.array([0xFFFFFFFF, 0x0000FFFF, 0xFFFF0000, 0x0F0F0F0F], np.uint32), k)
        pf = (rng.random(k) < 0.3).astype(np.uint32) * np.uint32(fb.PASS_SHADOW)
        entry = int(rng.integers(0, 3))
        uf = fb.UPDATE_ALL if full else fb.UPDATE_INCREMENTAL
        if entry == 0:
            ctx.update_and_cull(ffs, uf, cam_mask=cam, pass_flags=pf)
        elif entry == 1:
            ctx.render_prep(update_flags=uf, frusta=ffs, cam_mask=cam, pass_flags=pf)
        else:
            ctx.update_transforms(uf)
            ctx.cull(ffs, cam_mask=cam, pass_flags=pf)
        # --- the oracle, from scratch
        og = ob.Graph.build(parent, flags, mask, local, aabb)
        og.L.orc_graph_drop_messages(og.h)
        og.update_hierarchical_data()
        live = np.nonzero(reachable(parent, flags))[0].astype(np.uint32)
        assert_same_hierarchy(og, ctx, live)
        for f, fo in enumerate(fos):
            want = np.sort(og.from_graph(fo, int(cam[f]), bool(pf[f] & fb.PASS_SHADOW)))
            got = np.sort(ctx.get_visible(f))
            assert np.array_equal(got, want), f"seed {seed} step {step} entry {entry} frustum {f}/{k}: {got.size} vs {want.size}"
        og.free()
