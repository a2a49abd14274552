"""N2 — animation sampling on the device: fyx_animate against the oracle's restatement of
AnimationContainer::update_animations (scene/animation/mod.rs:83-88), Animation::tick (fyrox-animation/src/lib.rs:
471-496, 895-914), TrackDataContainer::fetch (container.rs:162-301), Curve::value_at (fyrox-math/src/curve.rs:252-309)
and the pose application (scene/animation/mod.rs:147-179, scene/transform.rs:202-262).

Vector3 / UnitQuaternion tracks: bit-exact global transforms and time positions over many frames (the span hints and
the Transform's dirty-or-different rule are state, so one wrong bit anywhere shows up later).  UnitQuaternionEuler
tracks go through sin/cos, which the reference takes from the platform libm: compared at 2e-6 absolute on the
rotation-matrix entries of depth-1 nodes.
"""
import ctypes as C

import numpy as np
import pytest

import fyrox_b200 as fb
import oracle_binding as ob
from helpers import NONE, UNIT_BOX, assert_same_hierarchy, bits_equal, camera_frustum

pytestmark = pytest.mark.gpu

KD, TD = ob.KEY_DTYPE, ob.TRACK_DTYPE


class Builder:
    """Builds the (tracks, keys) arrays of one animation."""

    def __init__(self, rng):
        self.rng = rng
        self.keys = []
        self.tracks = []

    def curve(self, n_keys, lo=-3.0, hi=3.0, grid=0.25, kinds=(0, 1, 2)):
        rng = self.rng
        first = len(self.keys)
        # locations on a 0.25 grid (time positions step by 0.125: they land exactly on keys half the time), duplicates allowed
        locs = np.sort(rng.integers(0, 10, n_keys)).astype(np.float32) * np.float32(grid)
        for loc in locs:
            kind = int(rng.choice(kinds))
            lt, rt = (float(np.float32(np.tan(rng.uniform(-1.2, 1.2)))), float(np.float32(np.tan(rng.uniform(-1.2, 1.2))))) if kind == 2 else (0.0, 0.0)
            self.keys.append((float(loc), float(np.float32(rng.uniform(lo, hi))), kind, lt, rt))
        return first, n_keys

    def track(self, node, binding, kind, n_curves=None, enabled=1, key_range=(0, 7), **kw):
        need = {ob.TV_REAL: 1, ob.TV_VECTOR2: 2, ob.TV_VECTOR3: 3, ob.TV_VECTOR4: 4, ob.TV_QUAT_EULER: 3, ob.TV_QUAT: 4}[kind]
        nc = need if n_curves is None else n_curves
        t = np.zeros((), TD)
        t["target_node"], t["binding"], t["value_kind"], t["enabled"], t["n_curves"] = node, binding, kind, enabled, nc
        for c in range(nc):
            if kind == ob.TV_QUAT and c == 3:  # keep the quaternion away from 0 (0/0 on normalisation is not a parity case)
                f, n = self.curve(int(self.rng.integers(1, 7)), lo=0.6, hi=1.0, kinds=kw.get("kinds", (0, 1, 2)))
            else:
                f, n = self.curve(int(self.rng.integers(key_range[0], key_range[1])), **kw)
            t["first_key"][c], t["n_keys"][c] = f, n
        self.tracks.append(t)

    def arrays(self):
        k = np.zeros(len(self.keys), KD)
        for i, r in enumerate(self.keys):
            k[i] = r
        return np.array(self.tracks, TD) if self.tracks else np.zeros(0, TD), k


def make_graph(rng, n, depth1=False):
    parent = np.full(n, NONE, np.uint32)
    if depth1:
        parent[1:] = 0
    else:
        parent[1:] = (rng.random(n - 1) * np.arange(1, n)).astype(np.uint32)
    flags = np.full(n, fb.NODE_DEFAULT | fb.NODE_RENDERABLE, np.uint32)
    flags[0] = fb.NODE_DEFAULT
    pos = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    rot = rng.normal(size=(n, 4)).astype(np.float32)
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    scale = rng.uniform(0.5, 1.5, (n, 3)).astype(np.float32)
    pos[0], rot[0], scale[0] = 0, (0, 0, 0, 1), 1
    return parent, flags, np.concatenate([pos, rot, scale], axis=1)


def oracle_transforms(trs, statics=None):
    n = len(trs)
    arr = (ob.Transform * n)()
    for i in range(n):
        t = arr[i]
        ob.lib().orc_transform_identity(C.byref(t))
        t.local_position[:] = trs[i, 0:3].tolist()
        t.local_rotation[:] = trs[i, 3:7].tolist()
        t.local_scale[:] = trs[i, 7:10].tolist()
        if statics is not None:
            s = statics[i]
            t.pre_rotation[:] = s[0:4].tolist()
            t.post_rotation_matrix[:] = s[4:13].tolist()
            t.rotation_offset[:] = s[13:16].tolist()
            t.rotation_pivot[:] = s[16:19].tolist()
            t.scaling_offset[:] = s[19:22].tolist()
            t.scaling_pivot[:] = s[22:25].tolist()
    return arr


def load_pair(ctx, parent, flags, trs, statics=None):
    n = len(parent)
    tr = oracle_transforms(trs, statics)
    local = np.empty((n, 16), np.float32)
    for i in range(n):
        ob.lib().orc_transform_calculate_local(C.byref(tr[i]), ob.fp(local[i]))
    aabb = np.tile(UNIT_BOX, (n, 1))
    og = ob.Graph.build(parent, flags, None, local, aabb)
    og.update_hierarchical_data()
    ctx.set_topology(parent, flags, None, aabb)
    if statics is not None:
        ctx.set_transform_statics(statics)
    ctx.set_local_trs(trs)
    ctx.update_transforms(fb.UPDATE_ALL)
    assert_same_hierarchy(og, ctx)
    return og, tr


def times_equal(ctx, anims):
    got = ctx.anim_time_positions(0, len(anims))
    want = np.array([a.time_position for a in anims], np.float32)
    assert bits_equal(got, want).all(), (got, want)


@pytest.mark.parametrize("with_statics", [False, True])
def test_animation_players_match_oracle_bit_for_bit(ctx, with_statics):
    rng = np.random.default_rng(101 + with_statics)
    n = 700
    parent, flags, trs = make_graph(rng, n)
    dead = [40, 41]  # free pool records: a track bound to one is logged and skipped
    flags[dead] = 0
    parent[np.isin(parent, dead)] = 0
    statics = None
    if with_statics:
        statics = np.zeros((n, 25), np.float32)
        q = rng.normal(size=(n, 4)).astype(np.float32)
        statics[:, 0:4] = q / np.linalg.norm(q, axis=1, keepdims=True)
        statics[:, 4:13] = rng.normal(size=(n, 9)).astype(np.float32)
        statics[:, 13:25] = rng.uniform(-2, 2, (n, 12)).astype(np.float32)
    og, tr = load_pair(ctx, parent, flags, trs, statics)

    # animation 0: looped over [0, 2], the whole "skeleton": position + rotation + scale tracks
    b0 = Builder(rng)
    for node in range(1, 260):
        b0.track(node, ob.BIND_POSITION, ob.TV_VECTOR3)
        b0.track(node, ob.BIND_ROTATION, ob.TV_QUAT, lo=-1.0, hi=1.0)
        if node % 3 == 0:
            b0.track(node, ob.BIND_SCALE, ob.TV_VECTOR3, lo=0.5, hi=2.0)
    # animation 1: clamped to [0.5, 1.5], plays backwards, overrides some of animation 0's nodes, plus oddities
    b1 = Builder(rng)
    for node in range(200, 330):
        b1.track(node, ob.BIND_ROTATION, ob.TV_QUAT, lo=-1.0, hi=1.0)
    b1.track(5, ob.BIND_POSITION, ob.TV_VECTOR3, enabled=0)                 # disabled binding
    b1.track(6, ob.BIND_POSITION, ob.TV_VECTOR3, n_curves=2)                # too few curves: fetch() is None
    b1.track(7, ob.BIND_ROTATION, ob.TV_VECTOR3)                            # wrong value type: logged, skipped
    b1.track(8, ob.BIND_POSITION, ob.TV_QUAT)                               # wrong value type
    b1.track(40, ob.BIND_POSITION, ob.TV_VECTOR3)                           # dead node
    b1.track(n + 50, ob.BIND_POSITION, ob.TV_VECTOR3)                       # handle out of range
    b1.track(9, ob.BIND_SCALE, ob.TV_VECTOR3, key_range=(0, 2))             # empty / single-key curves
    b1.track(10, ob.BIND_POSITION, ob.TV_VECTOR3, kinds=(0,))               # Constant keys: stepf
    b1.track(10, ob.BIND_POSITION, ob.TV_VECTOR3, kinds=(2,))               # same property twice in one animation: the later wins
    # animation 2: disabled at first
    b2 = Builder(rng)
    for node in range(100, 400, 7):
        b2.track(node, ob.BIND_POSITION, ob.TV_VECTOR3, kinds=(1, 2))
    specs = [(b0, dict(speed=1.0, looped=True, time_slice=(0.0, 2.0), time_position=0.0, enabled=True)),
             (b1, dict(speed=-0.75, looped=False, time_slice=(0.5, 1.5), time_position=1.25, enabled=True)),
             (b2, dict(speed=2.5, looped=True, time_slice=(0.25, 1.75), time_position=7.0, enabled=False))]
    anims = []
    for b, kw in specs:
        t, k = b.arrays()
        anims.append(ob.Animation(t, k, **kw))
        ctx.anim_add(t, k, **kw)
    times_equal(ctx, anims)  # set_time_position wrapped / clamped the initial positions

    _, ff = camera_frustum(zfar=500.0)
    dt = 0.125
    for frame in range(14):
        if frame == 5:
            ob.lib().orc_animation_set_enabled(anims[2].h, 1)
            ctx.anim_set_enabled(2, True)
        if frame == 8:
            ob.lib().orc_animation_set_track_enabled(anims[0].h, 3, 0)
            ctx.anim_set_track_enabled(0, 3, False)
            ob.lib().orc_animation_set_speed(anims[0].h, -3.0)
            ctx.anim_set_speed(0, -3.0)
            ob.lib().orc_animation_set_time_position(anims[1].h, -9.0)
            ctx.anim_set_time_position(1, -9.0)
        if frame == 10:
            dt = 0.3  # off the key grid
        ob.update_animations(anims, dt, og, tr)
        og.update()
        if frame % 2:
            ctx.animate(dt)
            ctx.update_and_cull([ff], fb.UPDATE_INCREMENTAL)
        else:  # the one-call frame ticks the animation players first
            ctx.render_prep(update_flags=fb.UPDATE_INCREMENTAL, frusta=[ff], do_palettes=False, do_skin=False, animate_dt=dt)
        times_equal(ctx, anims)
        assert_same_hierarchy(og, ctx)
    # topology change: slots move, the animation keeps its state (hints, time) and follows the nodes
    flags2 = flags.copy()
    flags2[[60, 61, 62]] = 0
    parent2 = parent.copy()
    parent2[np.isin(parent2, [60, 61, 62])] = 0
    cur = np.array([[*tr[i].local_position, *tr[i].local_rotation, *tr[i].local_scale] for i in range(n)], np.float32)
    local = np.empty((n, 16), np.float32)
    for i in range(n):
        ob.lib().orc_transform_calculate_local(C.byref(tr[i]), ob.fp(local[i]))
    aabb = np.tile(UNIT_BOX, (n, 1))
    og2 = ob.Graph.build(parent2, flags2, None, local, aabb)
    og2.update_hierarchical_data()
    ctx.set_topology(parent2, flags2, None, aabb)
    if statics is not None:
        ctx.set_transform_statics(statics)
    ctx.set_local_trs(cur)
    ctx.update_transforms(fb.UPDATE_ALL)
    assert_same_hierarchy(og2, ctx)
    for frame in range(3):
        ob.update_animations(anims, dt, og2, tr)
        og2.update()
        ctx.animate(dt)
        ctx.update_transforms(fb.UPDATE_INCREMENTAL)
        times_equal(ctx, anims)
        assert_same_hierarchy(og2, ctx)


def test_blend_groups_match_oracle_bit_for_bit(ctx):
    """fyx_anim_blend_group: BlendAnimations over PlayAnimation sources with constant weights in a one-layer, one-state
    machine (machine/mod.rs:344-382, node/blend.rs:136-166, pose.rs:41-101), next to a directly applied animation."""
    rng = np.random.default_rng(404)
    n = 420
    parent, flags, trs = make_graph(rng, n)
    og, tr = load_pair(ctx, parent, flags, trs)

    def build(nodes, pos=True, rot=True, scale=False, kinds=(0, 1, 2)):
        b = Builder(rng)
        for node in nodes:
            if pos:
                b.track(node, ob.BIND_POSITION, ob.TV_VECTOR3, kinds=kinds)
            if rot:
                b.track(node, ob.BIND_ROTATION, ob.TV_QUAT, lo=-1.0, hi=1.0, kinds=kinds)
            if scale:
                b.track(node, ob.BIND_SCALE, ob.TV_VECTOR3, lo=0.5, hi=2.0, kinds=kinds)
        return b.arrays()

    specs = [
        (build(range(1, 101)), dict(speed=1.0, looped=True, time_slice=(0.0, 2.0), time_position=0.0)),                       # 0: A
        (build(range(50, 151), scale=True), dict(speed=0.5, looped=True, time_slice=(0.0, 2.25), time_position=1.0)),         # 1: B
        (build(range(1, 151), pos=False), dict(speed=-1.0, looped=False, time_slice=(0.25, 2.0), time_position=2.0)),         # 2: C
        (build(range(200, 251)), dict(speed=1.0, looped=True, time_slice=(0.0, 2.0), time_position=0.5)),                     # 3: D
        (build(range(200, 251), kinds=(1,)), dict(speed=2.0, looped=True, time_slice=(0.0, 1.5), time_position=0.0)),         # 4: E
        (build(range(300, 351), scale=True), dict(speed=1.0, looped=True, time_slice=(0.0, 2.0), time_position=0.25)),        # 5: F (direct)
    ]
    anims = []
    for (t, k), kw in specs:
        anims.append(ob.Animation(t, k, **kw))
        ctx.anim_add(t, k, **kw)
    w1, w2 = [0.9, 0.3, 0.6], [1.0, 0.45]
    g1 = ctx.anim_blend_group([0, 1, 2], w1)
    g2 = ctx.anim_blend_group([4, 3], w2)  # source order is the group's, not the animation ids'
    assert (g1, g2) == (1, 2)
    with pytest.raises(fb.FyxError):
        ctx.anim_blend_group([0, 5], [1, 1])  # already in a group
    dt = 0.125
    for frame in range(11):
        if frame == 4:
            ob.lib().orc_animation_set_enabled(anims[1].h, 0)  # B stops ticking; its last pose keeps blending in
            ctx.anim_set_enabled(1, False)
        if frame == 6:
            w1 = [0.0, 0.75, 0.1]
            ctx.anim_set_blend_weights(g1, w1)
        if frame == 8:
            ob.lib().orc_animation_set_enabled(anims[1].h, 1)
            ctx.anim_set_enabled(1, True)
            ob.lib().orc_animation_set_enabled(anims[0].h, 0)  # now the FIRST source is the stale one
            ctx.anim_set_enabled(0, False)
        ob.update_animations([anims[5]], dt, og, tr)
        ob.blend_group_update([anims[0], anims[1], anims[2]], w1, dt, og, tr)
        ob.blend_group_update([anims[4], anims[3]], w2, dt, og, tr)
        og.update()
        ctx.animate(dt)
        ctx.update_transforms(fb.UPDATE_INCREMENTAL)
        times_equal(ctx, anims)
        assert_same_hierarchy(og, ctx)


def test_euler_rotation_tracks_within_libm_tolerance(ctx):
    rng = np.random.default_rng(7)
    n = 300
    parent, flags, trs = make_graph(rng, n, depth1=True)
    trs[:, 7:10] = 1.0  # unit scale: the rotation block of G is the rotation matrix itself
    og, tr = load_pair(ctx, parent, flags, trs)
    b = Builder(rng)
    for node in range(1, n):
        b.track(node, ob.BIND_ROTATION, ob.TV_QUAT_EULER, lo=-6.5, hi=6.5, kinds=(1, 2))  # Track::new_rotation's default kind
    t, k = b.arrays()
    kw = dict(speed=1.0, looped=True, time_slice=(0.0, 2.25), time_position=0.1, enabled=True)
    anim = ob.Animation(t, k, **kw)
    ctx.anim_add(t, k, **kw)
    for frame in range(6):
        ob.update_animations([anim], 0.21, og, tr)
        og.update()
        ctx.animate(0.21)
        ctx.update_transforms(fb.UPDATE_INCREMENTAL)
        times_equal(ctx, [anim])
        G, Go = ctx.get_global_matrices(), og.global_transforms()
        rot_cols = [0, 1, 2, 4, 5, 6, 8, 9, 10]
        assert np.abs(G[:, rot_cols] - Go[:, rot_cols]).max() <= 2e-6
        assert bits_equal(G[:, 12:16], Go[:, 12:16]).all()  # translations do not depend on sin/cos


def test_animation_api_errors(ctx):
    parent, flags, trs = make_graph(np.random.default_rng(1), 10)
    ctx.set_topology(parent, flags)
    ctx.animate(0.1)  # no animations: a no-op
    b = Builder(np.random.default_rng(2))
    b.track(1, ob.BIND_POSITION, ob.TV_VECTOR3)
    t, k = b.arrays()
    bad = t.copy()
    bad["first_key"][0][0] = len(k) + 5
    with pytest.raises(fb.FyxError):
        ctx.anim_add(bad, k)
    bad = t.copy()
    bad["binding"] = 3  # ValueBinding::Property
    with pytest.raises(fb.FyxError):
        ctx.anim_add(bad, k)
    with pytest.raises(fb.FyxError):
        ctx.anim_add(t, k, time_slice=(2.0, 1.0))
    assert ctx.anim_add(t, k) == 0
    with pytest.raises(fb.FyxError):
        ctx.anim_set_enabled(3, True)
    with pytest.raises(fb.FyxError):
        ctx.anim_set_track_enabled(0, 9, True)
    ctx.anim_clear()
    assert ctx.anim_add(t, k, time_slice=(0.0, 1.0)) == 0
    ctx.animate(0.25)
    ctx.update_transforms(fb.UPDATE_INCREMENTAL)
    assert ctx.anim_time_positions(0, 1)[0] == np.float32(0.25)
