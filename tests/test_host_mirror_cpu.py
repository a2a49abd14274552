"""Host mirrors that need no GPU: fyrox_b200.animation (curves / tracks / animations and their flattening for
fyx_anim_add) and fyrox_b200.lod (LOD groups and the last-write-wins resolution for fyx_set_lod_ranges), pinned by the
reference's own curve tests (K14) and checked against the oracle."""
import ctypes as C
import json
import os

import numpy as np

import fyrox_b200 as fb
import oracle_binding as ob
from fyrox_b200 import _lib as L
from fyrox_b200.animation import (KEY_DTYPE, TRACK_DTYPE, Animation, AnimationContainer, Curve, CurveKey, CurveKeyKind, Track, TrackBinding,
                                  TrackDataContainer, TrackValueKind, ValueBinding)
from fyrox_b200.lod import LevelOfDetail, LodGroup, resolve_lod_ranges
from helpers import random_graph

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def test_k14_curve_keys_stay_sorted_like_the_reference():
    k = KATS["K14_curve_key_order"]
    c = Curve()
    for loc in k["insert"]:
        c.add_key(CurveKey(loc, 0.0, CurveKeyKind.constant()))
    assert [x.location for x in c.keys] == k["expected"]
    c = Curve([CurveKey(l, v) for l, v in k["from_vec"]])  # stable sort: the two equal keys keep their order
    assert [[x.location, x.value] for x in c.keys] == k["from_vec_expected"]
    c = Curve()
    c.add_key(CurveKey(0.0, 5.0, CurveKeyKind.constant()))
    c.add_key(CurveKey(1.0, 10.0, CurveKeyKind.linear()))
    assert c.max_location() == k["max_location"] and c.keys_values() == [5.0, 10.0] and not c.is_empty()
    c.add_key(CurveKey())  # CurveKey::default(): location 0 goes in FRONT of the existing key at 0 (partition_point on <)
    assert [[x.location, x.value] for x in c.keys] == k["after_add_default"]
    c.move_key(0, 20.0)
    assert [[x.location, x.value] for x in c.keys] == k["after_move_key0_to_20"]
    assert CurveKeyKind() == CurveKeyKind.constant() and CurveKeyKind.new_cubic(0.0, 0.0) == CurveKeyKind.cubic(0.0, 0.0)
    c.clear()
    assert c.is_empty() and c.max_location() == 0.0


def test_layouts_match_the_c_abi():
    assert KEY_DTYPE.itemsize == C.sizeof(L.fyx_curve_key) == 20 and TRACK_DTYPE.itemsize == C.sizeof(L.fyx_anim_track) == 52
    assert KEY_DTYPE == fb.Context.KEY_DTYPE and TRACK_DTYPE == fb.Context.TRACK_DTYPE
    assert KEY_DTYPE == ob.KEY_DTYPE and TRACK_DTYPE == ob.TRACK_DTYPE
    assert [TrackValueKind.components_count(k) for k in range(6)] == [1, 2, 3, 4, 3, 4]
    assert Track.new_rotation().frames.kind == TrackValueKind.UnitQuaternionEuler and len(Track.new_rotation().frames.curves) == 3


def test_flattened_animation_drives_the_oracle():
    """An Animation built with the reference's API, flattened for fyx_anim_add, is what the oracle's Animation takes: a
    linear position track (0,0,0)@0 -> (2,4,8)@2 on node 1 plus a track without binding (skipped) and a disabled one."""
    a = Animation("walk")
    pos = Track.new_position()
    for axis, end in enumerate((2.0, 4.0, 8.0)):
        pos.frames.curves[axis].add_key(CurveKey(2.0, end, CurveKeyKind.linear()))  # out of order on purpose
        pos.frames.curves[axis].add_key(CurveKey(0.0, 0.0, CurveKeyKind.linear()))
    a.add_track_with_binding(TrackBinding(1), pos)
    orphan = Track.new_scale()
    a.tracks.append(orphan)  # in the tracks data but without a binding: update_pose skips it
    off = Track.new_scale()
    off.frames.curves[0].add_key(CurveKey(0.0, 9.0))
    a.add_track_with_binding(TrackBinding(1, enabled=False), off)
    a.fit_length_to_content()
    assert a.time_slice == (0.0, 2.0)
    a.time_position = 0.5
    tracks, keys, kw = a.flatten()
    assert len(tracks) == 2 and tracks[0]["binding"] == ValueBinding.Position and tracks[1]["enabled"] == 0 and len(keys) == 7
    oa = ob.Animation(tracks, keys, **kw)
    parent = np.array([0xFFFFFFFF, 0], np.uint32)
    og = ob.Graph.build(parent, None, None, np.tile(np.eye(4, dtype=np.float32).reshape(16), (2, 1)), None)
    tr = (ob.Transform * 2)()
    for i in range(2):
        ob.lib().orc_transform_identity(C.byref(tr[i]))
    ob.update_animations([oa], 0.5, og, tr)
    assert tuple(tr[1].local_position) == (0.5, 1.0, 2.0) and tuple(tr[1].local_scale) == (1.0, 1.0, 1.0)
    assert oa.time_position == 1.0

    class FakeCtx:  # AnimationContainer.upload hands the animations over in pool order
        def __init__(self):
            self.calls = []

        def anim_add(self, t, k, **kw):
            self.calls.append((len(t), len(k), kw["speed"]))
            return len(self.calls) - 1

    cont = AnimationContainer()
    b = Animation("idle")
    b.set_speed(2.0)
    assert cont.add(a) == 0 and cont.add(b) == 1
    fc = FakeCtx()
    assert cont.upload(fc) == [0, 1] and fc.calls == [(2, 7, 1.0), (0, 0, 2.0)]


def test_level_of_detail_constructor_and_setters():
    l = LevelOfDetail(0.8, 0.2, [3])  # LevelOfDetail::new: begin = min(begin, end), end = max(end, begin), both clamped
    assert (l.begin(), l.end()) == (float(np.float32(0.2)), float(np.float32(0.2)))
    l = LevelOfDetail(-1.0, 7.0, [3])
    assert (l.begin(), l.end()) == (0.0, 1.0)
    l.set_begin(2.0)  # clamps to 1; not > end (1.0): no swap
    assert (l.begin(), l.end()) == (1.0, 1.0)
    l = LevelOfDetail(0.25, 0.5, [1])
    l.set_end(0.1)  # end < begin: swapped
    assert (l.begin(), l.end()) == (float(np.float32(0.1)), 0.25)


def test_resolved_lod_ranges_reproduce_the_reference_loop():
    """resolve_lod_ranges (owners in pool order, levels, objects, last write wins, dead objects skipped) gives, per object, a
    range whose verdict equals the lod_filter the oracle computes with the literal loop of from_graph."""
    rng = np.random.default_rng(77)
    parent, flags, mask, local, aabb = random_graph(rng, 1500, p_orphan=0.02)
    n = len(parent)
    og = ob.Graph.build(parent, flags, mask, local, aabb)
    og.update_hierarchical_data()
    alive = np.nonzero((flags & fb.NODE_ALIVE) != 0)[0]
    groups = {}
    for o in rng.choice(alive, 25, replace=False):
        lv = [LevelOfDetail(b, e, rng.integers(1, n, rng.integers(1, 6)).tolist()) for (b, e) in ((0.0, 0.3), (0.3, 0.55), (0.55, 1.0))]
        groups[int(o)] = LodGroup(lv)
        og.set_lod_group(int(o), [(l.begin(), l.end(), l.objects) for l in lv])
    idx, ranges = resolve_lod_ranges(groups, is_alive=lambda i: bool(flags[i] & fb.NODE_ALIVE))
    assert idx.size and np.all(np.diff(idx.astype(np.int64)) > 0) and ranges.shape == (idx.size, 2)
    G = og.global_transforms()
    for eye, zn, zf in (((0, 0, 60), 0.1, 200.0), ((35, 5, -10), 0.5, 90.0)):
        want = og.lod_filter(eye, zn, zf)
        got = np.ones(n, bool)
        e = np.array(eye, np.float32)
        for k, x in enumerate(idx):
            d = e - G[x, 12:15]
            dist = np.sqrt(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]), dtype=np.float32)
            nrm = np.float32(np.float32(dist - np.float32(zn)) / np.float32(np.float32(zf) - np.float32(zn)))
            got[x] = (nrm >= ranges[k, 0]) and (nrm <= ranges[k, 1])
        assert np.array_equal(got, want)


def test_cpp_host_mirror_of_animation_and_lod():
    """tests/cpp/test_host_cpu.cpp: the same reference curve tests (K14) and host logic on the C++ mirror
    (fyrox_b200/host/fyrox_anim_host.hpp); pure host code, no CUDA call."""
    import subprocess

    d = os.path.join(HERE, "cpp")
    subprocess.run(["make", "-C", d, "test_host_cpu"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(d, "test_host_cpu")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
