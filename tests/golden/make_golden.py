#!/usr/bin/env python
"""Emit tests/golden/reference_kats.json: the known-answer vectors of the reference's own unit tests
for the render-prep path (SURVEY.md §8c K1–K10, K16) and of the animation-sampling step before it (K11–K15), transcribed from the cited test sources.

The reference is Rust (no toolchain here), so the vectors cannot be produced by running it.  When
/root/reference exists (the build container) this script also checks that each cited file still
contains the quoted fragment, so a transcription cannot drift from the source silently.
"""
import json
import os

REF = "/root/reference"
FMAX = 3.4028234663852886e38
INV_SQRT3 = 0.57735026  # literal used by the reference test

I4 = [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]

KATS = {
    "K1_aabb_transform": {
        "source": "fyrox-math/src/aabb.rs:357-371",
        "quote": "assert_eq!(transformed_aabb.max, Vector3::new(3.0, 3.0, 3.0));",
        "aabb": [0.0, 0.0, 0.0, 1.0, 1.0, 1.0],
        "translation": [1.0, 1.0, 1.0],
        "scaling": [2.0, 2.0, 2.0],
        "expected": [1.0, 1.0, 1.0, 3.0, 3.0, 3.0],
    },
    "K2_aabb_basics": {
        "source": "fyrox-math/src/aabb.rs:373-523",
        "quote": "assert_eq!(_box.max, Vector3::new(-f32::MAX, -f32::MAX, -f32::MAX));",
        "default": [FMAX, FMAX, FMAX, -FMAX, -FMAX, -FMAX],
        "unit": [-0.5, -0.5, -0.5, 0.5, 0.5, 0.5],
        "add_point": {"points": [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], "expected": [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]},
        "add_box": {"start": [0.0] * 6, "box": [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], "expected": [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]},
        "corners_of_radius_1": [
            [-1.0, -1.0, -1.0], [-1.0, -1.0, 1.0], [1.0, -1.0, 1.0], [1.0, -1.0, -1.0],
            [-1.0, 1.0, -1.0], [-1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, -1.0],
        ],
        "is_valid": {"default": False, "after_add_1_1_1": True, "after_add_m1": True},
        "is_degenerate": {"unit": False, "collapsed": True},
    },
    "K3_frustum_from_identity": {
        "source": "fyrox-math/src/frustum.rs:370-453",
        "quote": "Plane::from_abcd(0.0, 0.0, -1.0, 1.0).unwrap(),",
        "vp_row_major": I4,
        "planes_abcd": [[1.0, 0.0, 0.0, 1.0], [-1.0, 0.0, 0.0, 1.0], [0.0, -1.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [0.0, 0.0, -1.0, 1.0], [0.0, 0.0, 1.0, 1.0]],
        "corners": [[-1.0, 1.0, 1.0], [-1.0, -1.0, 1.0], [1.0, -1.0, 1.0], [1.0, 1.0, 1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, -1.0], [1.0, -1.0, -1.0], [1.0, 1.0, -1.0]],
    },
    "K4_frustum_queries": {
        "source": "fyrox-math/src/frustum.rs:471-537",
        "quote": "assert!(!f.is_intersects_point_cloud(&[Vector3::new(-1.0, -2.0, 1.0)]));",
        "cloud_true": [[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]],
        "cloud_false": [[-1.0, -2.0, 1.0]],
        "aabb_true": [-0.5, -0.5, -0.5, 0.5, 0.5, 0.5],
        "aabb_false": [5.0, 5.0, 5.0, 15.0, 15.0, 15.0],
        "offset_true": [1.0, 1.0, 1.0],
        "offset_false": [10.0, 10.0, 10.0],
        "contains_true": [0.0, 0.0, 0.0],
        "contains_false": [10.0, 10.0, 10.0],
    },
    "K5_plane": {
        "source": "fyrox-math/src/plane.rs:131-205",
        "quote": "normal: Vector3::new(0.57735026, 0.57735026, 0.57735026),",
        "from_abcd_1110": {"normal": [INV_SQRT3] * 3, "d": 0.0},
        "from_abcd_0000_is_none": True,
        "dot": {"normal": [0.0, 0.0, 1.0], "d": 0.0, "point": [1.0, 1.0, 1.0], "expected": 1.0},
        "intersection_of_axis_planes": [0.0, 0.0, 0.0],
    },
    "K6_hierarchy_propagation": {
        "source": "fyrox-impl/src/scene/graph/mod.rs:2646-2739",
        "quote": "assert_eq!(graph[d].global_position(), Vector3::new(2.0, 1.0, 1.0));",
        "local_positions": {"a": [1.0, 0.0, 0.0], "b": [0.0, 1.0, 0.0], "c": [0.0, 0.0, 1.0], "d": [1.0, 1.0, 1.0]},
        "b_visibility": False,
        "b_enabled": False,
        "first": {
            "global_positions": {"a": [1.0, 0.0, 0.0], "b": [1.0, 1.0, 0.0], "c": [1.0, 1.0, 1.0], "d": [2.0, 1.0, 1.0]},
            "global_visibility": {"a": True, "b": False, "c": False, "d": True},
            "global_enabled": {"a": True, "b": False, "c": False, "d": True},
        },
        "second": {
            "global_positions": {"a": [1.0, 0.0, 0.0], "b": [1.0, 2.0, 0.0], "c": [1.0, 2.0, 1.0], "d": [2.0, 1.0, 1.0]},
            "global_visibility": {"a": True, "b": True, "c": True, "d": True},
            "global_enabled": {"a": False, "b": False, "c": False, "d": False},
        },
    },
    "K7_global_scale": {
        "source": "fyrox-impl/src/scene/graph/mod.rs:2602-2644",
        "quote": "assert_eq!(graph.global_scale(c), Vector3::new(3.0, 4.0, 6.0));",
        "local_scales": {"a": [1.0, 1.0, 2.0], "b": [3.0, 2.0, 1.0], "c": [1.0, 2.0, 3.0]},
        "expected": {"a": [1.0, 1.0, 2.0], "b": [3.0, 2.0, 2.0], "c": [3.0, 4.0, 6.0]},
    },
    "K8_matrix4_ext": {
        "source": "fyrox-math/src/lib.rs:1480-1501",
        "quote": "fn matrix4_ext_for_matrix4()",
        "side": [1.0, 0.0, 0.0], "up": [0.0, 1.0, 0.0], "look": [0.0, 0.0, 1.0], "position": [0.0, 0.0, 0.0],
        "linear_indices": {"side": [0, 1, 2], "up": [4, 5, 6], "look": [8, 9, 10], "position": [12, 13, 14]},
    },
    "K9_sorting_index": {
        "source": "fyrox-impl/src/renderer/bundle.rs:1288-1331",
        "quote": "fn test_calculate_sorting_index()",
        "range_center": (2**64 - 1) // 2,
        "cases": [{"z": 0.0, "delta": 0}, {"z": 1.0, "delta": 1000}, {"z": 2.0, "delta": 2000}, {"z": -3.0, "delta": -3000}],
    },
    "K11_wrapf": {
        "source": "fyrox-math/src/lib.rs:1141-1147",
        "quote": "assert_eq!(wrapf(12.0, 5.0, 10.0), 7.0);",
        "cases": [[5.0, 0.0, 10.0, 5.0], [5.0, 0.0, 0.0, 0.0], [2.0, 5.0, 10.0, 7.0], [12.0, 5.0, 10.0, 7.0]],
    },
    "K12_curve_value_at": {
        "source": "fyrox-math/src/curve.rs:428-471",
        "quote": "assert_eq!(curve.value_at(0.5, &mut 0), 0.5);",
        # keys are (location, value, kind) with kind Linear; every query starts from hint 0
        "steps": [
            {"keys": [], "queries": [[0.0, 0.0]]},
            {"keys": [[0.0, 1.0]], "queries": [[-1.0, 1.0], [1.0, 1.0], [0.0, 1.0]]},
            {"keys": [[0.0, 1.0], [1.0, 0.0]], "queries": [[-1.0, 1.0], [2.0, 0.0], [0.5, 0.5]]},
            {"keys": [[0.0, 1.0], [1.0, 0.0], [2.0, 1.0]], "queries": [[-1.0, 1.0], [3.0, 1.0], [0.0, 1.0], [2.0, 1.0], [0.5, 0.5]]},
        ],
    },
    "K13_curve_key_interpolate": {
        "source": "fyrox-math/src/curve.rs:536-575",
        "quote": "assert_eq!(key3.interpolate(&key4, 1.0), 30.0);",
        # kinds: 0 Constant, 1 Linear, 2 Cubic (new_cubic(0.0, 0.0): both tangents tan(0) = 0)
        "keys": {"key": [0.0, 5.0, 0], "key2": [1.0, 10.0, 1], "key3": [2.0, 20.0, 2], "key4": [3.0, 30.0, 2]},
        "cases": [
            ["key", "key2", 1.0, 10.0], ["key", "key2", 0.0, 5.0], ["key", "key3", 1.0, 20.0], ["key", "key3", 0.0, 5.0],
            ["key2", "key", 1.0, 5.0], ["key2", "key", 0.0, 10.0], ["key2", "key3", 1.0, 20.0], ["key2", "key3", 0.0, 10.0],
            ["key3", "key", 1.0, 5.0], ["key3", "key", 0.0, 20.0], ["key3", "key2", 1.0, 10.0], ["key3", "key2", 0.0, 20.0],
            ["key3", "key4", 1.0, 30.0], ["key3", "key4", 0.0, 20.0],
        ],
    },
    "K14_curve_key_order": {
        "source": "fyrox-math/src/curve.rs:410-427",
        "quote": "assert_eq!(curve.keys[0].location, -5.0);",
        "insert": [0.0, -1.0, 3.0, 2.0, -5.0],
        "expected": [-5.0, -1.0, 0.0, 2.0, 3.0],
        # test_curve_from_vec (curve.rs:570-579): keys (location, value) given as [key2, key3, key, key4 = key2.clone()]
        "from_vec": [[0.0, 0.0], [1.0, 1.0], [-1.0, -1.0], [0.0, 0.0]],
        "from_vec_expected": [[-1.0, -1.0], [0.0, 0.0], [0.0, 0.0], [1.0, 1.0]],
        # test_curve (curve.rs:487-520): keys 5@0 (Constant), 10@1 (Linear); a default key (0@0) added lands in front; moved to 20 it ends last
        "max_location": 1.0,
        "after_add_default": [[0.0, 0.0], [0.0, 5.0], [1.0, 10.0]],
        "after_move_key0_to_20": [[0.0, 5.0], [1.0, 10.0], [20.0, 0.0]],
    },
    "K15_quat_from_euler": {
        # A relation, not a literal: the reference asserts that ITS product qz*qy*qx equals nalgebra's closed form, with
        # UnitQuaternion's PartialEq (all coordinates equal, or all negated).  The oracle restates both sides — quat_mul's
        # operation order and from_euler_angles (w = cr*cp*cy + sr*sp*sy, i = sr*cp*cy - cr*sp*sy, j = cr*sp*cy + sr*cp*sy,
        # k = cr*cp*sy - sr*sp*cy with (s, c) = sin_cos(angle * 0.5), nalgebra 0.3x geometry/quaternion_construction.rs, recalled) —
        # and must reproduce the equality BIT FOR BIT; with these angles the coordinates are 1 and +-4.37e-8 whose last ulps
        # depend on the order of the sums.
        "source": "fyrox-math/src/lib.rs:1460-1477",
        "quote": "UnitQuaternion::from_euler_angles(",
        "euler": ["pi", "pi", "pi"],
        "order": "XYZ",
    },
    "K16_vertex_buffer_attributes": {
        # the reference's own interleaved test vertex (repr(C): position 3f, tex 2f, tex 2f, normal 3f, tangent 4f, bone weights 4f,
        # bone indices 4 x u8 = 76 bytes) and what its attribute views read back (test_view_original_equal / test_attribute_view).
        # Offsets follow from the declared attribute sizes in order (buffer.rs: offset = running sum of size * sizeof(type)).
        "source": "fyrox-impl/src/scene/mesh/buffer.rs:1687-1828,1882-1901",
        "quote": "bone_indices: Vector4::new(1, 2, 3, 4),",
        "stride": 76,
        "offsets": {"position": 0, "tex_coord": 12, "second_tex_coord": 20, "normal": 28, "tangent": 40, "bone_weights": 56, "bone_indices": 72},
        "vertices": [
            {"position": [1.0, 2.0, 3.0], "tex_coord": [0.0, 1.0], "second_tex_coord": [1.0, 0.0], "normal": [0.0, 1.0, 0.0],
             "tangent": [1.0, 0.0, 0.0, 1.0], "bone_weights": [0.25, 0.25, 0.25, 0.25], "bone_indices": [1, 2, 3, 4]},
            {"position": [3.0, 2.0, 1.0], "tex_coord": [1.0, 0.0], "second_tex_coord": [1.0, 0.0], "normal": [0.0, 1.0, 0.0],
             "tangent": [1.0, 0.0, 0.0, 1.0], "bone_weights": [0.25, 0.25, 0.25, 0.25], "bone_indices": [1, 2, 3, 4]},
            {"position": [1.0, 1.0, 1.0], "tex_coord": [1.0, 1.0], "second_tex_coord": [1.0, 0.0], "normal": [0.0, 1.0, 0.0],
             "tangent": [1.0, 0.0, 0.0, 1.0], "bone_weights": [0.25, 0.25, 0.25, 0.25], "bone_indices": [1, 2, 3, 4]},
        ],
    },
    "K10_handle_numbering": {
        "source": "fyrox-impl/src/scene/graph/mod.rs:408-424",
        "quote": "root: Handle::NONE,",
        "root": [0, 1],
        "first_added": [1, 1],
    },
}


def verify_against_reference():
    missing = []
    for name, k in KATS.items():
        path = os.path.join(REF, k["source"].rsplit(":", 1)[0])
        if not os.path.exists(path):
            missing.append((name, "file missing: " + path))
            continue
        if k["quote"] not in open(path, encoding="utf-8").read():
            missing.append((name, "quote not found in " + path))
    return missing


if __name__ == "__main__":
    if os.path.isdir(REF):
        bad = verify_against_reference()
        if bad:
            raise SystemExit("golden transcription drifted from the reference: %r" % bad)
        print("all %d quoted fragments found in %s" % (len(KATS), REF))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(KATS, f, indent=1, sort_keys=True)
    print("wrote", out)
