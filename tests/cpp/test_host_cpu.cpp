// CPU-only test of the C++ host mirror of the reference's animation data model and LOD groups
// (fyrox_b200/host/fyrox_anim_host.hpp): the reference's own curve tests (fyrox-math/src/curve.rs:410-427, 487-520,
// 570-579 — K14 of tests/golden/reference_kats.json) and the host logic the C ABI expects from its caller.
// No CUDA call is made; run by tests/test_host_mirror_cpu.py.
#include <cstdio>
#include <cstring>

#include "../../fyrox_b200/host/fyrox_anim_host.hpp"

using namespace fyrox;

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                  \
        }                                                                \
    } while (0)

static void test_curve_key_insertion_order() // curve.rs:410-427
{
    Curve curve;
    for (float loc : {0.0f, -1.0f, 3.0f, 2.0f, -5.0f}) curve.add_key(CurveKey(loc, 0.0f, CurveKeyKind::Constant()));
    const float want[5] = {-5.0f, -1.0f, 0.0f, 2.0f, 3.0f};
    for (int i = 0; i < 5; ++i) CHECK(curve.keys()[i].location == want[i]);
}

static void test_curve_from_vec() // curve.rs:570-579
{
    CurveKey key(-1.0f, -1.0f), key2(0.0f, 0.0f), key3(1.0f, 1.0f), key4 = key2;
    Curve curve(std::vector<CurveKey>{key2, key3, key, key4});
    CHECK(curve.name.empty());
    CHECK((curve.keys() == std::vector<CurveKey>{key, key2, key4, key3}));
}

static void test_curve_keys_and_move() // curve.rs:487-520
{
    Curve curve;
    CurveKey key(0.0f, 5.0f, CurveKeyKind::Constant()), key2(1.0f, 10.0f, CurveKeyKind::Linear());
    curve.add_key(key);
    curve.add_key(key2);
    CHECK((curve.keys() == std::vector<CurveKey>{key, key2}));
    CHECK(curve.max_location() == 1.0f);
    CHECK(!curve.is_empty());
    Curve curve2 = curve;
    CurveKey key3; // CurveKey::default()
    curve2.add_key(key3);
    CHECK((curve2.keys() == std::vector<CurveKey>{key3, key, key2}));
    curve2.move_key(0, 20.0f);
    CHECK((curve2.keys() == std::vector<CurveKey>{key, key2, CurveKey(20.0f, 0.0f)}));
    curve.clear();
    CHECK(curve.is_empty());
    CHECK((CurveKeyKind() == CurveKeyKind::Constant()));                          // curve.rs:523-533
    CHECK((CurveKeyKind::new_cubic(0.0f, 0.0f) == CurveKeyKind::Cubic(0.0f, 0.0f)));
}

static void test_flatten()
{
    CHECK(sizeof(fyx_curve_key) == 20 && sizeof(fyx_anim_track) == 52);
    Animation a;
    Track pos = Track::new_position();
    const float ends[3] = {2.0f, 4.0f, 8.0f};
    for (int axis = 0; axis < 3; ++axis) {
        pos.frames.curves[axis].add_key(CurveKey(2.0f, ends[axis], CurveKeyKind::Linear())); // out of order on purpose
        pos.frames.curves[axis].add_key(CurveKey(0.0f, 0.0f, CurveKeyKind::Linear()));
    }
    a.add_track_with_binding(TrackBinding{1u, true}, pos);
    a.tracks.push_back(Track::new_scale()); // in the tracks data, no binding: skipped like update_pose does
    Track off = Track::new_scale();
    off.frames.curves[0].add_key(CurveKey(0.0f, 9.0f));
    a.add_track_with_binding(TrackBinding{1u, false}, off);
    Track rot = Track::new_rotation();
    CHECK(rot.frames.kind == TrackValueKind::UnitQuaternionEuler && rot.frames.curves.size() == 3);
    a.fit_length_to_content();
    CHECK(a.time_slice_start == 0.0f && a.time_slice_end == 2.0f);
    a.time_position = 0.5f;
    std::vector<fyx_anim_track> t;
    std::vector<fyx_curve_key> k;
    fyx_animation_desc d;
    a.flatten(t, k, d);
    CHECK(t.size() == 2 && k.size() == 7 && d.n_tracks == 2 && d.n_keys == 7 && d.tracks == t.data() && d.keys == k.data());
    CHECK(t[0].target_node == 1 && t[0].binding == FYX_BIND_POSITION && t[0].value_kind == FYX_TV_VECTOR3 && t[0].enabled == 1 && t[0].n_curves == 3);
    CHECK(t[0].first_key[1] == 2 && t[0].n_keys[1] == 2 && k[2].location == 0.0f && k[3].location == 2.0f && k[3].value == 4.0f && k[3].kind == FYX_KEY_LINEAR);
    CHECK(t[1].enabled == 0 && t[1].binding == FYX_BIND_SCALE && t[1].n_keys[0] == 1 && t[1].n_keys[1] == 0);
    CHECK(d.speed == 1.0f && d.looped == 1 && d.enabled == 1 && d.time_position == 0.5f && d.time_slice_end == 2.0f && d.struct_size == sizeof d);
}

static void test_lod()
{
    LevelOfDetail l(0.8f, 0.2f, {3}); // LevelOfDetail::new: begin = min(begin, end), end = max(end, begin), clamped
    CHECK(l.begin() == 0.2f && l.end() == 0.2f);
    LevelOfDetail m(-1.0f, 7.0f, {3});
    CHECK(m.begin() == 0.0f && m.end() == 1.0f);
    LevelOfDetail n(0.25f, 0.5f, {1});
    n.set_end(0.1f);
    CHECK(n.begin() == 0.1f && n.end() == 0.25f);
    // owners 9 and 4 both list object 2; owner 9 is visited later (pool order), so its level's range stands; object 7 is dead
    std::map<uint32_t, LodGroup> groups;
    groups[9].levels = {LevelOfDetail(0.5f, 1.0f, {2, 7})};
    groups[4].levels = {LevelOfDetail(0.0f, 0.3f, {2, 5}), LevelOfDetail(0.3f, 1.0f, {6, 5})};
    std::vector<uint32_t> idx;
    std::vector<float> be;
    resolve_lod_ranges(groups, [](uint32_t i) { return i != 7; }, idx, be);
    CHECK((idx == std::vector<uint32_t>{2, 5, 6}));
    CHECK((be == std::vector<float>{0.5f, 1.0f, 0.3f, 1.0f, 0.3f, 1.0f}));
}

int main()
{
    test_curve_key_insertion_order();
    test_curve_from_vec();
    test_curve_keys_and_move();
    test_flatten();
    test_lod();
    if (failures) {
        std::printf("%d check(s) failed\n", failures);
        return 1;
    }
    std::printf("host mirror (animation, LOD): all checks passed\n");
    return 0;
}
