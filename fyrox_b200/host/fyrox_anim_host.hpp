// fyrox_anim_host.hpp — C++ host mirror of the reference's animation data model and LOD groups (header-only).
//
// The objects a Fyrox host owns, with the reference's names and argument meaning, and the two pieces of host logic
// the C ABI expects from its caller:
//   * Animation::flatten  — AnimationTracksData + TrackBindings -> fyx_anim_track[] / fyx_curve_key[] for fyx_anim_add
//                           
//   * resolve_lod_ranges  — the write order of from_graph's lod_filter loop (renderer/bundle.rs:898-916) resolved to one
//                           range per object for fyx_set_lod_ranges
// Nothing here samples a curve or tests a distance: that is the device's job.  Reference: fyrox-math/src/curve.rs,
// fyrox-animation/src/{track,container,lib}.rs, fyrox-impl/src/scene/base.rs:61-160.
// Python twin: fyrox_b200/animation.py, fyrox_b200/lod.py.  Pinned without a GPU by tests/cpp/test_host_cpu.cpp.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/fyrox_b200.h"

namespace fyrox {

// fyrox-math/src/curve.rs:33-55
struct CurveKeyKind {
    uint32_t kind = FYX_KEY_CONSTANT;
    float left_tangent = 0.0f, right_tangent = 0.0f; // tan(angle), Cubic only
    static CurveKeyKind Constant() { return {}; }
    static CurveKeyKind Linear() { return {FYX_KEY_LINEAR, 0.0f, 0.0f}; }
    static CurveKeyKind Cubic(float left_tangent, float right_tangent) { return {FYX_KEY_CUBIC, left_tangent, right_tangent}; }
    static CurveKeyKind new_cubic(float left_angle_radians, float right_angle_radians) // curve.rs:47-54
    {
        return {FYX_KEY_CUBIC, std::tan(left_angle_radians), std::tan(right_angle_radians)};
    }
    bool operator==(const CurveKeyKind &o) const { return kind == o.kind && left_tangent == o.left_tangent && right_tangent == o.right_tangent; }
};

// curve.rs:57-75
struct CurveKey {
    float location = 0.0f, value = 0.0f;
    CurveKeyKind kind;
    CurveKey() = default;
    CurveKey(float location_, float value_, CurveKeyKind kind_ = {}) : location(location_), value(value_), kind(kind_) {}
    bool operator==(const CurveKey &o) const { return location == o.location && value == o.value && kind == o.kind; }
};

// curve.rs:150-255: keys stay sorted by location; equal locations keep their order (Rust's sort_by is stable, add_key
// inserts in front of the first key that is not smaller)
class Curve {
  public:
    Curve() = default;
    explicit Curve(std::vector<CurveKey> keys) : keys_(std::move(keys)) { sort_keys(); } // From<Vec<CurveKey>>
    void clear() { keys_.clear(); }
    bool is_empty() const { return keys_.empty(); }
    const std::vector<CurveKey> &keys() const { return keys_; }
    void add_key(const CurveKey &k)
    {
        auto pos = std::partition_point(keys_.begin(), keys_.end(), [&](const CurveKey &x) { return x.location < k.location; });
        keys_.insert(pos, k);
    }
    void move_key(size_t key_id, float location)
    {
        if (key_id < keys_.size()) {
            keys_[key_id].location = location;
            sort_keys();
        }
    }
    float max_location() const { return keys_.empty() ? 0.0f : keys_.back().location; }
    std::string name;

  private:
    void sort_keys() { std::stable_sort(keys_.begin(), keys_.end(), [](const CurveKey &a, const CurveKey &b) { return a.location < b.location; }); }
    std::vector<CurveKey> keys_;
};

// fyrox-animation/src/container.rs:41-76
enum class TrackValueKind : uint32_t { Real = FYX_TV_REAL, Vector2 = FYX_TV_VECTOR2, Vector3 = FYX_TV_VECTOR3, Vector4 = FYX_TV_VECTOR4,
                                       UnitQuaternionEuler = FYX_TV_QUAT_EULER, UnitQuaternion = FYX_TV_QUAT };
inline size_t components_count(TrackValueKind k)
{
    switch (k) {
    case TrackValueKind::Real: return 1;
    case TrackValueKind::Vector2: return 2;
    case TrackValueKind::Vector3: return 3;
    case TrackValueKind::Vector4: return 4;
    case TrackValueKind::UnitQuaternionEuler: return 3;
    default: return 4;
    }
}
// fyrox-animation/src/value.rs:358-374 (property bindings are not supported on the device)
enum class ValueBinding : uint32_t { Position = FYX_BIND_POSITION, Scale = FYX_BIND_SCALE, Rotation = FYX_BIND_ROTATION };

// container.rs:99-160, 303-312
struct TrackDataContainer {
    TrackValueKind kind = TrackValueKind::Vector3;
    std::vector<Curve> curves;
    explicit TrackDataContainer(TrackValueKind k = TrackValueKind::Vector3) : kind(k), curves(components_count(k)) {}
    float time_length() const
    {
        float length = 0.0f;
        for (const Curve &c : curves)
            if (c.max_location() > length) length = c.max_location();
        return length;
    }
};

// fyrox-animation/src/track.rs:95-205
struct Track {
    TrackDataContainer frames;
    ValueBinding binding = ValueBinding::Position;
    uint64_t id; // stands for the Uuid
    explicit Track(TrackDataContainer c = TrackDataContainer(), ValueBinding b = ValueBinding::Position) : frames(std::move(c)), binding(b), id(next_id()) {}
    static Track new_position() { return Track(TrackDataContainer(TrackValueKind::Vector3), ValueBinding::Position); }
    static Track new_rotation() { return Track(TrackDataContainer(TrackValueKind::UnitQuaternionEuler), ValueBinding::Rotation); } // track.rs:139-145
    static Track new_scale() { return Track(TrackDataContainer(TrackValueKind::Vector3), ValueBinding::Scale); }
    float time_length() const { return frames.time_length(); }

  private:
    static uint64_t next_id()
    {
        static uint64_t n = 0;
        return ++n;
    }
};

// track.rs:36-93
struct TrackBinding {
    uint32_t target = FYX_NONE; // Handle<Node>::index()
    bool enabled = true;
};

// fyrox-animation/src/lib.rs:269-496, 755-790, 922-945
class Animation {
  public:
    std::string name;
    std::vector<Track> tracks; // AnimationTracksData::tracks
    std::map<uint64_t, TrackBinding> track_bindings;
    float speed = 1.0f, time_position = 0.0f, time_slice_start = 0.0f, time_slice_end = 0.0f; // Animation::default
    bool enabled = true, looped = true;

    void add_track_with_binding(const TrackBinding &binding, Track track)
    {
        track_bindings[track.id] = binding;
        tracks.push_back(std::move(track));
    }
    void set_time_slice(float start, float end)
    {
        assert(start <= end); // lib.rs:446
        time_slice_start = start;
        time_slice_end = end;
    }
    void fit_length_to_content() // lib.rs:410-421
    {
        time_slice_start = 0.0f;
        for (const Track &t : tracks)
            if (t.time_length() > time_slice_end) time_slice_end = t.time_length();
    }
    // tracks in AnimationTracksData order; a track without a binding is left out like update_pose skips it (lib.rs:903-905).
    // `desc` points into the two vectors: keep them alive until fyx_anim_add has returned.
    void flatten(std::vector<fyx_anim_track> &out_tracks, std::vector<fyx_curve_key> &out_keys, fyx_animation_desc &desc) const
    {
        out_tracks.clear();
        out_keys.clear();
        for (const Track &t : tracks) {
            auto b = track_bindings.find(t.id);
            if (b == track_bindings.end()) continue;
            fyx_anim_track r{};
            r.target_node = b->second.target;
            r.binding = (uint32_t)t.binding;
            r.value_kind = (uint32_t)t.frames.kind;
            r.enabled = b->second.enabled ? 1u : 0u;
            r.n_curves = (uint32_t)std::min<size_t>(t.frames.curves.size(), 4);
            for (uint32_t c = 0; c < r.n_curves; ++c) {
                r.first_key[c] = (uint32_t)out_keys.size();
                r.n_keys[c] = (uint32_t)t.frames.curves[c].keys().size();
                for (const CurveKey &k : t.frames.curves[c].keys())
                    out_keys.push_back(fyx_curve_key{k.location, k.value, k.kind.kind, k.kind.left_tangent, k.kind.right_tangent});
            }
            out_tracks.push_back(r);
        }
        desc = fyx_animation_desc{};
        desc.struct_size = sizeof(fyx_animation_desc);
        desc.n_tracks = (uint32_t)out_tracks.size();
        desc.tracks = out_tracks.data();
        desc.n_keys = (uint32_t)out_keys.size();
        desc.keys = out_keys.data();
        desc.speed = speed;
        desc.time_position = time_position;
        desc.time_slice_start = time_slice_start;
        desc.time_slice_end = time_slice_end;
        desc.looped = looped ? 1u : 0u;
        desc.enabled = enabled ? 1u : 0u;
    }
};

// lib.rs:947-1100 reduced to what the device needs: animations in pool order = the order update_animations walks them
class AnimationContainer {
  public:
    size_t add(Animation a)
    {
        animations.push_back(std::move(a));
        return animations.size() - 1;
    }
    // hands every animation to the context in pool order; returns the device ids (or an empty vector on the first error)
    std::vector<uint32_t> upload(fyx_ctx *ctx) const
    {
        std::vector<uint32_t> ids;
        std::vector<fyx_anim_track> t;
        std::vector<fyx_curve_key> k;
        for (const Animation &a : animations) {
            fyx_animation_desc d;
            a.flatten(t, k, d);
            uint32_t id = 0;
            if (fyx_anim_add(ctx, &d, &id) != FYX_OK) return {};
            ids.push_back(id);
        }
        return ids;
    }
    std::vector<Animation> animations;
};

// ---- LOD groups: fyrox-impl/src/scene/base.rs:61-160 ----
class LevelOfDetail {
  public:
    LevelOfDetail(float begin, float end, std::vector<uint32_t> objects_) : objects(std::move(objects_))
    {
        for (uint32_t o : objects) assert(o != FYX_NONE && "Invalid handles are not allowed"); // base.rs:75-78
        begin = std::min(begin, end); // base.rs:79-80
        end = std::max(end, begin);
        begin_ = clamp01(begin);
        end_ = clamp01(end);
    }
    void set_begin(float percent)
    {
        begin_ = clamp01(percent);
        if (begin_ > end_) std::swap(begin_, end_);
    }
    float begin() const { return begin_; }
    void set_end(float percent)
    {
        end_ = clamp01(percent);
        if (end_ < begin_) std::swap(begin_, end_);
    }
    float end() const { return end_; }
    std::vector<uint32_t> objects;

  private:
    static float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); } // f32::clamp; NaN stays NaN
    float begin_, end_;
};

struct LodGroup {
    std::vector<LevelOfDetail> levels;
};

// The loop of from_graph (renderer/bundle.rs:898-916) writes lod_filter[object] for every object of every level of every
// owner, owners in pool order: the verdict written last stands.  Resolved here to one (begin, end) per object, dead
// objects (try_get_node fails) skipped; out_idx ascending, out_begin_end 2 floats per object — fyx_set_lod_ranges' input.
inline void resolve_lod_ranges(const std::map<uint32_t, LodGroup> &groups_by_owner, const std::function<bool(uint32_t)> &is_alive,
                               std::vector<uint32_t> &out_idx, std::vector<float> &out_begin_end)
{
    std::map<uint32_t, std::pair<float, float>> last;
    for (const auto &g : groups_by_owner) // std::map iterates in key (= pool) order
        for (const LevelOfDetail &l : g.second.levels)
            for (uint32_t obj : l.objects) {
                if (is_alive && !is_alive(obj)) continue;
                last[obj] = {l.begin(), l.end()};
            }
    out_idx.clear();
    out_begin_end.clear();
    for (const auto &e : last) {
        out_idx.push_back(e.first);
        out_begin_end.push_back(e.second.first);
        out_begin_end.push_back(e.second.second);
    }
}

} // namespace fyrox
