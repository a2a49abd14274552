"""The cull kernels do not run the reference's 8 corners x 6 planes loop (fyrox-math/src/frustum.rs:205-219): they test, per
plane, the ONE corner built from the per-axis bound whose product with the normal component is the larger one, and claim
identical booleans because IEEE rounding is monotone (fyx_math.cuh, frustum_intersects_aabb).  This file checks that claim
on the CPU, in numpy float32 (one rounding per operation, like the kernels' __fmul_rn / __fadd_rn), over millions of random
and adversarial planes / boxes — including zeros of both signs, denormals, huge and tiny magnitudes, degenerate boxes — and
the same for the corner-in-box fallback evaluated through per-axis masks of DISTINCT corner coordinates."""
import numpy as np

f32 = np.float32


def s_of(n, d, px, py, pz):
    """Plane::dot + d in the reference's order: ((nx*px + ny*py) + nz*pz) + d, one rounding per op."""
    return ((n[:, 0] * px + n[:, 1] * py) + n[:, 2] * pz) + d


def literal_all_behind(n, d, lo, hi):
    """all 8 corners have s <= 0 (frustum.rs:208-214 for one plane)"""
    allb = np.ones(len(n), bool)
    for c in range(8):
        px = np.where(c & 1, hi[:, 0], lo[:, 0])
        py = np.where(c & 2, hi[:, 1], lo[:, 1])
        pz = np.where(c & 4, hi[:, 2], lo[:, 2])
        allb &= s_of(n, d, px, py, pz) <= 0
    return allb


def trick_all_behind(n, d, lo, hi):
    """the kernels' form: operand picked by the sign of the normal component (n < 0 -> min, else max), one corner"""
    vx = np.where(n[:, 0] < 0, lo[:, 0], hi[:, 0])
    vy = np.where(n[:, 1] < 0, lo[:, 1], hi[:, 1])
    vz = np.where(n[:, 2] < 0, lo[:, 2], hi[:, 2])
    return s_of(n, d, vx, vy, vz) <= 0


def special_values(rng, m):
    pool = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-45, -1e-45, 1e-38, -1e-38, 1e-20, 3.0, -3.0, 1e10, -1e10, 1e17, -1e17, 0.1, -0.1,
                     1.0000001, 0.99999994, 16777216.0, -16777217.0, 1e-7, -1e-7], f32)
    return pool[rng.integers(0, len(pool), m)]


def make_cases(rng, m, mode):
    if mode == "random":
        n = rng.normal(size=(m, 3)).astype(f32)
        d = (rng.normal(size=m) * 10).astype(f32)
        a = (rng.normal(size=(m, 3)) * 20).astype(f32)
        b = (rng.normal(size=(m, 3)) * 20).astype(f32)
    elif mode == "scaled":  # wildly different magnitudes per component: cancellation and absorption everywhere
        e = lambda *s, top=12: (10.0 ** rng.uniform(-12, top, s)).astype(f32)
        # plane normals are unit vectors in the engine; here up to 1e6 so that no product overflows (|n|*|bound| <= 1e24):
        # the claim is about rounding, not about inf - inf
        n = (rng.normal(size=(m, 3)).astype(f32)) * e(m, 3, top=5)
        d = rng.normal(size=m).astype(f32) * e(m)
        a = rng.normal(size=(m, 3)).astype(f32) * e(m, 3)
        b = rng.normal(size=(m, 3)).astype(f32) * e(m, 3)
    else:  # "special": zeros of both signs, denormals, ties
        n = special_values(rng, m * 3).reshape(m, 3)
        d = special_values(rng, m)
        a = special_values(rng, m * 3).reshape(m, 3)
        b = special_values(rng, m * 3).reshape(m, 3)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    # the kernels take this path only for "tame" boxes (finite, |bound| <= 1e18, min <= max); -0/+0 pairs count as ordered
    keep = (np.abs(lo) <= 1e18).all(1) & (np.abs(hi) <= 1e18).all(1)
    return n[keep], d[keep], lo[keep], hi[keep]


def test_max_corner_equals_the_eight_corner_loop():
    rng = np.random.default_rng(2024)
    total = 0
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        for mode, m in (("random", 2_000_000), ("scaled", 2_000_000), ("special", 2_000_000)):
            n, d, lo, hi = make_cases(rng, m, mode)
            a, b = literal_all_behind(n, d, lo, hi), trick_all_behind(n, d, lo, hi)
            assert np.isfinite(s_of(n, d, hi[:, 0], hi[:, 1], hi[:, 2])).all()
            assert np.array_equal(a, b), (mode, int((a != b).sum()), n[a != b][:3], d[a != b][:3], lo[a != b][:3], hi[a != b][:3])
            total += len(n)
    assert total > 5_000_000


def test_corner_in_box_through_distinct_coordinate_masks():
    """Fallback of Frustum::is_intersects_aabb (frustum.rs:238-243): any frustum corner inside the box (inclusive).  The kernels
    evaluate it as the AND over axes of 8-bit masks 'corners whose coordinate on this axis lies in [lo, hi]', built from the
    DISTINCT coordinates per axis (bit-pattern distinct: -0 and +0 stay separate entries)."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        # frustum-like corner sets with many repeated coordinates, zeros of both signs included
        vals = np.array([-0.0, 0.0, -0.01, 0.01, -1.0, 1.0, -120.0, 120.0, 2.5], f32)
        corners = vals[rng.integers(0, len(vals), (8, 3))]
        m = 4000
        a = (rng.normal(size=(m, 3)) * 2).astype(f32)
        b = (rng.normal(size=(m, 3)) * 2).astype(f32)
        snap = rng.random((m, 3)) < 0.3  # box bounds that coincide with corner coordinates (inclusive compares matter)
        a = np.where(snap, vals[rng.integers(0, len(vals), (m, 3))], a)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        want = np.zeros(m, bool)
        for c in range(8):
            want |= ((corners[c] >= lo) & (corners[c] <= hi)).all(1)
        alive = np.full(m, 0xFF, np.uint32)
        for ax in range(3):
            bits = corners[:, ax].view(np.uint32)
            mask = np.zeros(m, np.uint32)
            for u in np.unique(bits):
                v = np.array([u], np.uint32).view(f32)[0]
                cm = np.uint32(sum(1 << i for i in range(8) if bits[i] == u))
                mask |= np.where((v >= lo[:, ax]) & (v <= hi[:, ax]), cm, np.uint32(0))
            alive &= mask
        assert np.array_equal(alive != 0, want), trial
