"""k_fold_bones gives one warp to a skinned mesh: lanes take the bones round-robin, keep (value, first position) candidates
with strict compares, and merge them with a butterfly reduction that prefers the smaller position among numerically equal
values.  The claim (fyx_kernels.cu) is that this equals the reference's sequential `add_point` loop
(fyrox-math/src/aabb.rs:86-106: strict < / > updates in bone order) BIT FOR BIT — including which of -0 / +0 survives.
Checked here on the CPU by emulating the warp in numpy float32 on inputs full of ties, zeros of both signs and NaNs."""
import numpy as np

f32 = np.float32


def sequential(start_min, start_max, pts):
    mn, mx = f32(start_min), f32(start_max)
    for p in pts:
        if p < mn:
            mn = p
        if p > mx:
            mx = p
    return mn, mx


def warp_fold(start_min, start_max, pts):
    lanes = 32
    mn = np.full(lanes, start_min, f32)
    mx = np.full(lanes, start_max, f32)
    kmn = np.zeros(lanes, np.int64)
    kmx = np.zeros(lanes, np.int64)
    for b, p in enumerate(pts):  # bone b goes to lane b % 32, key b + 1
        l = b % lanes
        if p < mn[l]:
            mn[l], kmn[l] = p, b + 1
        if p > mx[l]:
            mx[l], kmx[l] = p, b + 1
    o = 16
    while o:
        omn, okmn, omx, okmx = mn.copy(), kmn.copy(), mx.copy(), kmx.copy()
        for l in range(lanes):
            q = l ^ o
            if omn[q] < mn[l] or (omn[q] == mn[l] and okmn[q] < kmn[l]):
                mn[l], kmn[l] = omn[q], okmn[q]
            if omx[q] > mx[l] or (omx[q] == mx[l] and okmx[q] < kmx[l]):
                mx[l], kmx[l] = omx[q], okmx[q]
        o >>= 1
    assert len(set(mn.view(np.uint32).tolist())) == 1 and len(set(mx.view(np.uint32).tolist())) == 1  # every lane agrees
    return mn[0], mx[0]


def test_warp_fold_equals_sequential_add_point_bit_for_bit():
    rng = np.random.default_rng(5)
    pool = np.array([0.0, -0.0, 1.0, -1.0, 1.0, 2.5, -2.5, 1e-45, -1e-45, np.nan, 3.0, -3.0, 0.0, -0.0], f32)
    with np.errstate(invalid="ignore"):
        for trial in range(4000):
            n = int(rng.integers(0, 200))
            pts = pool[rng.integers(0, len(pool), n)] if trial % 2 else rng.normal(size=n).astype(f32)
            smin, smax = pool[rng.integers(0, len(pool) - 5)], pool[rng.integers(0, len(pool) - 5)]
            if np.isnan(smin) or np.isnan(smax):
                continue
            a = sequential(smin, smax, pts)
            b = warp_fold(smin, smax, pts)
            assert a[0].view(np.uint32) == b[0].view(np.uint32) and a[1].view(np.uint32) == b[1].view(np.uint32), (trial, pts[:8], a, b)
