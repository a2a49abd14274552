// fyx_math.cuh — op-order-exact f32 arithmetic shared by the kernels.
//
// The reference is Rust + nalgebra: every `*` and `+` rounds once, nothing is contracted into FMA
// (SURVEY.md Appendix A).  On the device every product/sum therefore goes through __fmul_rn /
// __fadd_rn, which nvcc never fuses (the library is also built with -fmad=false).  Parenthesisation
// below IS the specification; do not "simplify".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fyx {

#if defined(__CUDA_ARCH__)
#define FYX_MUL(a, b) __fmul_rn((a), (b))
#define FYX_ADD(a, b) __fadd_rn((a), (b))
#else
#define FYX_MUL(a, b) ((a) * (b))
#define FYX_ADD(a, b) ((a) + (b))
#endif

// Affine 4x4 kept as its three upper rows; the bottom row is exactly (+0,+0,+0,1) for every matrix
// Transform::matrix() can produce (scene/transform.rs:476-539) and for every product of such
// matrices in nalgebra's order.  r[i] = (M[i,0], M[i,1], M[i,2], M[i,3]).
struct Affine {
    float4 r0, r1, r2;
};

__host__ __device__ __forceinline__ Affine affine_identity()
{
    Affine a;
    a.r0 = make_float4(1.f, 0.f, 0.f, 0.f);
    a.r1 = make_float4(0.f, 1.f, 0.f, 0.f);
    a.r2 = make_float4(0.f, 0.f, 1.f, 0.f);
    return a;
}

// One row of C = A·B for affine A, B in nalgebra's gemm order (Appendix A1):
//   C[i,j] = ((A[i,0]*B[0,j] + A[i,1]*B[1,j]) + A[i,2]*B[2,j]) + A[i,3]*B[3,j]
// with B[3,:] = (+0,+0,+0,1).  The A[i,3]*(+0) term is kept: it turns a -0 partial sum into +0
// exactly as the reference's full 4x4 product does.
__host__ __device__ __forceinline__ float4 affine_mul_row(const float4 a, const Affine &b)
{
    float4 c;
    const float z = FYX_MUL(a.w, 0.0f);
    c.x = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(a.x, b.r0.x), FYX_MUL(a.y, b.r1.x)), FYX_MUL(a.z, b.r2.x)), z);
    c.y = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(a.x, b.r0.y), FYX_MUL(a.y, b.r1.y)), FYX_MUL(a.z, b.r2.y)), z);
    c.z = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(a.x, b.r0.z), FYX_MUL(a.y, b.r1.z)), FYX_MUL(a.z, b.r2.z)), z);
    c.w = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(a.x, b.r0.w), FYX_MUL(a.y, b.r1.w)), FYX_MUL(a.z, b.r2.w)), a.w);
    return c;
}

// Graph::update_global_transform_recursively: G = parent.G * local (scene/graph/mod.rs:1216);
// palette: bone.G * inv_bind (scene/mesh/mod.rs:787-788).
__host__ __device__ __forceinline__ Affine affine_mul(const Affine &a, const Affine &b)
{
    Affine c;
    c.r0 = affine_mul_row(a.r0, b);
    c.r1 = affine_mul_row(a.r1, b);
    c.r2 = affine_mul_row(a.r2, b);
    return c;
}

// One row of AxisAlignedBoundingBox::transform (fyrox-math/src/aabb.rs:264-287, Appendix A4):
// min = max = M[i,3]; for j = 0,1,2: a = M[i,j]*lmin[j], b = M[i,j]*lmax[j]; a<b ? (min+=a,max+=b) : (min+=b,max+=a).
// Returns (min_i, max_i).
__host__ __device__ __forceinline__ float2 aabb_transform_row(const float4 row, const float2 lx, const float2 ly,
                                                              const float2 lz)
{
    float mn = row.w, mx = row.w;
    float a, b;
    a = FYX_MUL(row.x, lx.x); b = FYX_MUL(row.x, lx.y);
    if (a < b) { mn = FYX_ADD(mn, a); mx = FYX_ADD(mx, b); } else { mn = FYX_ADD(mn, b); mx = FYX_ADD(mx, a); }
    a = FYX_MUL(row.y, ly.x); b = FYX_MUL(row.y, ly.y);
    if (a < b) { mn = FYX_ADD(mn, a); mx = FYX_ADD(mx, b); } else { mn = FYX_ADD(mn, b); mx = FYX_ADD(mx, a); }
    a = FYX_MUL(row.z, lz.x); b = FYX_MUL(row.z, lz.y);
    if (a < b) { mn = FYX_ADD(mn, a); mx = FYX_ADD(mx, b); } else { mn = FYX_ADD(mn, b); mx = FYX_ADD(mx, a); }
    return make_float2(mn, mx);
}

// Frustum in kernel-parameter space.  planes as in frustum.rs:26-30; corner_min/max = component-wise
// bounds of the 8 corners (exact min/max, no rounding) used only as a conservative early-out for
// the corner-in-AABB fallback.
struct FrustumDev {
    float4 plane[6];
    float cx[8], cy[8], cz[8];
    float cmin[3], cmax[3];
    uint32_t cam_mask;
    uint32_t pass_flags;
    uint32_t psel;   // bit (3*p + axis): the plane-p normal component on that axis is negative
    uint32_t pad_;
};

// Frustum::is_intersects_aabb (fyrox-math/src/frustum.rs:222-245) on (min,max) pairs per axis.
//
// Cloud test (frustum.rs:205-219): for each plane, all 8 corners have s = (n·p) + d <= 0  ⇒ outside.
// s(corner) = ((nx*px + ny*py) + nz*pz) + d with one rounding per op (plane.rs:78-80, Appendix A6).
// Rounding is monotone, so the corner built from the per-axis larger products has the largest s of
// the eight: "all eight <= 0"  ⇔  that corner's s <= 0 — the same booleans as the reference loop in
// 10 ops per plane instead of 56.  The argument needs NaN-free arithmetic and min <= max: boxes with a
// non-finite or huge (>1e18) bound or an inverted axis take the literal 8-corner loop instead (never in
// practice; the branch is uniform).
__device__ __forceinline__ bool frustum_intersects_aabb(const FrustumDev &f, const float2 x, const float2 y,
                                                        const float2 z)
{
    bool cloud = true;
    const float kBig = 1e18f;
    const bool tame = (fabsf(x.x) <= kBig) & (fabsf(x.y) <= kBig) & (fabsf(y.x) <= kBig) & (fabsf(y.y) <= kBig) &
                      (fabsf(z.x) <= kBig) & (fabsf(z.y) <= kBig) & (x.x <= x.y) & (y.x <= y.y) & (z.x <= z.y);
    if (tame) {
        // n*p is monotone in p (rounding is monotone), so max(fl(n*min), fl(n*max)) is fl(n*max) for n >= 0
        // and fl(n*min) for n < 0: pick the operand first (psel, built on the host) and multiply once.
        // For n == ±0 both products are zeros; either choice gives the same booleans.
        const uint32_t sel = f.psel;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const float4 pl = f.plane[p];
            const float vx = (sel >> (3 * p + 0)) & 1u ? x.x : x.y;
            const float vy = (sel >> (3 * p + 1)) & 1u ? y.x : y.y;
            const float vz = (sel >> (3 * p + 2)) & 1u ? z.x : z.y;
            const float s = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(pl.x, vx), FYX_MUL(pl.y, vy)), FYX_MUL(pl.z, vz)), pl.w);
            cloud &= !(s <= 0.0f);
        }
    } else {
        // literal restatement of the 8-corner loop (NaN-correct)
        const float xs[2] = {x.x, x.y}, ys[2] = {y.x, y.y}, zs[2] = {z.x, z.y};
        for (int p = 0; p < 6; ++p) {
            const float4 pl = f.plane[p];
            int back = 0;
            for (int c = 0; c < 8; ++c) {
                const float s = FYX_ADD(FYX_ADD(FYX_ADD(FYX_MUL(pl.x, xs[c & 1]), FYX_MUL(pl.y, ys[(c >> 1) & 1])),
                                                FYX_MUL(pl.z, zs[(c >> 2) & 1])), pl.w);
                back += (s <= 0.0f) ? 1 : 0;
            }
            if (back >= 8) cloud = false;
        }
    }
    if (cloud) return true;
    // Fallback: any frustum corner inside the AABB, inclusive compares (aabb.rs:193-200).
    // Early-out: if the AABB misses the corners' bounding box on any axis no corner can be inside.
    if (x.y < f.cmin[0] || x.x > f.cmax[0] || y.y < f.cmin[1] || y.x > f.cmax[1] || z.y < f.cmin[2] || z.x > f.cmax[2])
        return false;
    bool inside = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        inside |= (f.cx[c] >= x.x) & (f.cx[c] <= x.y) & (f.cy[c] >= y.x) & (f.cy[c] <= y.y) & (f.cz[c] >= z.x) &
                  (f.cz[c] <= z.y);
    }
    return inside;
}

} // namespace fyx
