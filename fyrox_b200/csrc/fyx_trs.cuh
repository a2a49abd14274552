// fyx_trs.cuh — Transform::calculate_local_transform on the device (shared by the TRS upload kernels of
// fyx_kernels.cu and the animation kernels of fyx_anim.cu).  Same expression order as the oracle.
#pragma once
#include "fyx_internal.h"

namespace fyx {

__device__ __forceinline__ bool finite4(const float4 v)
{
    return (fabsf(v.x) <= 3.402823466e38f) & (fabsf(v.y) <= 3.402823466e38f) & (fabsf(v.z) <= 3.402823466e38f) &
           (fabsf(v.w) <= 3.402823466e38f);
}

// UnitQuaternion::to_rotation_matrix (nalgebra), column-major 3x3 — same expression order as the oracle
__device__ __forceinline__ void quat_to_rot(const float q[4], float r[9])
{
    const float i = q[0], j = q[1], k = q[2], w = q[3];
    const float ww = FYX_MUL(w, w), ii = FYX_MUL(i, i), jj = FYX_MUL(j, j), kk = FYX_MUL(k, k);
    const float ij = FYX_MUL(FYX_MUL(i, j), 2.0f), wk = FYX_MUL(FYX_MUL(w, k), 2.0f), wj = FYX_MUL(FYX_MUL(w, j), 2.0f);
    const float ik = FYX_MUL(FYX_MUL(i, k), 2.0f), jk = FYX_MUL(FYX_MUL(j, k), 2.0f), wi = FYX_MUL(FYX_MUL(w, i), 2.0f);
    r[0] = FYX_ADD(FYX_ADD(FYX_ADD(ww, ii), -jj), -kk); // a - b is a + (-b) exactly
    r[3] = FYX_ADD(ij, -wk);
    r[6] = FYX_ADD(wj, ik);
    r[1] = FYX_ADD(wk, ij);
    r[4] = FYX_ADD(FYX_ADD(FYX_ADD(ww, -ii), jj), -kk);
    r[7] = FYX_ADD(jk, -wi);
    r[2] = FYX_ADD(ik, -wj);
    r[5] = FYX_ADD(wi, jk);
    r[8] = FYX_ADD(FYX_ADD(FYX_ADD(ww, -ii), -jj), kk);
}

__device__ __forceinline__ float dot3s(const float a0, const float b0, const float a1, const float b1, const float a2, const float b2)
{
    return FYX_ADD(FYX_ADD(FYX_MUL(a0, b0), FYX_MUL(a1, b1)), FYX_MUL(a2, b2)); // x*y + z*w + u*v, left to right
}

// Transform::calculate_local_transform (scene/transform.rs:421-540), expression by expression (Rust's
// a + b - c ... is left-associative; x - y is x + (-y)).  `st` = the node's pivots / offsets / pre- and
// post-rotation (nullptr with HAS_STATICS false: the defaults).
template <bool HAS_STATICS>
__device__ __forceinline__ void trs_to_local(const fyx_trs &t, const fyx_transform_statics *st, Affine &A)
{
    float prq[4] = {0.f, 0.f, 0.f, 1.f};
    float por[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    float ro[3] = {0.f, 0.f, 0.f}, rp[3] = {0.f, 0.f, 0.f}, so[3] = {0.f, 0.f, 0.f}, sp[3] = {0.f, 0.f, 0.f};
    if (HAS_STATICS) {
        const fyx_transform_statics s = *st;
#pragma unroll
        for (int i = 0; i < 4; ++i) prq[i] = s.pre_rotation[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) por[i] = s.post_rotation_matrix[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ro[i] = s.rotation_offset[i];
            rp[i] = s.rotation_pivot[i];
            so[i] = s.scaling_offset[i];
            sp[i] = s.scaling_pivot[i];
        }
    }
    float pr[9], r[9];
    quat_to_rot(prq, pr);
    quat_to_rot(t.rotation, r);
    const float sx = t.scale[0], sy = t.scale[1], sz = t.scale[2];
    float av[9], f[9];
#pragma unroll
    for (int c = 0; c < 3; ++c)       // a(3c + i) = pr[i]*r[3c] + pr[3+i]*r[3c+1] + pr[6+i]*r[3c+2]
#pragma unroll
        for (int i = 0; i < 3; ++i) av[3 * c + i] = dot3s(pr[i], r[3 * c], pr[3 + i], r[3 * c + 1], pr[6 + i], r[3 * c + 2]);
#pragma unroll
    for (int c = 0; c < 3; ++c)       // f(3c + i) = por[3c]*a(i) + por[3c+1]*a(3+i) + por[3c+2]*a(6+i)
#pragma unroll
        for (int i = 0; i < 3; ++i) f[3 * c + i] = dot3s(por[3 * c], av[i], por[3 * c + 1], av[3 + i], por[3 * c + 2], av[6 + i]);
    float m[12]; // m0..m2, m4..m6, m8..m10 and the translation m12..m14
    m[0] = FYX_MUL(sx, f[0]); m[1] = FYX_MUL(sx, f[1]); m[2] = FYX_MUL(sx, f[2]);
    m[3] = FYX_MUL(sy, f[3]); m[4] = FYX_MUL(sy, f[4]); m[5] = FYX_MUL(sy, f[5]);
    m[6] = FYX_MUL(sz, f[6]); m[7] = FYX_MUL(sz, f[7]); m[8] = FYX_MUL(sz, f[8]);
    const float s3[3] = {sx, sy, sz};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // m(12+i) = ro + rp + t - rp.x*f(i) - rp.y*f(3+i) - rp.z*f(6+i) + so.x*f(i) + k0 + so.y*f(3+i) + k1 + so.z*f(6+i) + k2
        //           - sx*k0 - sy*k1 - sz*k2,   k0 = sp.x*f(i), k1 = sp.y*f(3+i), k2 = sp.z*f(6+i)
        const float f0 = f[i], f1 = f[3 + i], f2 = f[6 + i];
        const float k0 = FYX_MUL(sp[0], f0), k1 = FYX_MUL(sp[1], f1), k2 = FYX_MUL(sp[2], f2);
        float v = FYX_ADD(FYX_ADD(ro[i], rp[i]), t.position[i]);
        v = FYX_ADD(v, -FYX_MUL(rp[0], f0));
        v = FYX_ADD(v, -FYX_MUL(rp[1], f1));
        v = FYX_ADD(v, -FYX_MUL(rp[2], f2));
        v = FYX_ADD(v, FYX_MUL(so[0], f0));
        v = FYX_ADD(v, k0);
        v = FYX_ADD(v, FYX_MUL(so[1], f1));
        v = FYX_ADD(v, k1);
        v = FYX_ADD(v, FYX_MUL(so[2], f2));
        v = FYX_ADD(v, k2);
        v = FYX_ADD(v, -FYX_MUL(s3[0], k0));
        v = FYX_ADD(v, -FYX_MUL(s3[1], k1));
        v = FYX_ADD(v, -FYX_MUL(s3[2], k2));
        m[9 + i] = v;
    }
    // Matrix4::new(m0, m4, m8, m12, m1, ...) is row-major: rows of the local matrix
    A.r0 = make_float4(m[0], m[3], m[6], m[9]);
    A.r1 = make_float4(m[1], m[4], m[7], m[10]);
    A.r2 = make_float4(m[2], m[5], m[8], m[11]);

}

} // namespace fyx
