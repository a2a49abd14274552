/*
 * fyrox_oracle.c — CPU restatement of Fyrox's render-prep hot path.  See fyrox_oracle.h for the
 * rules (test infrastructure only) and the parity status.  Every function cites the reference
 * source (paths relative to /root/reference) it follows.
 *
 * Arithmetic: IEEE-754 binary32, one rounding per * and per + (compile with -ffp-contract=off).
 * Structure: like the reference, the graph is a pool of separately heap-allocated nodes with
 * children vectors and recursive DFS — this file doubles as the "ref-faithful-1T" CPU baseline.
 */
#include "fyrox_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* nalgebra pieces                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* nalgebra Vector3::dot, 3-element fast path: a + b + c, left to right. */
static inline float dot3(const float u[3], const float v[3])
{
    float a = u[0] * v[0];
    float b = u[1] * v[1];
    float c = u[2] * v[2];
    return (a + b) + c;
}

/* nalgebra Vector3::cross */
static inline void cross3(const float u[3], const float v[3], float out[3])
{
    out[0] = u[1] * v[2] - u[2] * v[1];
    out[1] = u[2] * v[0] - u[0] * v[2];
    out[2] = u[0] * v[1] - u[1] * v[0];
}

void orc_mat4_identity(float out[16])
{
    memset(out, 0, 16 * sizeof(float));
    out[0] = out[5] = out[10] = out[15] = 1.0f;
}

/* nalgebra Matrix4 * Matrix4 (static 4x4 → gemm → per-column gemv → axcpy), Appendix A1:
 * C[i,j] = ((A[i,0]*B[0,j] + A[i,1]*B[1,j]) + A[i,2]*B[2,j]) + A[i,3]*B[3,j].
 * Call sites: scene/graph/mod.rs:1216, scene/mesh/mod.rs:497,787-788, renderer/bundle.rs:894. */
void orc_mat4_mul(const float a[16], const float b[16], float out[16])
{
    float r[16];
    for (int j = 0; j < 4; ++j) {
        for (int i = 0; i < 4; ++i) {
            float y = a[0 * 4 + i] * b[j * 4 + 0];
            y = y + a[1 * 4 + i] * b[j * 4 + 1];
            y = y + a[2 * 4 + i] * b[j * 4 + 2];
            y = y + a[3 * 4 + i] * b[j * 4 + 3];
            r[j * 4 + i] = y;
        }
    }
    memcpy(out, r, sizeof r);
}

/* nalgebra Matrix4::transform_point (Appendix A11): (M3x3*p + t) / n, n = row3·p + m33, skipped if n == 0.
 * Call site: scene/mesh/mod.rs:515-518. */
void orc_mat4_transform_point(const float m[16], const float p[3], float out[3])
{
    float n = ((m[3] * p[0] + m[7] * p[1]) + m[11] * p[2]) + m[15];
    float t[3];
    for (int i = 0; i < 3; ++i)
        t[i] = ((m[0 + i] * p[0] + m[4 + i] * p[1]) + m[8 + i] * p[2]) + m[12 + i];
    if (n != 0.0f) {
        t[0] = t[0] / n;
        t[1] = t[1] / n;
        t[2] = t[2] / n;
    }
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
}

/* nalgebra UnitQuaternion::to_rotation_matrix; q = (i,j,k,w); out column-major 3x3.
 * Call site: scene/transform.rs:161,424-425. */
void orc_quat_to_rotation_matrix(const float q[4], float r[9])
{
    float i = q[0], j = q[1], k = q[2], w = q[3];
    float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f;
    float ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
    /* Matrix3::new is row-major: rows below */
    float m11 = ww + ii - jj - kk, m12 = ij - wk,           m13 = wj + ik;
    float m21 = wk + ij,           m22 = ww - ii + jj - kk, m23 = jk - wi;
    float m31 = ik - wj,           m32 = wi + jk,           m33 = ww - ii - jj + kk;
    r[0] = m11; r[1] = m21; r[2] = m31;
    r[3] = m12; r[4] = m22; r[5] = m32;
    r[6] = m13; r[7] = m23; r[8] = m33;
}

/* Input generators (host side, not on the parity path): classic right-handed look-at and
 * nalgebra Perspective3::new / Orthographic3::new formulas (scene/camera.rs:100,459). */
void orc_look_at_rh(const float eye[3], const float target[3], const float up[3], float out[16])
{
    float f[3] = { target[0] - eye[0], target[1] - eye[1], target[2] - eye[2] };
    float fl = sqrtf(dot3(f, f));
    f[0] /= fl; f[1] /= fl; f[2] /= fl;
    float s[3];
    cross3(f, up, s);
    float sl = sqrtf(dot3(s, s));
    s[0] /= sl; s[1] /= sl; s[2] /= sl;
    float u[3];
    cross3(s, f, u);
    out[0] = s[0]; out[4] = s[1]; out[8]  = s[2]; out[12] = -dot3(s, eye);
    out[1] = u[0]; out[5] = u[1]; out[9]  = u[2]; out[13] = -dot3(u, eye);
    out[2] = -f[0]; out[6] = -f[1]; out[10] = -f[2]; out[14] = dot3(f, eye);
    out[3] = 0.0f; out[7] = 0.0f; out[11] = 0.0f; out[15] = 1.0f;
}

void orc_perspective(float aspect, float fovy, float znear, float zfar, float out[16])
{
    memset(out, 0, 16 * sizeof(float));
    float m22 = 1.0f / tanf(fovy / 2.0f);
    out[5] = m22;                                   /* (1,1) */
    out[0] = m22 / aspect;                          /* (0,0) */
    out[10] = (zfar + znear) / (znear - zfar);      /* (2,2) */
    out[14] = zfar * znear * 2.0f / (znear - zfar); /* (2,3) */
    out[11] = -1.0f;                                /* (3,2) */
    out[15] = 0.0f;
}

void orc_orthographic(float l, float r, float b, float t, float zn, float zf, float out[16])
{
    memset(out, 0, 16 * sizeof(float));
    out[0] = 2.0f / (r - l);
    out[5] = 2.0f / (t - b);
    out[10] = -2.0f / (zf - zn);
    out[12] = -(r + l) / (r - l);
    out[13] = -(t + b) / (t - b);
    out[14] = -(zf + zn) / (zf - zn);
    out[15] = 1.0f;
}

/* ------------------------------------------------------------------------------------------ */
/* fyrox-math/src/plane.rs                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* Plane::from_abcd — plane.rs:63-75 */
int orc_plane_from_abcd(float a, float b, float c, float d, orc_plane *out)
{
    float nrm[3] = { a, b, c };
    float len = sqrtf(dot3(nrm, nrm)); /* Vector3::norm = sqrt(norm_squared), Appendix A5 */
    if (len == 0.0f)
        return 0;
    float coeff = 1.0f / len;
    out->n[0] = a * coeff;
    out->n[1] = b * coeff;
    out->n[2] = c * coeff;
    out->d = d * coeff;
    return 1;
}

/* Plane::dot — plane.rs:78-80 */
float orc_plane_dot(const orc_plane *p, const float pt[3])
{
    return dot3(p->n, pt) + p->d;
}

/* Plane::intersection_point — plane.rs:94-102 */
void orc_plane_intersection_point(const orc_plane *a, const orc_plane *b, const orc_plane *c, float out[3])
{
    float bc[3], ca[3], ab[3];
    cross3(b->n, c->n, bc);
    float f = -1.0f / dot3(a->n, bc);
    cross3(c->n, a->n, ca);
    cross3(a->n, b->n, ab);
    for (int i = 0; i < 3; ++i) {
        float v1 = bc[i] * a->d;
        float v2 = ca[i] * b->d;
        float v3 = ab[i] * c->d;
        out[i] = ((v1 + v2) + v3) * f;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* fyrox-math/src/frustum.rs                                                                   */
/* ------------------------------------------------------------------------------------------ */

enum { F_LEFT = 0, F_RIGHT = 1, F_TOP = 2, F_BOTTOM = 3, F_FAR = 4, F_NEAR = 5 };

/* Frustum::from_view_projection_matrix — frustum.rs:54-82 (m[] is nalgebra linear = column-major) */
int orc_frustum_from_view_projection_matrix(const float m[16], orc_frustum *out)
{
    orc_plane *p = out->planes;
    if (!orc_plane_from_abcd(m[3] + m[0], m[7] + m[4], m[11] + m[8], m[15] + m[12], &p[F_LEFT])) return 0;
    if (!orc_plane_from_abcd(m[3] - m[0], m[7] - m[4], m[11] - m[8], m[15] - m[12], &p[F_RIGHT])) return 0;
    if (!orc_plane_from_abcd(m[3] - m[1], m[7] - m[5], m[11] - m[9], m[15] - m[13], &p[F_TOP])) return 0;
    if (!orc_plane_from_abcd(m[3] + m[1], m[7] + m[5], m[11] + m[9], m[15] + m[13], &p[F_BOTTOM])) return 0;
    if (!orc_plane_from_abcd(m[3] - m[2], m[7] - m[6], m[11] - m[10], m[15] - m[14], &p[F_FAR])) return 0;
    if (!orc_plane_from_abcd(m[3] + m[2], m[7] + m[6], m[11] + m[10], m[15] + m[14], &p[F_NEAR])) return 0;

    orc_plane_intersection_point(&p[F_LEFT], &p[F_TOP], &p[F_FAR], out->corners[0]);
    orc_plane_intersection_point(&p[F_LEFT], &p[F_BOTTOM], &p[F_FAR], out->corners[1]);
    orc_plane_intersection_point(&p[F_RIGHT], &p[F_BOTTOM], &p[F_FAR], out->corners[2]);
    orc_plane_intersection_point(&p[F_RIGHT], &p[F_TOP], &p[F_FAR], out->corners[3]);
    orc_plane_intersection_point(&p[F_LEFT], &p[F_TOP], &p[F_NEAR], out->corners[4]);
    orc_plane_intersection_point(&p[F_LEFT], &p[F_BOTTOM], &p[F_NEAR], out->corners[5]);
    orc_plane_intersection_point(&p[F_RIGHT], &p[F_BOTTOM], &p[F_NEAR], out->corners[6]);
    orc_plane_intersection_point(&p[F_RIGHT], &p[F_TOP], &p[F_NEAR], out->corners[7]);
    return 1;
}

/* Frustum::default — frustum.rs:32-43 */
void orc_frustum_default(orc_frustum *out)
{
    float m[16];
    orc_perspective(1.0f, 1.57079632679489661923f /* FRAC_PI_2 */, 0.01f, 1024.0f, m);
    orc_frustum_from_view_projection_matrix(m, out);
}

/* Frustum::is_intersects_point_cloud — frustum.rs:205-219 */
int orc_frustum_is_intersects_point_cloud(const orc_frustum *f, const float *pts, size_t n)
{
    for (int pl = 0; pl < 6; ++pl) {
        size_t back_points = 0;
        for (size_t i = 0; i < n; ++i) {
            if (orc_plane_dot(&f->planes[pl], pts + 3 * i) <= 0.0f) {
                back_points += 1;
                if (back_points >= n)
                    return 0;
            }
        }
    }
    return 1;
}

static void aabb_corners_for_cull(const orc_aabb *a, float c[8][3])
{
    /* corner order of frustum.rs:223-232 (same as AABB::corners, aabb.rs) */
    const float *mn = a->min, *mx = a->max;
    float t[8][3] = {
        { mn[0], mn[1], mn[2] }, { mn[0], mn[1], mx[2] }, { mx[0], mn[1], mx[2] }, { mx[0], mn[1], mn[2] },
        { mn[0], mx[1], mn[2] }, { mn[0], mx[1], mx[2] }, { mx[0], mx[1], mx[2] }, { mx[0], mx[1], mn[2] },
    };
    memcpy(c, t, sizeof t);
}

/* Frustum::is_intersects_aabb — frustum.rs:222-245 */
int orc_frustum_is_intersects_aabb(const orc_frustum *f, const orc_aabb *aabb)
{
    float corners[8][3];
    aabb_corners_for_cull(aabb, corners);
    if (orc_frustum_is_intersects_point_cloud(f, &corners[0][0], 8))
        return 1;
    for (int i = 0; i < 8; ++i)
        if (orc_aabb_is_contains_point(aabb, f->corners[i]))
            return 1;
    return 0;
}

/* Frustum::is_intersects_aabb_offset — frustum.rs:247-275 */
int orc_frustum_is_intersects_aabb_offset(const orc_frustum *f, const orc_aabb *aabb, const float off[3])
{
    float corners[8][3];
    aabb_corners_for_cull(aabb, corners);
    for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 3; ++k)
            corners[i][k] = corners[i][k] + off[k];
    if (orc_frustum_is_intersects_point_cloud(f, &corners[0][0], 8))
        return 1;
    for (int i = 0; i < 8; ++i)
        if (orc_aabb_is_contains_point(aabb, f->corners[i]))
            return 1;
    return 0;
}

/* Frustum::is_contains_point — frustum.rs (all planes: dot > 0) */
int orc_frustum_is_contains_point(const orc_frustum *f, const float pt[3])
{
    for (int pl = 0; pl < 6; ++pl)
        if (orc_plane_dot(&f->planes[pl], pt) <= 0.0f)
            return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* fyrox-math/src/aabb.rs                                                                      */
/* ------------------------------------------------------------------------------------------ */

void orc_aabb_default(orc_aabb *a) /* aabb.rs:31-39 */
{
    a->min[0] = a->min[1] = a->min[2] = FLT_MAX;
    a->max[0] = a->max[1] = a->max[2] = -FLT_MAX;
}

void orc_aabb_unit(orc_aabb *a) /* aabb.rs:42-45 */
{
    a->min[0] = a->min[1] = a->min[2] = -0.5f;
    a->max[0] = a->max[1] = a->max[2] = 0.5f;
}

void orc_aabb_add_point(orc_aabb *s, const float a[3]) /* aabb.rs:86-106 */
{
    if (a[0] < s->min[0]) s->min[0] = a[0];
    if (a[1] < s->min[1]) s->min[1] = a[1];
    if (a[2] < s->min[2]) s->min[2] = a[2];
    if (a[0] > s->max[0]) s->max[0] = a[0];
    if (a[1] > s->max[1]) s->max[1] = a[1];
    if (a[2] > s->max[2]) s->max[2] = a[2];
}

void orc_aabb_add_box(orc_aabb *s, const orc_aabb *b) /* aabb.rs add_box: add_point(min), add_point(max) */
{
    orc_aabb_add_point(s, b->min);
    orc_aabb_add_point(s, b->max);
}

void orc_aabb_corners(const orc_aabb *a, float out[8][3]) { aabb_corners_for_cull(a, out); }

static int vec_all_nan_or_inf(const float v[3]) /* aabb.rs:166-169: x.iter().all(|e| nan || inf) */
{
    for (int i = 0; i < 3; ++i)
        if (!(isnan(v[i]) || isinf(v[i]))) return 0;
    return 1;
}

int orc_aabb_is_valid(const orc_aabb *a) /* aabb.rs:164-176 */
{
    return a->max[0] >= a->min[0] && a->max[1] >= a->min[1] && a->max[2] >= a->min[2] &&
           !vec_all_nan_or_inf(a->min) && !vec_all_nan_or_inf(a->max);
}

int orc_aabb_is_degenerate(const orc_aabb *a) /* aabb.rs:179-181: max == min */
{
    return a->max[0] == a->min[0] && a->max[1] == a->min[1] && a->max[2] == a->min[2];
}

int orc_aabb_is_contains_point(const orc_aabb *a, const float p[3]) /* aabb.rs:193-200 */
{
    return p[0] >= a->min[0] && p[0] <= a->max[0] && p[1] >= a->min[1] && p[1] <= a->max[1] &&
           p[2] >= a->min[2] && p[2] <= a->max[2];
}

/* AxisAlignedBoundingBox::transform — aabb.rs:264-287; basis()/position() lib.rs:767-774 */
void orc_aabb_transform(const orc_aabb *s, const float m[16], orc_aabb *out)
{
    orc_aabb t;
    for (int i = 0; i < 3; ++i)
        t.min[i] = t.max[i] = m[12 + i];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            float bij = m[j * 4 + i];
            float a = bij * s->min[j];
            float b = bij * s->max[j];
            if (a < b) {
                t.min[i] += a;
                t.max[i] += b;
            } else {
                t.min[i] += b;
                t.max[i] += a;
            }
        }
    }
    *out = t;
}

/* ------------------------------------------------------------------------------------------ */
/* fyrox-impl/src/scene/transform.rs                                                           */
/* ------------------------------------------------------------------------------------------ */

void orc_transform_identity(orc_transform *t) /* transform.rs:176-200 */
{
    memset(t, 0, sizeof *t);
    t->local_rotation[3] = 1.0f;
    t->pre_rotation[3] = 1.0f;
    t->local_scale[0] = t->local_scale[1] = t->local_scale[2] = 1.0f;
    t->post_rotation_matrix[0] = t->post_rotation_matrix[4] = t->post_rotation_matrix[8] = 1.0f;
}

/* Transform::calculate_local_transform — transform.rs:421-540 (expression order kept verbatim;
 * Rust's a + b + c - d ... is left-associative) */
void orc_transform_calculate_local(const orc_transform *t, float out[16])
{
    const float *por = t->post_rotation_matrix;
    float pr[9], r[9];
    orc_quat_to_rotation_matrix(t->pre_rotation, pr);
    orc_quat_to_rotation_matrix(t->local_rotation, r);

    float sx = t->local_scale[0], sy = t->local_scale[1], sz = t->local_scale[2];
    float tx = t->local_position[0], ty = t->local_position[1], tz = t->local_position[2];
    float rpx = t->rotation_pivot[0], rpy = t->rotation_pivot[1], rpz = t->rotation_pivot[2];
    float rox = t->rotation_offset[0], roy = t->rotation_offset[1], roz = t->rotation_offset[2];
    float spx = t->scaling_pivot[0], spy = t->scaling_pivot[1], spz = t->scaling_pivot[2];
    float sox = t->scaling_offset[0], soy = t->scaling_offset[1], soz = t->scaling_offset[2];

    float a0 = pr[0] * r[0] + pr[3] * r[1] + pr[6] * r[2];
    float a1 = pr[1] * r[0] + pr[4] * r[1] + pr[7] * r[2];
    float a2 = pr[2] * r[0] + pr[5] * r[1] + pr[8] * r[2];
    float a3 = pr[0] * r[3] + pr[3] * r[4] + pr[6] * r[5];
    float a4 = pr[1] * r[3] + pr[4] * r[4] + pr[7] * r[5];
    float a5 = pr[2] * r[3] + pr[5] * r[4] + pr[8] * r[5];
    float a6 = pr[0] * r[6] + pr[3] * r[7] + pr[6] * r[8];
    float a7 = pr[1] * r[6] + pr[4] * r[7] + pr[7] * r[8];
    float a8 = pr[2] * r[6] + pr[5] * r[7] + pr[8] * r[8];
    float f0 = por[0] * a0 + por[1] * a3 + por[2] * a6;
    float f1 = por[0] * a1 + por[1] * a4 + por[2] * a7;
    float f2 = por[0] * a2 + por[1] * a5 + por[2] * a8;
    float f3 = por[3] * a0 + por[4] * a3 + por[5] * a6;
    float f4 = por[3] * a1 + por[4] * a4 + por[5] * a7;
    float f5 = por[3] * a2 + por[4] * a5 + por[5] * a8;
    float f6 = por[6] * a0 + por[7] * a3 + por[8] * a6;
    float f7 = por[6] * a1 + por[7] * a4 + por[8] * a7;
    float f8 = por[6] * a2 + por[7] * a5 + por[8] * a8;
    float m0 = sx * f0, m1 = sx * f1, m2 = sx * f2, m3 = 0.0f;
    float m4 = sy * f3, m5 = sy * f4, m6 = sy * f5, m7 = 0.0f;
    float m8 = sz * f6, m9 = sz * f7, m10 = sz * f8, m11 = 0.0f;
    float k0 = spx * f0, k1 = spy * f3, k2 = spz * f6;
    float m12 = rox + rpx + tx - rpx * f0 - rpy * f3 - rpz * f6 + sox * f0 + k0 + soy * f3 + k1 + soz * f6 + k2 -
                sx * k0 - sy * k1 - sz * k2;
    float k3 = spx * f1, k4 = spy * f4, k5 = spz * f7;
    float m13 = roy + rpy + ty - rpx * f1 - rpy * f4 - rpz * f7 + sox * f1 + k3 + soy * f4 + k4 + soz * f7 + k5 -
                sx * k3 - sy * k4 - sz * k5;
    float k6 = spx * f2, k7 = spy * f5, k8 = spz * f8;
    float m14 = roz + rpz + tz - rpx * f2 - rpy * f5 - rpz * f8 + sox * f2 + k6 + soy * f5 + k7 + soz * f8 + k8 -
                sx * k6 - sy * k7 - sz * k8;
    float m15 = 1.0f;
    /* Matrix4::new(m0, m4, m8, m12, m1, ...) is row-major ⇒ column-major storage is m0..m15 in order */
    float m[16] = { m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15 };
    memcpy(out, m, sizeof m);
}

/* RenderContext::calculate_sorting_index — renderer/bundle.rs:118-127 */
uint64_t orc_calculate_sorting_index(const float view[16], const float gp[3])
{
    const uint64_t RANGE_CENTER = UINT64_MAX / 2;
    float v[3];
    orc_mat4_transform_point(view, gp, v);
    float zf = v[2] * 1000.0f;
    int64_t d;
    /* Rust `as i64` saturates and maps NaN to 0 */
    if (isnan(zf)) d = 0;
    else if (zf >= 9223372036854775807.0f) d = INT64_MAX;
    else if (zf <= -9223372036854775808.0f) d = INT64_MIN;
    else d = (int64_t)zf;
    /* u64::saturating_add_signed */
    if (d >= 0) {
        uint64_t ud = (uint64_t)d;
        return (RANGE_CENTER > UINT64_MAX - ud) ? UINT64_MAX : RANGE_CENTER + ud;
    } else {
        uint64_t ud = (uint64_t)(-(d + 1)) + 1u;
        return (ud > RANGE_CENTER) ? 0 : RANGE_CENTER - ud;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* scene graph                                                                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t n_bones;
    uint32_t *bones;
    uint32_t n_verts;
    unsigned char *verts; /* owned copy, interleaved */
    orc_vertex_layout layout;
} orc_surface;

typedef struct orc_node {
    int kind;
    /* Base (scene/base.rs:389-483) */
    float local_matrix[16]; /* Transform::matrix() cache */
    int visibility, enabled, frustum_culling, cast_shadows, is_light;
    int is_probe;     /* ReflectionProbe node */
    int static_batch; /* BatchingMode::Static: collect_render_data returns RdcControlFlow::Break when the mesh is rendered */
    uint32_t render_mask;
    uint32_t parent;
    uint32_t *children;
    uint32_t n_children, cap_children;
    float global_transform[16];
    int global_visibility, global_enabled;
    float inv_bind_pose[16];
    /* Mesh (scene/mesh/mod.rs:328-377) */
    orc_aabb local_bounding_box;
    orc_aabb world_bounding_box;
    orc_surface *surfaces;
    uint32_t n_surfaces;
    /* Base::lod_group (scene/base.rs:129-160,417-418): levels of (begin, end, objects) */
    uint32_t n_lod_levels;
    float *lod_begin, *lod_end;
    uint32_t *lod_obj_begin; /* n_lod_levels + 1 offsets into lod_objects */
    uint32_t *lod_objects;
} orc_node;

enum { MSG_TRANSFORM = 1, MSG_VISIBILITY = 2, MSG_ENABLED = 4 };
typedef struct { uint32_t node; uint8_t kind; } orc_msg;

struct orc_graph {
    orc_node **records; /* pool: index → heap node or NULL (free) */
    uint32_t capacity, cap_alloc;
    uint32_t root;
    orc_msg *msgs;
    size_t n_msgs, cap_msgs;
};

static orc_node *node_at(const orc_graph *g, uint32_t i)
{
    return (i < g->capacity) ? g->records[i] : NULL; /* Pool::try_borrow */
}

static void push_msg(orc_graph *g, uint32_t node, uint8_t kind)
{
    if (g->n_msgs == g->cap_msgs) {
        g->cap_msgs = g->cap_msgs ? g->cap_msgs * 2 : 64;
        g->msgs = (orc_msg *)realloc(g->msgs, g->cap_msgs * sizeof(orc_msg));
    }
    g->msgs[g->n_msgs].node = node;
    g->msgs[g->n_msgs].kind = kind;
    g->n_msgs++;
}

static orc_node *new_node(int kind)
{
    orc_node *n = (orc_node *)calloc(1, sizeof(orc_node));
    n->kind = kind;
    orc_mat4_identity(n->local_matrix);
    orc_mat4_identity(n->global_transform);
    orc_mat4_identity(n->inv_bind_pose);
    /* BaseBuilder defaults, scene/base.rs:1255-1271,1388,1399-1400 */
    n->visibility = 1; n->enabled = 1; n->frustum_culling = 1; n->cast_shadows = 1;
    n->render_mask = 0xFFFFFFFFu;
    n->parent = ORC_NONE;
    n->global_visibility = 1; n->global_enabled = 1;
    orc_aabb_default(&n->local_bounding_box); /* Mesh with no surfaces: AABB::default() */
    orc_aabb_default(&n->world_bounding_box);
    return n;
}

static uint32_t pool_spawn(orc_graph *g, orc_node *n)
{
    /* Pool::spawn reuses free records first (fyrox-core/src/pool/mod.rs); free-list order is not
     * modelled — first free index is used. */
    for (uint32_t i = 0; i < g->capacity; ++i)
        if (!g->records[i]) { g->records[i] = n; return i; }
    if (g->capacity == g->cap_alloc) {
        g->cap_alloc = g->cap_alloc ? g->cap_alloc * 2 : 16;
        g->records = (orc_node **)realloc(g->records, g->cap_alloc * sizeof(orc_node *));
    }
    g->records[g->capacity] = n;
    return g->capacity++;
}

orc_graph *orc_graph_new(void)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->root = ORC_NONE;
    orc_graph_add_node(g, ORC_KIND_PIVOT); /* graph/mod.rs:408-424: root pivot, index 0 */
    return g;
}

static void free_node(orc_node *n)
{
    if (!n) return;
    for (uint32_t s = 0; s < n->n_surfaces; ++s) {
        free(n->surfaces[s].bones);
        free(n->surfaces[s].verts);
    }
    free(n->surfaces);
    free(n->children);
    free(n->lod_begin);
    free(n->lod_end);
    free(n->lod_obj_begin);
    free(n->lod_objects);
    free(n);
}

void orc_graph_free(orc_graph *g)
{
    if (!g) return;
    for (uint32_t i = 0; i < g->capacity; ++i) free_node(g->records[i]);
    free(g->records);
    free(g->msgs);
    free(g);
}

uint32_t orc_graph_capacity(const orc_graph *g) { return g->capacity; }
uint32_t orc_graph_root(const orc_graph *g) { return g->root; }

static void children_push(orc_node *p, uint32_t c)
{
    if (p->n_children == p->cap_children) {
        p->cap_children = p->cap_children ? p->cap_children * 2 : 4;
        p->children = (uint32_t *)realloc(p->children, p->cap_children * sizeof(uint32_t));
    }
    p->children[p->n_children++] = c;
}

/* Graph::isolate_node — graph/mod.rs:2143-2160 */
static void isolate_node(orc_graph *g, uint32_t h)
{
    orc_node *n = node_at(g, h);
    uint32_t ph = n->parent;
    n->parent = ORC_NONE;
    orc_node *p = node_at(g, ph);
    if (p) {
        for (uint32_t i = 0; i < p->n_children; ++i) {
            if (p->children[i] == h) {
                memmove(p->children + i, p->children + i + 1, (p->n_children - i - 1) * sizeof(uint32_t));
                p->n_children--;
                break;
            }
        }
    }
}

/* Graph::link_nodes — graph/mod.rs:2114-2131 */
void orc_graph_link_nodes(orc_graph *g, uint32_t child, uint32_t parent)
{
    isolate_node(g, child);
    node_at(g, child)->parent = parent;
    children_push(node_at(g, parent), child);
    push_msg(g, child, MSG_TRANSFORM);
}

/* Graph::add_node(_at_handle) — graph/mod.rs:2044-2088 + Base::on_connected_to_graph base.rs:520-540 */
uint32_t orc_graph_add_node(orc_graph *g, int kind)
{
    uint32_t h = pool_spawn(g, new_node(kind));
    if (g->root == ORC_NONE)
        g->root = h;
    else
        orc_graph_link_nodes(g, h, g->root);
    push_msg(g, h, MSG_TRANSFORM);
    push_msg(g, h, MSG_VISIBILITY);
    push_msg(g, h, MSG_ENABLED);
    return h;
}

/* Graph::remove_node — graph/mod.rs:2091-2111 */
void orc_graph_remove_node(orc_graph *g, uint32_t node)
{
    if (!node_at(g, node)) return;
    isolate_node(g, node);
    uint32_t *stack = NULL;
    size_t ns = 0, cs = 0;
#define PUSH(x) do { if (ns == cs) { cs = cs ? cs * 2 : 16; stack = (uint32_t *)realloc(stack, cs * 4); } stack[ns++] = (x); } while (0)
    PUSH(node);
    while (ns) {
        uint32_t h = stack[--ns];
        orc_node *n = node_at(g, h);
        if (!n) continue;
        for (uint32_t i = 0; i < n->n_children; ++i) PUSH(n->children[i]);
        free_node(n);
        g->records[h] = NULL;
    }
#undef PUSH
    free(stack);
}

orc_graph *orc_graph_build(uint32_t capacity, const uint32_t *parent, const uint32_t *flags,
                           const uint32_t *render_mask, const float *local_m16, const float *local_aabb6)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->root = capacity ? 0 : ORC_NONE;
    g->capacity = g->cap_alloc = capacity;
    g->records = (orc_node **)calloc(capacity ? capacity : 1, sizeof(orc_node *));
    for (uint32_t i = 0; i < capacity; ++i) {
        uint32_t f = flags ? flags[i] : (ORC_FLAG_VISIBILITY | ORC_FLAG_ENABLED | ORC_FLAG_FRUSTUM_CULLING |
                                         ORC_FLAG_CAST_SHADOWS | ORC_FLAG_ALIVE);
        if (!(f & ORC_FLAG_ALIVE)) continue; /* free pool record */
        orc_node *n = new_node((f & ORC_FLAG_RENDERABLE) ? ORC_KIND_MESH : ORC_KIND_PIVOT);
        n->visibility = !!(f & ORC_FLAG_VISIBILITY);
        n->enabled = !!(f & ORC_FLAG_ENABLED);
        n->frustum_culling = !!(f & ORC_FLAG_FRUSTUM_CULLING);
        n->cast_shadows = !!(f & ORC_FLAG_CAST_SHADOWS);
        n->is_light = !!(f & ORC_FLAG_LIGHT);
        n->static_batch = !!(f & ORC_FLAG_STATIC_BATCH);
        n->is_probe = !!(f & ORC_FLAG_REFLECTION_PROBE);
        if (render_mask) n->render_mask = render_mask[i];
        if (local_m16) memcpy(n->local_matrix, local_m16 + 16 * (size_t)i, 64);
        if (local_aabb6 && n->kind == ORC_KIND_MESH) {
            memcpy(n->local_bounding_box.min, local_aabb6 + 6 * (size_t)i, 12);
            memcpy(n->local_bounding_box.max, local_aabb6 + 6 * (size_t)i + 3, 12);
        }
        g->records[i] = n;
    }
    for (uint32_t i = 0; i < capacity; ++i) {
        orc_node *n = g->records[i];
        if (!n || !parent || parent[i] == ORC_NONE) continue;
        n->parent = parent[i];
        orc_node *p = node_at(g, parent[i]);
        if (p) children_push(p, i);
    }
    return g;
}

/* ---- setters ---- */
void orc_node_set_local_matrix(orc_graph *g, uint32_t i, const float m[16])
{
    orc_node *n = node_at(g, i);
    if (!n) return;
    memcpy(n->local_matrix, m, 64);
    push_msg(g, i, MSG_TRANSFORM); /* local_transform_mut() → TrackedProperty::deref_mut, base.rs:343-352 */
}
void orc_node_set_local_transform(orc_graph *g, uint32_t i, const orc_transform *t)
{
    float m[16];
    orc_transform_calculate_local(t, m);
    orc_node_set_local_matrix(g, i, m);
}
void orc_node_set_visibility(orc_graph *g, uint32_t i, int v)
{
    orc_node *n = node_at(g, i);
    if (!n) return;
    n->visibility = !!v;
    push_msg(g, i, MSG_VISIBILITY);
}
void orc_node_set_enabled(orc_graph *g, uint32_t i, int v)
{
    orc_node *n = node_at(g, i);
    if (!n) return;
    n->enabled = !!v;
    push_msg(g, i, MSG_ENABLED);
}
void orc_node_set_frustum_culling(orc_graph *g, uint32_t i, int v) { orc_node *n = node_at(g, i); if (n) n->frustum_culling = !!v; }
void orc_node_set_cast_shadows(orc_graph *g, uint32_t i, int v) { orc_node *n = node_at(g, i); if (n) n->cast_shadows = !!v; }
void orc_node_set_render_mask(orc_graph *g, uint32_t i, uint32_t m) { orc_node *n = node_at(g, i); if (n) n->render_mask = m; }
void orc_node_set_inv_bind_pose(orc_graph *g, uint32_t i, const float m[16]) { orc_node *n = node_at(g, i); if (n) memcpy(n->inv_bind_pose, m, 64); }
void orc_mesh_set_local_aabb(orc_graph *g, uint32_t i, const orc_aabb *a) { orc_node *n = node_at(g, i); if (n) n->local_bounding_box = *a; }

uint32_t orc_mesh_add_surface(orc_graph *g, uint32_t mesh, uint32_t n_bones, const uint32_t *bones,
                              uint32_t n_verts, const void *verts, const orc_vertex_layout *layout)
{
    orc_node *n = node_at(g, mesh);
    if (!n) return ORC_NONE;
    n->surfaces = (orc_surface *)realloc(n->surfaces, (n->n_surfaces + 1) * sizeof(orc_surface));
    orc_surface *s = &n->surfaces[n->n_surfaces];
    memset(s, 0, sizeof *s);
    s->n_bones = n_bones;
    if (n_bones) {
        s->bones = (uint32_t *)malloc(n_bones * sizeof(uint32_t));
        memcpy(s->bones, bones, n_bones * sizeof(uint32_t));
    }
    if (verts && n_verts && layout) {
        s->n_verts = n_verts;
        s->layout = *layout;
        s->verts = (unsigned char *)malloc((size_t)n_verts * layout->stride);
        memcpy(s->verts, verts, (size_t)n_verts * layout->stride);
    }
    return n->n_surfaces++;
}

/* Mesh::local_bounding_box — scene/mesh/mod.rs:631-656, extend_aabb_from_vertex_buffer :557-568 */
void orc_mesh_recalc_local_aabb(orc_graph *g, uint32_t mesh)
{
    orc_node *n = node_at(g, mesh);
    if (!n) return;
    orc_aabb bb;
    orc_aabb_default(&bb);
    for (uint32_t si = 0; si < n->n_surfaces; ++si) {
        const orc_surface *s = &n->surfaces[si];
        for (uint32_t v = 0; v < s->n_verts; ++v) {
            float p[3];
            memcpy(p, s->verts + (size_t)v * s->layout.stride + s->layout.position_offset, 12);
            orc_aabb_add_point(&bb, p);
        }
    }
    n->local_bounding_box = bb;
}

/* ---- hierarchy propagation ---- */

static int mesh_is_skinned(const orc_node *n)
{
    for (uint32_t s = 0; s < n->n_surfaces; ++s)
        if (n->surfaces[s].n_bones) return 1;
    return 0;
}

/* Mesh::on_global_transform_changed — scene/mesh/mod.rs:667-689.  Bone global positions are read at
 * their CURRENT stored value (DFS-order dependent). */
static void on_global_transform_changed(const orc_graph *g, orc_node *n, const float new_g[16])
{
    if (n->kind != ORC_KIND_MESH) return;
    orc_aabb w;
    orc_aabb_transform(&n->local_bounding_box, new_g, &w);
    if (mesh_is_skinned(n)) {
        for (uint32_t s = 0; s < n->n_surfaces; ++s) {
            for (uint32_t b = 0; b < n->surfaces[s].n_bones; ++b) {
                const orc_node *bn = node_at(g, n->surfaces[s].bones[b]);
                if (bn) orc_aabb_add_point(&w, &bn->global_transform[12]); /* global_position() */
            }
        }
    }
    n->world_bounding_box = w;
}

/* Graph::update_global_transform_recursively — scene/graph/mod.rs:1199-1241 */
static void update_global_transform_recursively(orc_graph *g, uint32_t h)
{
    orc_node *n = node_at(g, h);
    if (!n) return;
    float parent_g[16];
    const orc_node *p = node_at(g, n->parent);
    if (p) memcpy(parent_g, p->global_transform, 64);
    else orc_mat4_identity(parent_g);
    float new_g[16];
    orc_mat4_mul(parent_g, n->local_matrix, new_g);
    on_global_transform_changed(g, n, new_g);
    memcpy(n->global_transform, new_g, 64);
    for (uint32_t i = 0; i < n->n_children; ++i)
        update_global_transform_recursively(g, n->children[i]);
}

/* Graph::update_enabled_flag_recursively — scene/graph/mod.rs:1166-1180 */
static void update_enabled_flag_recursively(orc_graph *g, uint32_t h)
{
    orc_node *n = node_at(g, h);
    if (!n) return;
    const orc_node *p = node_at(g, n->parent);
    int parent_enabled = p ? p->global_enabled : 1;
    n->global_enabled = parent_enabled && n->enabled;
    for (uint32_t i = 0; i < n->n_children; ++i)
        update_enabled_flag_recursively(g, n->children[i]);
}

/* Graph::update_visibility_recursively — scene/graph/mod.rs:1182-1197 */
static void update_visibility_recursively(orc_graph *g, uint32_t h)
{
    orc_node *n = node_at(g, h);
    if (!n) return;
    const orc_node *p = node_at(g, n->parent);
    int parent_vis = p ? p->global_visibility : 1;
    n->global_visibility = parent_vis && n->visibility;
    for (uint32_t i = 0; i < n->n_children; ++i)
        update_visibility_recursively(g, n->children[i]);
}

/* Graph::update_hierarchical_data — scene/graph/mod.rs:1272-1292 */
void orc_graph_update_hierarchical_data(orc_graph *g)
{
    update_global_transform_recursively(g, g->root);
    update_enabled_flag_recursively(g, g->root);
    update_visibility_recursively(g, g->root);
}

void orc_graph_drop_messages(orc_graph *g) { g->n_msgs = 0; }

typedef struct { uint32_t node; uint8_t flags; } orc_root;

typedef struct {
    orc_graph *g;
    uint8_t *visited;
    orc_root *roots;
    int64_t *root_slot; /* node index → slot in roots, -1 if the node is not a root candidate */
} msg_ctx;

static void mark_recursive(msg_ctx *c, uint32_t from, uint32_t origin, uint8_t flag)
{
    if (from < c->g->capacity) {
        c->visited[from] |= flag;
        if (from != origin && c->root_slot[from] >= 0) {
            /* remove a descendant from the list of potential roots (graph/mod.rs:1366-1374) */
            orc_root *r = &c->roots[c->root_slot[from]];
            r->flags &= (uint8_t)~flag;
            if (r->flags == 0) c->root_slot[from] = -1; /* roots.remove(&h) */
        }
    }
    orc_node *n = node_at(c->g, from);
    if (n)
        for (uint32_t i = 0; i < n->n_children; ++i)
            mark_recursive(c, n->children[i], origin, flag);
}

/* Graph::process_node_messages — scene/graph/mod.rs:1303-1399.  The reference keeps roots in an
 * FxHashMap (iteration order unspecified); this restatement visits them in insertion order. */
void orc_graph_update(orc_graph *g)
{
    size_t cap = g->capacity ? g->capacity : 1;
    msg_ctx c;
    c.g = g;
    c.visited = (uint8_t *)calloc(cap, 1);
    c.root_slot = (int64_t *)malloc(cap * sizeof(int64_t));
    for (size_t i = 0; i < cap; ++i) c.root_slot[i] = -1;
    c.roots = NULL;
    size_t n_roots = 0, cap_roots = 0;
    for (size_t mi = 0; mi < g->n_msgs; ++mi) {
        orc_msg m = g->msgs[mi];
        if (m.node >= g->capacity) continue;
        if (c.visited[m.node] & m.kind) continue;
        c.visited[m.node] |= m.kind;
        /* roots.entry(node).or_insert(NONE).insert(flag) */
        if (c.root_slot[m.node] < 0) {
            if (n_roots == cap_roots) {
                cap_roots = cap_roots ? cap_roots * 2 : 16;
                c.roots = (orc_root *)realloc(c.roots, cap_roots * sizeof(orc_root));
            }
            c.roots[n_roots].node = m.node;
            c.roots[n_roots].flags = 0;
            c.root_slot[m.node] = (int64_t)n_roots++;
        }
        c.roots[c.root_slot[m.node]].flags |= m.kind;
        mark_recursive(&c, m.node, m.node, m.kind);
    }
    g->n_msgs = 0;
    for (size_t r = 0; r < n_roots; ++r) {
        if (c.root_slot[c.roots[r].node] != (int64_t)r) continue; /* removed (or superseded) entry */
        uint8_t fl = c.roots[r].flags;
        if (fl & MSG_TRANSFORM) update_global_transform_recursively(g, c.roots[r].node);
        if (fl & MSG_VISIBILITY) update_visibility_recursively(g, c.roots[r].node);
        if (fl & MSG_ENABLED) update_enabled_flag_recursively(g, c.roots[r].node);
    }
    free(c.roots);
    free(c.root_slot);
    free(c.visited);
}

/* ---- queries ---- */
void orc_node_global_transform(const orc_graph *g, uint32_t i, float out[16]) { const orc_node *n = node_at(g, i); if (n) memcpy(out, n->global_transform, 64); else orc_mat4_identity(out); }
void orc_node_local_matrix(const orc_graph *g, uint32_t i, float out[16]) { const orc_node *n = node_at(g, i); if (n) memcpy(out, n->local_matrix, 64); else orc_mat4_identity(out); }
int orc_node_global_visibility(const orc_graph *g, uint32_t i) { const orc_node *n = node_at(g, i); return n ? n->global_visibility : 0; }
int orc_node_is_globally_enabled(const orc_graph *g, uint32_t i) { const orc_node *n = node_at(g, i); return n ? n->global_enabled : 0; }
uint32_t orc_node_parent(const orc_graph *g, uint32_t i) { const orc_node *n = node_at(g, i); return n ? n->parent : ORC_NONE; }

/* NodeTrait::world_bounding_box: Mesh → cached (mesh/mod.rs:659-661); Base → unit box transformed by
 * the current global transform (scene/base.rs:741-750) */
void orc_node_world_bounding_box(const orc_graph *g, uint32_t i, orc_aabb *out)
{
    const orc_node *n = node_at(g, i);
    if (!n) { orc_aabb_default(out); return; }
    if (n->kind == ORC_KIND_MESH) { *out = n->world_bounding_box; return; }
    orc_aabb u;
    orc_aabb_unit(&u);
    orc_aabb_transform(&u, n->global_transform, out);
}

/* NodeTrait::should_be_rendered — scene/node/mod.rs:231-256 */
int orc_node_should_be_rendered(const orc_graph *g, uint32_t i, const orc_frustum *f, uint32_t render_mask)
{
    const orc_node *n = node_at(g, i);
    if (!n) return 0;
    if ((n->render_mask & render_mask) == 0) return 0;
    if (!n->global_visibility) return 0;
    if (!n->global_enabled) return 0;
    if (n->frustum_culling && f) {
        orc_aabb w;
        orc_node_world_bounding_box(g, i, &w);
        if (!orc_frustum_is_intersects_aabb(f, &w)) return 0;
    }
    return 1;
}

/* Graph::global_scale — scene/graph/mod.rs:1835-1845 (local scales supplied by the caller, xyz per node) */
void orc_graph_global_scale(const orc_graph *g, uint32_t h, const float *ls, float out[3])
{
    float s[3] = { 1.0f, 1.0f, 1.0f };
    const orc_node *n;
    while ((n = node_at(g, h)) != NULL) {
        s[0] = s[0] * ls[3 * (size_t)h + 0];
        s[1] = s[1] * ls[3 * (size_t)h + 1];
        s[2] = s[2] * ls[3 * (size_t)h + 2];
        h = n->parent;
    }
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2];
}

/* ---- the cull driver ---- */
typedef struct {
    const orc_graph *g;
    const orc_frustum *f;
    uint32_t render_mask;
    int shadow_pass;
    uint32_t *out;
    size_t cap, count;
} cull_ctx;

/* iterate_recursive — renderer/bundle.rs:988-1004; Mesh::collect_render_data — scene/mesh/mod.rs:691-725.
 * A statically batched mesh that IS rendered returns RdcControlFlow::Break (mesh/mod.rs:725): the DFS does not descend
 * into its children; one that fails should_be_rendered / the shadow test returns Continue like everybody else. */
static void iterate_recursive(uint32_t h, cull_ctx *c)
{
    const orc_node *n = node_at(c->g, h);
    if (!n) return;
    if (n->kind == ORC_KIND_MESH) {
        if (orc_node_should_be_rendered(c->g, h, c->f, c->render_mask) && !(c->shadow_pass && !n->cast_shadows)) {
            if (c->count < c->cap) c->out[c->count] = h;
            c->count++;
            if (n->static_batch) return; /* RdcControlFlow::Break */
        }
    }
    for (uint32_t i = 0; i < n->n_children; ++i)
        iterate_recursive(n->children[i], c);
}

size_t orc_from_graph(const orc_graph *g, const orc_frustum *f, uint32_t render_mask, int shadow_pass,
                      uint32_t *out_idx, size_t cap)
{
    cull_ctx c = { g, f, render_mask, shadow_pass, out_idx, cap, 0 };
    if (g->root != ORC_NONE) iterate_recursive(g->root, &c);
    return c.count;
}

/* Base::set_lod_group — scene/base.rs:805-807.  Levels carry the (already clamped, LevelOfDetail::new :74-86) range and
 * the handles of the nodes that represent that level. */
void orc_node_set_lod_group(orc_graph *g, uint32_t i, uint32_t n_levels, const float *begin, const float *end,
                            const uint32_t *obj_begin, const uint32_t *objects)
{
    orc_node *n = node_at(g, i);
    if (!n) return;
    free(n->lod_begin); free(n->lod_end); free(n->lod_obj_begin); free(n->lod_objects);
    n->n_lod_levels = n_levels;
    n->lod_begin = n->lod_end = NULL;
    n->lod_obj_begin = n->lod_objects = NULL;
    if (!n_levels) return;
    n->lod_begin = (float *)malloc(4 * (size_t)n_levels);
    n->lod_end = (float *)malloc(4 * (size_t)n_levels);
    n->lod_obj_begin = (uint32_t *)malloc(4 * ((size_t)n_levels + 1));
    memcpy(n->lod_begin, begin, 4 * (size_t)n_levels);
    memcpy(n->lod_end, end, 4 * (size_t)n_levels);
    memcpy(n->lod_obj_begin, obj_begin, 4 * ((size_t)n_levels + 1));
    uint32_t total = obj_begin[n_levels];
    n->lod_objects = (uint32_t *)malloc(4 * (size_t)(total ? total : 1));
    memcpy(n->lod_objects, objects, 4 * (size_t)total);
}

/* The lod_filter of RenderDataBundleStorage::from_graph — renderer/bundle.rs:898-916: every node starts visible; for
 * every node with a LOD group, in pool order, every object of every level is marked by whether its normalised distance
 * to the observer lies in the level's range (later writes win).  metric_distance = |a - b| (nalgebra: sqrt of the
 * left-to-right dot of the difference). */
void orc_lod_filter(const orc_graph *g, const float observer_translation[3], float z_near, float z_far, uint8_t *filter)
{
    for (uint32_t i = 0; i < g->capacity; ++i) filter[i] = 1;
    for (uint32_t i = 0; i < g->capacity; ++i) {
        const orc_node *n = node_at(g, i);
        if (!n || !n->n_lod_levels) continue;
        for (uint32_t l = 0; l < n->n_lod_levels; ++l)
            for (uint32_t k = n->lod_obj_begin[l]; k < n->lod_obj_begin[l + 1]; ++k) {
                const uint32_t obj = n->lod_objects[k];
                const orc_node *o = node_at(g, obj);
                if (!o) continue; /* try_get_node failed */
                float dx = observer_translation[0] - o->global_transform[12];
                float dy = observer_translation[1] - o->global_transform[13];
                float dz = observer_translation[2] - o->global_transform[14];
                float distance = sqrtf(dx * dx + dy * dy + dz * dz);
                float z_range = z_far - z_near;
                float normalized = (distance - z_near) / z_range;
                filter[obj] = (uint8_t)(normalized >= n->lod_begin[l] && normalized <= n->lod_end[l]);
            }
    }
}

typedef struct {
    cull_ctx c;
    const uint8_t *filter;
} cull_lod_ctx;

/* iterate_recursive with the LOD filter — renderer/bundle.rs:988-1004: a filtered-out node hides its whole sub-tree */
static void iterate_recursive_lod(uint32_t h, cull_lod_ctx *x)
{
    const orc_node *n = node_at(x->c.g, h);
    if (!n) return;
    if (!x->filter[h]) return;
    if (n->kind == ORC_KIND_MESH) {
        if (orc_node_should_be_rendered(x->c.g, h, x->c.f, x->c.render_mask) && !(x->c.shadow_pass && !n->cast_shadows)) {
            if (x->c.count < x->c.cap) x->c.out[x->c.count] = h;
            x->c.count++;
            if (n->static_batch) return; /* RdcControlFlow::Break */
        }
    }
    for (uint32_t i = 0; i < n->n_children; ++i) iterate_recursive_lod(n->children[i], x);
}

size_t orc_from_graph_lod(const orc_graph *g, const orc_frustum *f, uint32_t render_mask, int shadow_pass,
                          const float observer_translation[3], float z_near, float z_far, uint32_t *out_idx, size_t cap)
{
    uint8_t *filter = (uint8_t *)malloc(g->capacity ? g->capacity : 1);
    orc_lod_filter(g, observer_translation, z_near, z_far, filter);
    cull_lod_ctx x = { { g, f, render_mask, shadow_pass, out_idx, cap, 0 }, filter };
    if (g->root != ORC_NONE) iterate_recursive_lod(g->root, &x);
    free(filter);
    return x.c.count;
}

/* ---- palette + LBS ---- */

/* bone_matrices closure — scene/mesh/mod.rs:781-793 (invalid handle ⇒ identity) */
uint32_t orc_mesh_bone_matrices(const orc_graph *g, uint32_t mesh, uint32_t surface, float *out)
{
    const orc_node *n = node_at(g, mesh);
    if (!n || surface >= n->n_surfaces) return 0;
    const orc_surface *s = &n->surfaces[surface];
    for (uint32_t b = 0; b < s->n_bones; ++b) {
        const orc_node *bn = node_at(g, s->bones[b]);
        if (bn) orc_mat4_mul(bn->global_transform, bn->inv_bind_pose, out + 16 * (size_t)b);
        else orc_mat4_identity(out + 16 * (size_t)b);
    }
    return s->n_bones;
}

/* Skinned branch of Mesh::accurate_world_bounding_box — scene/mesh/mod.rs:501-522 (positions);
 * normals: standard.shader:192-195, mat3(m)*n*w accumulated in k order (Appendix A12). */
void orc_skin_vertices(const float *pal, uint32_t n_verts, const void *verts, const orc_vertex_layout *l,
                       float *out_pos, float *out_nrm)
{
    const unsigned char *base = (const unsigned char *)verts;
    for (uint32_t v = 0; v < n_verts; ++v) {
        const unsigned char *vp = base + (size_t)v * l->stride;
        float p[3], nr[3], w[4];
        unsigned char bi[4];
        memcpy(p, vp + l->position_offset, 12);
        memcpy(nr, vp + l->normal_offset, 12);
        memcpy(w, vp + l->bone_weights_offset, 16);
        memcpy(bi, vp + l->bone_indices_offset, 4);
        float pos[3] = { 0.0f, 0.0f, 0.0f };
        float nrm[3] = { 0.0f, 0.0f, 0.0f };
        for (int k = 0; k < 4; ++k) {
            const float *m = pal + 16 * (size_t)bi[k];
            float t[3];
            orc_mat4_transform_point(m, p, t);
            pos[0] += t[0] * w[k];
            pos[1] += t[1] * w[k];
            pos[2] += t[2] * w[k];
            if (out_nrm) {
                for (int i = 0; i < 3; ++i) {
                    float r = (m[0 + i] * nr[0] + m[4 + i] * nr[1]) + m[8 + i] * nr[2];
                    nrm[i] += r * w[k];
                }
            }
        }
        memcpy(out_pos + 3 * (size_t)v, pos, 12);
        if (out_nrm) memcpy(out_nrm + 3 * (size_t)v, nrm, 12);
    }
}

/* IEEE binary16 -> binary32, exact (what texelFetch of an RGB16F texel yields). */
static float orc_half_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                man <<= 1;
                ++e;
            } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* N4 — the blend-shape stage of the standard shader followed by the skinning above
 * (fyrox-material/src/shader/standard/opengl/standard.shader:167-173):
 *   for i in 0..blendShapesCount: inputPosition.xyz += offsets.position * weight; inputNormal += offsets.normal * weight;
 * offsets = S_FetchBlendShapeOffsets(storage, gl_VertexID, i) (fyrox-graphics-gl/src/shaders/shared.glsl:371-378): the
 * three RGB16F texels of vertex v in layer i of the volume texture BlendShapesContainer::from_lists fills
 * (scene/mesh/surface.rs:92-218): record v of the layer = 9 halfs (position, normal, tangent).
 * weight = BlendShape::weight / 100.0 (scene/mesh/mod.rs:794-798), passed in already divided.
 * One rounding per product and per sum (shader semantics, unpinned: GLSL would allow a fused multiply-add). */
void orc_skin_vertices_blend(const float *pal, uint32_t n_verts, const void *verts, const orc_vertex_layout *l, uint32_t n_shapes,
                             const uint16_t *records, uint32_t layer_stride, const float *weights, float *out_pos, float *out_nrm)
{
    const unsigned char *base = (const unsigned char *)verts;
    unsigned char *tmp = (unsigned char *)malloc(l->stride);
    for (uint32_t v = 0; v < n_verts; ++v) {
        memcpy(tmp, base + (size_t)v * l->stride, l->stride);
        float p[3], nr[3];
        memcpy(p, tmp + l->position_offset, 12);
        memcpy(nr, tmp + l->normal_offset, 12);
        for (uint32_t i = 0; i < n_shapes; ++i) {
            const uint16_t *r = records + ((size_t)i * layer_stride + v) * 9;
            const float w = weights[i];
            for (int k = 0; k < 3; ++k) {
                p[k] = p[k] + orc_half_to_float(r[k]) * w;
                nr[k] = nr[k] + orc_half_to_float(r[3 + k]) * w;
            }
        }
        memcpy(tmp + l->position_offset, p, 12);
        memcpy(tmp + l->normal_offset, nr, 12);
        orc_skin_vertices(pal, 1, tmp, l, out_pos + 3 * (size_t)v, out_nrm ? out_nrm + 3 * (size_t)v : NULL);
    }
    free(tmp);
}

uint32_t orc_mesh_skin(const orc_graph *g, uint32_t mesh, uint32_t surface, float *out_pos, float *out_nrm)
{
    const orc_node *n = node_at(g, mesh);
    if (!n || surface >= n->n_surfaces) return 0;
    const orc_surface *s = &n->surfaces[surface];
    if (!s->n_bones || !s->n_verts) return 0;
    float *pal = (float *)malloc((size_t)s->n_bones * 64);
    /* accurate_world_bounding_box indexes graph[b] (panics on invalid); palette closure uses identity */
    orc_mesh_bone_matrices(g, mesh, surface, pal);
    orc_skin_vertices(pal, s->n_verts, s->verts, &s->layout, out_pos, out_nrm);
    free(pal);
    return s->n_verts;
}

int orc_node_is_alive(const orc_graph *g, uint32_t n) { return node_at(g, n) != NULL; }

/* N4 — the reflection-probe part of the node loop of RenderDataBundleStorage::from_graph (renderer/bundle.rs:918-925): for every
 * alive node in pool order that is a ReflectionProbe whose world bounding box contains the observer's translation (inclusive,
 * AxisAlignedBoundingBox::is_contains_point, aabb.rs:193-200), `storage.environment_map = Some(probe)` — the LAST one wins.
 * No visibility / enabled / reachability test there. */
uint32_t orc_select_reflection_probe(const orc_graph *g, const float observer_translation[3])
{
    uint32_t pick = ORC_NONE;
    for (uint32_t i = 0; i < g->capacity; ++i) {
        const orc_node *n = node_at(g, i);
        if (!n || !n->is_probe) continue;
        orc_aabb w;
        orc_node_world_bounding_box(g, i, &w);
        if (orc_aabb_is_contains_point(&w, observer_translation)) pick = i;
    }
    return pick;
}

/* N4 (light list) — the `options.collect_lights` part of RenderDataBundleStorage::from_graph (renderer/bundle.rs:926-974):
 * every alive node, in pool order, that is a BaseLight, whose world bounding box the observer's frustum intersects and
 * that is globally visible and enabled.  No reachability, LOD, render-mask or frustum_culling-flag test here. */
size_t orc_collect_lights(const orc_graph *g, const orc_frustum *f, uint32_t *out, size_t cap)
{
    size_t c = 0;
    for (uint32_t i = 0; i < g->capacity; ++i) {
        const orc_node *n = node_at(g, i);
        if (!n || !n->is_light) continue;
        orc_aabb w;
        orc_node_world_bounding_box(g, i, &w);
        if (orc_frustum_is_intersects_aabb(f, &w) && n->global_visibility && n->global_enabled) {
            if (c < cap) out[c] = i;
            c++;
        }
    }
    return c;
}

/* N3 — what Mesh::collect_render_data pushes for a node with ONE surface (scene/mesh/mod.rs:700 sort index of
 * global_position(); :731-737 world = identity if the surface is skinned, else global_transform()) and what
 * RenderDataBundle::write_uniforms derives from it (renderer/bundle.rs:483-487: world, view_projection * world).
 * A node without surfaces is treated as unskinned.  Returns the sort index. */
uint64_t orc_node_instance(const orc_graph *g, uint32_t node, const float view[16], const float vp[16],
                           float world[16], float wvp[16])
{
    const orc_node *n = node_at(g, node);
    if (!n) {
        orc_mat4_identity(world);
        orc_mat4_mul(vp, world, wvp);
        return 0;
    }
    int skinned = n->static_batch; /* a static batch is pushed with world_transform = identity (mesh/mod.rs:716) */
    for (uint32_t si = 0; si < n->n_surfaces; ++si)
        if (n->surfaces[si].n_bones) skinned = 1;
    if (skinned) orc_mat4_identity(world);
    else memcpy(world, n->global_transform, 64);
    orc_mat4_mul(vp, world, wvp);
    const float gp[3] = {n->global_transform[12], n->global_transform[13], n->global_transform[14]};
    return orc_calculate_sorting_index(view, gp);
}

/* N3, meshes with several surfaces — Mesh::collect_render_data, scene/mesh/mod.rs:726-805: one SurfaceInstanceData per surface;
 * `world = if is_skinned { identity } else { global_transform }` per SURFACE (:731-737), the node's sort index for all of them
 * (:700), a static batch pushes the identity (:716).  Returns the sort index; out_skinned = the surface has bones. */
uint64_t orc_node_surface_instance(const orc_graph *g, uint32_t node, uint32_t surface, const float view[16], const float vp[16],
                                   float world[16], float wvp[16], int *out_skinned)
{
    const orc_node *n = node_at(g, node);
    if (out_skinned) *out_skinned = 0;
    if (!n || surface >= n->n_surfaces) {
        orc_mat4_identity(world);
        orc_mat4_mul(vp, world, wvp);
        return 0;
    }
    const int skinned = n->surfaces[surface].n_bones != 0;
    if (out_skinned) *out_skinned = skinned;
    if (skinned || n->static_batch) orc_mat4_identity(world);
    else memcpy(world, n->global_transform, 64);
    orc_mat4_mul(vp, world, wvp);
    const float gp[3] = {n->global_transform[12], n->global_transform[13], n->global_transform[14]};
    return orc_calculate_sorting_index(view, gp);
}

/* bone_matrices of ONE surface, zero padded to 255 (renderer/bundle.rs:484-496); 0 = the surface is not skinned */
int orc_surface_bone_block(const orc_graph *g, uint32_t mesh, uint32_t surface, float out[255 * 16])
{
    const orc_node *n = node_at(g, mesh);
    if (!n || surface >= n->n_surfaces || !n->surfaces[surface].n_bones) return 0;
    memset(out, 0, 255 * 64);
    orc_mesh_bone_matrices(g, mesh, surface, out);
    return 1;
}

/* N3 — the bone-matrix block RenderDataBundle::write_uniforms uploads for one instance (renderer/bundle.rs:484-496):
 * `matrices = [INIT; MAX_BONE_MATRICES]` (all-zero mat4, MAX_BONE_MATRICES = 255, fyrox-material/src/shader/mod.rs:613),
 * `matrices[0..n].copy_from_slice(&instance.bone_matrices)`; an instance with empty bone_matrices gets no block.
 * Returns 1 and fills out[255*16] if node `mesh` has a skinned surface (its first one), else 0. */
int orc_instance_bone_block(const orc_graph *g, uint32_t mesh, float out[255 * 16])
{
    const orc_node *n = node_at(g, mesh);
    if (!n) return 0;
    for (uint32_t si = 0; si < n->n_surfaces; ++si) {
        if (!n->surfaces[si].n_bones) continue;
        memset(out, 0, 255 * 64); /* +0.0 everywhere */
        orc_mesh_bone_matrices(g, mesh, si, out);
        return 1;
    }
    return 0;
}

/* Mesh::accurate_world_bounding_box — scene/mesh/mod.rs:468-526 */
void orc_mesh_accurate_world_bounding_box(const orc_graph *g, uint32_t mesh, orc_aabb *out)
{
    orc_aabb bb;
    orc_aabb_default(&bb);
    const orc_node *n = node_at(g, mesh);
    if (n) {
        for (uint32_t si = 0; si < n->n_surfaces; ++si) {
            const orc_surface *s = &n->surfaces[si];
            if (!s->n_verts) continue;
            if (!s->n_bones) {
                for (uint32_t v = 0; v < s->n_verts; ++v) {
                    float p[3], t[3];
                    memcpy(p, s->verts + (size_t)v * s->layout.stride + s->layout.position_offset, 12);
                    orc_mat4_transform_point(n->global_transform, p, t);
                    orc_aabb_add_point(&bb, t);
                }
            } else {
                float *pos = (float *)malloc((size_t)s->n_verts * 12);
                orc_mesh_skin(g, mesh, si, pos, NULL);
                for (uint32_t v = 0; v < s->n_verts; ++v) orc_aabb_add_point(&bb, pos + 3 * (size_t)v);
                free(pos);
            }
        }
    }
    *out = bb;
}
