// scenegen.cpp — deterministic synthetic scenes (see scenegen.h).  Host-only input generator.
#include "scenegen.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <omp.h>

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
// FYX_NODE_* input bits (include/fyrox_b200.h)
constexpr uint32_t F_VIS = 1u << 0, F_EN = 1u << 1, F_FC = 1u << 2, F_CS = 1u << 3, F_ALIVE = 1u << 4, F_REND = 1u << 5;

inline uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Pcg32 {
    uint64_t state, inc;
    Pcg32(uint64_t seed, uint64_t stream)
    {
        uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull);
        state = splitmix64(x);
        inc = splitmix64(x) | 1ull;
        next();
    }
    uint32_t next()
    {
        const uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    float uni() { return (float)(next() >> 8) * (1.0f / 16777216.0f); } // [0,1)
    float range(float a, float b) { return a + (b - a) * uni(); }
    float gauss()
    {
        float u1 = uni();
        if (u1 < 1e-7f) u1 = 1e-7f;
        const float u2 = uni();
        return std::sqrt(-2.0f * std::log(u1)) * std::cos(6.28318530718f * u2);
    }
};

struct Trs {
    float t[3];
    float q[4]; // i,j,k,w
    float s[3];
};

void quat_normalize(float q[4])
{
    float n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < 1e-12f) { q[0] = q[1] = q[2] = 0.f; q[3] = 1.f; return; }
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

// column-major 4x4 = T * R(q) * S; bottom row exactly (+0,+0,+0,1)
void trs_to_m16(const Trs &x, float m[16])
{
    const float i = x.q[0], j = x.q[1], k = x.q[2], w = x.q[3];
    const float r00 = 1 - 2 * (j * j + k * k), r01 = 2 * (i * j - k * w), r02 = 2 * (i * k + j * w);
    const float r10 = 2 * (i * j + k * w), r11 = 1 - 2 * (i * i + k * k), r12 = 2 * (j * k - i * w);
    const float r20 = 2 * (i * k - j * w), r21 = 2 * (j * k + i * w), r22 = 1 - 2 * (i * i + j * j);
    m[0] = r00 * x.s[0]; m[1] = r10 * x.s[0]; m[2] = r20 * x.s[0]; m[3] = 0.0f;
    m[4] = r01 * x.s[1]; m[5] = r11 * x.s[1]; m[6] = r21 * x.s[1]; m[7] = 0.0f;
    m[8] = r02 * x.s[2]; m[9] = r12 * x.s[2]; m[10] = r22 * x.s[2]; m[11] = 0.0f;
    m[12] = x.t[0]; m[13] = x.t[1]; m[14] = x.t[2]; m[15] = 1.0f;
}

enum Kind { K_ROOT, K_SECTOR, K_GROUP, K_LEAF, K_BONE, K_SKINNED };

} // namespace

struct sg_scene {
    sg_config cfg;
    // global layout
    uint32_t S = 0, G = 0, n_leaves = 0, first_group = 0, first_leaf = 0, first_unit = 0, unit_stride = 0;
    // local arrays
    std::vector<uint32_t> parent, flags, mask, gidx;
    std::vector<float> local, aabb;
    uint32_t n_rend = 0;
    // units
    std::vector<uint32_t> unit_global; // global unit id of each local unit
    std::vector<uint32_t> unit_first;  // local index of bone 0; bones contiguous, mesh = first + B
    std::vector<uint32_t> bone_nodes;  // n_units * B local indices
    std::vector<float> inv_bind;       // n_units * B * 16
    std::vector<Trs> bone_rest;        // n_units * B
};

namespace {

uint32_t bone_parent(uint32_t k) { return k <= 1 ? 0u : 1u + (k - 2u) / 2u; }

void gen_node(const sg_scene &sc, uint32_t gid, Kind kind, Trs &trs, uint32_t &flags, uint32_t &mask, float aabb[6])
{
    Pcg32 r(sc.cfg.seed, gid);
    trs.t[0] = trs.t[1] = trs.t[2] = 0.f;
    trs.q[0] = trs.q[1] = trs.q[2] = 0.f; trs.q[3] = 1.f;
    trs.s[0] = trs.s[1] = trs.s[2] = 1.f;
    float pr = 0.f;
    switch (kind) {
    case K_ROOT: pr = 0.f; break;
    case K_SECTOR: pr = 40.f; break;
    case K_GROUP: pr = 10.f; break;
    default: pr = 3.f; break;
    }
    if (kind != K_ROOT) {
        for (int i = 0; i < 3; ++i) trs.t[i] = r.range(-pr, pr);
        for (int i = 0; i < 4; ++i) trs.q[i] = r.gauss();
        quat_normalize(trs.q);
        if (kind == K_LEAF)
            for (int i = 0; i < 3; ++i) trs.s[i] = r.range(0.5f, 1.5f);
    }
    flags = F_ALIVE;
    if (kind == K_ROOT) {
        flags |= F_VIS | F_EN | F_FC | F_CS;
        mask = 0xFFFFFFFFu;
    } else {
        if (r.uni() < 0.97f) flags |= F_VIS;
        if (r.uni() < 0.99f) flags |= F_EN;
        if (r.uni() < 0.99f) flags |= F_FC;
        if (r.uni() < 0.90f) flags |= F_CS;
        mask = (r.uni() < 0.95f) ? 0xFFFFFFFFu : (1u << (r.next() & 31u));
    }
    if (kind == K_LEAF || kind == K_SKINNED) flags |= F_REND;
    if (kind == K_LEAF) {
        for (int i = 0; i < 3; ++i) {
            const float h = r.range(0.1f, 2.0f);
            aabb[i] = -h;
            aabb[3 + i] = h;
        }
    } else if (kind == K_SKINNED) { // refined from the vertices by the caller (sg_unit_vertices)
        for (int i = 0; i < 3; ++i) { aabb[i] = -4.0f; aabb[3 + i] = 4.0f; }
    } else { // AxisAlignedBoundingBox::unit()
        for (int i = 0; i < 3; ++i) { aabb[i] = -0.5f; aabb[3 + i] = 0.5f; }
    }
}

// affine inverse in double of a column-major 4x4 (float in, float out), bottom row exactly (0,0,0,1)
void affine_inverse(const double m[16], float out[16])
{
    const double a = m[0], b = m[4], c = m[8], d = m[1], e = m[5], f = m[9], g = m[2], h = m[6], i = m[10];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double id = 1.0 / det;
    double r[9];
    r[0] = A * id; r[1] = B * id; r[2] = C * id;                                       // column 0 of inverse (rows 0..2)
    r[3] = -(b * i - c * h) * id; r[4] = (a * i - c * g) * id; r[5] = -(a * h - b * g) * id; // column 1
    r[6] = (b * f - c * e) * id; r[7] = -(a * f - c * d) * id; r[8] = (a * e - b * d) * id;  // column 2
    const double tx = m[12], ty = m[13], tz = m[14];
    out[0] = (float)r[0]; out[1] = (float)r[1]; out[2] = (float)r[2]; out[3] = 0.0f;
    out[4] = (float)r[3]; out[5] = (float)r[4]; out[6] = (float)r[5]; out[7] = 0.0f;
    out[8] = (float)r[6]; out[9] = (float)r[7]; out[10] = (float)r[8]; out[11] = 0.0f;
    out[12] = (float)-(r[0] * tx + r[3] * ty + r[6] * tz);
    out[13] = (float)-(r[1] * tx + r[4] * ty + r[7] * tz);
    out[14] = (float)-(r[2] * tx + r[5] * ty + r[8] * tz);
    out[15] = 1.0f;
}

void mul_d(const double a[16], const float b[16], double out[16])
{
    double r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            double y = 0;
            for (int k = 0; k < 4; ++k) y += a[k * 4 + i] * (double)b[j * 4 + k];
            r[j * 4 + i] = y;
        }
    memcpy(out, r, sizeof r);
}

} // namespace

extern "C" sg_scene *sg_create(const sg_config *cfg)
{
    if (!cfg || cfg->n_nodes == 0 || cfg->bones_per_unit == 0 || cfg->bones_per_unit > 255) return nullptr;
    const uint32_t B = cfg->bones_per_unit;
    const uint64_t unit_nodes = (uint64_t)cfg->n_units * (B + 1);
    if (unit_nodes + 1 > cfg->n_nodes) return nullptr;
    sg_scene *sc = new sg_scene();
    sc->cfg = *cfg;
    if (sc->cfg.nranks < 1) { sc->cfg.nranks = 1; sc->cfg.rank = 0; }
    const uint32_t R = (uint32_t)sc->cfg.nranks, rank = (uint32_t)sc->cfg.rank;
    const uint32_t n_static = cfg->n_nodes - (uint32_t)unit_nodes;
    uint32_t S = (uint32_t)std::floor(std::cbrt((double)n_static));
    while (S > 0 && 1ull + S + (uint64_t)S * S > n_static) --S;
    // every rank must own at least one sector
    if (S < R) S = (1ull + R + (uint64_t)R * R <= n_static) ? R : 0;
    const uint32_t G = S * S;
    sc->S = S; sc->G = G;
    sc->first_group = 1 + S;
    sc->first_leaf = 1 + S + G;
    sc->n_leaves = n_static - 1 - S - G;
    sc->first_unit = sc->first_leaf + sc->n_leaves;
    sc->unit_stride = B + 1;

    auto owner_of_group = [&](uint32_t g) { return (g / S) % R; };
    auto owns_group_slot = [&](uint32_t x /* leaf or unit ordinal */) {
        if (G == 0) return (x % R) == rank;
        return owner_of_group(x % G) == rank;
    };

    // pass 1: local index assignment
    std::vector<uint32_t> sector_local(S, NONE), group_local(G, NONE);
    uint32_t n = 1; // root
    for (uint32_t s = 0; s < S; ++s)
        if (s % R == rank) sector_local[s] = n++;
    for (uint32_t g = 0; g < G; ++g)
        if (owner_of_group(g) == rank) group_local[g] = n++;
    std::vector<uint32_t> leaves;
    leaves.reserve(sc->n_leaves / R + 16);
    for (uint32_t l = 0; l < sc->n_leaves; ++l)
        if (owns_group_slot(l)) leaves.push_back(l);
    const uint32_t first_leaf_local = n;
    n += (uint32_t)leaves.size();
    for (uint32_t u = 0; u < cfg->n_units; ++u)
        if (owns_group_slot(u)) {
            sc->unit_global.push_back(u);
            sc->unit_first.push_back(n);
            n += B + 1;
        }
    const uint32_t cap = n;
    sc->parent.assign(cap, NONE);
    sc->flags.assign(cap, 0);
    sc->mask.assign(cap, 0);
    sc->gidx.assign(cap, 0);
    sc->local.assign((size_t)cap * 16, 0.f);
    sc->aabb.assign((size_t)cap * 6, 0.f);

    auto emit = [&](uint32_t li, uint32_t gid, Kind kind, uint32_t parent_local, Trs *keep) {
        Trs t;
        uint32_t f, m;
        gen_node(*sc, gid, kind, t, f, m, &sc->aabb[(size_t)li * 6]);
        trs_to_m16(t, &sc->local[(size_t)li * 16]);
        sc->parent[li] = parent_local;
        sc->flags[li] = f;
        sc->mask[li] = m;
        sc->gidx[li] = gid;
        if (keep) *keep = t;
    };
    emit(0, 0, K_ROOT, NONE, nullptr);
    for (uint32_t s = 0; s < S; ++s)
        if (sector_local[s] != NONE) emit(sector_local[s], 1 + s, K_SECTOR, 0, nullptr);
    for (uint32_t g = 0; g < G; ++g)
        if (group_local[g] != NONE) emit(group_local[g], sc->first_group + g, K_GROUP, sector_local[g / S], nullptr);
    const int64_t nl = (int64_t)leaves.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nl; ++i) {
        const uint32_t l = leaves[i];
        emit(first_leaf_local + (uint32_t)i, sc->first_leaf + l, K_LEAF, G ? group_local[l % G] : 0u, nullptr);
    }
    const int64_t nu = (int64_t)sc->unit_global.size();
    sc->bone_nodes.resize((size_t)nu * B);
    sc->inv_bind.resize((size_t)nu * B * 16);
    sc->bone_rest.resize((size_t)nu * B);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nu; ++i) {
        const uint32_t u = sc->unit_global[i];
        const uint32_t gbase = sc->first_unit + u * sc->unit_stride;
        const uint32_t lbase = sc->unit_first[i];
        const uint32_t group = G ? group_local[u % G] : 0u;
        std::vector<double> chain((size_t)B * 16);
        for (uint32_t k = 0; k < B; ++k) {
            const uint32_t pl = (k == 0) ? group : lbase + bone_parent(k);
            emit(lbase + k, gbase + k, K_BONE, pl, &sc->bone_rest[(size_t)i * B + k]);
            sc->bone_nodes[(size_t)i * B + k] = lbase + k;
            // rest pose in the unit's frame, in double; inverse bind pose = its inverse
            const float *L = &sc->local[(size_t)(lbase + k) * 16];
            double *C = &chain[(size_t)k * 16];
            if (k == 0) {
                for (int e = 0; e < 16; ++e) C[e] = (double)L[e];
            } else {
                mul_d(&chain[(size_t)bone_parent(k) * 16], L, C);
            }
            affine_inverse(C, &sc->inv_bind[((size_t)i * B + k) * 16]);
        }
        emit(lbase + B, gbase + B, K_SKINNED, group, nullptr);
    }
    uint32_t nr = 0;
    for (uint32_t i = 0; i < cap; ++i) nr += (sc->flags[i] & F_REND) ? 1u : 0u;
    sc->n_rend = nr;
    return sc;
}

extern "C" void sg_free(sg_scene *s) { delete s; }
extern "C" void sg_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
extern "C" uint32_t sg_capacity(const sg_scene *s) { return (uint32_t)s->parent.size(); }
extern "C" uint32_t sg_n_renderable(const sg_scene *s) { return s->n_rend; }
extern "C" const uint32_t *sg_parent(const sg_scene *s) { return s->parent.data(); }
extern "C" const uint32_t *sg_flags(const sg_scene *s) { return s->flags.data(); }
extern "C" const uint32_t *sg_render_mask(const sg_scene *s) { return s->mask.data(); }
extern "C" const float *sg_local_m16(const sg_scene *s) { return s->local.data(); }
extern "C" const float *sg_local_aabb(const sg_scene *s) { return s->aabb.data(); }
extern "C" const uint32_t *sg_global_index(const sg_scene *s) { return s->gidx.data(); }
extern "C" uint32_t sg_n_units(const sg_scene *s) { return (uint32_t)s->unit_global.size(); }
extern "C" uint32_t sg_unit_mesh_node(const sg_scene *s, uint32_t u) { return s->unit_first[u] + s->cfg.bones_per_unit; }
extern "C" const uint32_t *sg_unit_bone_nodes(const sg_scene *s, uint32_t u) { return &s->bone_nodes[(size_t)u * s->cfg.bones_per_unit]; }
extern "C" const float *sg_unit_inv_bind(const sg_scene *s, uint32_t u) { return &s->inv_bind[(size_t)u * s->cfg.bones_per_unit * 16]; }

extern "C" void sg_unit_vertices(const sg_scene *s, uint32_t u, void *out_verts, float out_aabb6[6])
{
    const uint32_t B = s->cfg.bones_per_unit, V = s->cfg.verts_per_unit;
    const uint32_t gid = s->first_unit + s->unit_global[u] * s->unit_stride + B; // mesh node's global id
    Pcg32 r(s->cfg.seed ^ 0xA5A5F00DCAFEull, gid);
    unsigned char *o = static_cast<unsigned char *>(out_verts);
    float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (uint32_t v = 0; v < V; ++v, o += 68) {
        float rec[16]; // pos3 uv2 nrm3 tan4 w4
        for (int i = 0; i < 3; ++i) {
            rec[i] = r.range(-4.0f, 4.0f);
            if (rec[i] < mn[i]) mn[i] = rec[i];
            if (rec[i] > mx[i]) mx[i] = rec[i];
        }
        rec[3] = r.uni(); rec[4] = r.uni();
        float nx = r.gauss(), ny = r.gauss(), nz = r.gauss();
        float nl = std::sqrt(nx * nx + ny * ny + nz * nz);
        if (nl < 1e-6f) { nx = 0; ny = 1; nz = 0; nl = 1; }
        rec[5] = nx / nl; rec[6] = ny / nl; rec[7] = nz / nl;
        rec[8] = 1.f; rec[9] = 0.f; rec[10] = 0.f; rec[11] = 1.f;
        // four influences near one bone: itself, its parent, a child, its sibling
        const uint32_t b0 = r.next() % B;
        uint32_t bi[4] = {b0, bone_parent(b0), b0, b0};
        const uint32_t child = (b0 == 0) ? 1u : 2u * b0;
        if (child < B) bi[2] = child;
        if (b0 >= 2) {
            const uint32_t t = b0 - 1;
            const uint32_t sib = ((t & 1u) ? t + 1 : t - 1) + 1;
            if (sib < B) bi[3] = sib;
        }
        float w[4], sum = 0.f;
        for (int k = 0; k < 4; ++k) w[k] = 0.05f + r.uni();
        const float z = r.uni();
        if (z < 0.125f) { w[3] = 0.f; }
        else if (z < 0.25f) { w[2] = 0.f; w[3] = 0.f; }
        for (int k = 0; k < 4; ++k) sum += w[k];
        for (int k = 0; k < 4; ++k) rec[12 + k] = w[k] / sum;
        memcpy(o, rec, 64);
        o[64] = (unsigned char)bi[0]; o[65] = (unsigned char)bi[1]; o[66] = (unsigned char)bi[2]; o[67] = (unsigned char)bi[3];
    }
    if (out_aabb6) {
        for (int i = 0; i < 3; ++i) { out_aabb6[i] = mn[i]; out_aabb6[3 + i] = mx[i]; }
    }
}

extern "C" void sg_units_vertices(const sg_scene *s, uint32_t u0, uint32_t count, void *out_verts, float *out_aabb6)
{
    const size_t per = (size_t)s->cfg.verts_per_unit * 68;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)count; ++i)
        sg_unit_vertices(s, u0 + (uint32_t)i, static_cast<unsigned char *>(out_verts) + per * (size_t)i,
                         out_aabb6 ? out_aabb6 + 6 * (size_t)i : nullptr);
}

static void animate_one(const sg_scene *s, int64_t e, uint32_t frame, Trs &t)
{
    const uint32_t li = s->bone_nodes[e];
    t = s->bone_rest[e];
    // small-angle perturbation about a per-bone axis, phase from the global id
    Pcg32 r(s->cfg.seed ^ 0x51DE5EEDull, s->gidx[li]);
    float ax[3] = {r.gauss(), r.gauss(), r.gauss()};
    float al = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    if (al < 1e-6f) { ax[0] = 1; ax[1] = 0; ax[2] = 0; al = 1; }
    const float phase = 6.2831853f * r.uni();
    const float ang = 0.15f * std::sin(0.37f * (float)(frame + 1) + phase);
    const float sh = std::sin(0.5f * ang) / al, ch = std::cos(0.5f * ang);
    const float d[4] = {ax[0] * sh, ax[1] * sh, ax[2] * sh, ch};
    const float *q = t.q;
    float qn[4] = {
        q[3] * d[0] + q[0] * d[3] + q[1] * d[2] - q[2] * d[1],
        q[3] * d[1] - q[0] * d[2] + q[1] * d[3] + q[2] * d[0],
        q[3] * d[2] + q[0] * d[1] - q[1] * d[0] + q[2] * d[3],
        q[3] * d[3] - q[0] * d[0] - q[1] * d[1] - q[2] * d[2],
    };
    quat_normalize(qn);
    memcpy(t.q, qn, sizeof qn);
}

extern "C" uint32_t sg_animate(const sg_scene *s, uint32_t frame, uint32_t *out_idx, float *out_m16)
{
    const uint32_t B = s->cfg.bones_per_unit;
    const int64_t n = (int64_t)s->unit_global.size() * B;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        Trs t;
        animate_one(s, e, frame, t);
        if (out_idx) out_idx[e] = s->bone_nodes[e];
        trs_to_m16(t, out_m16 + 16 * (size_t)e);
    }
    return (uint32_t)n;
}

extern "C" uint32_t sg_animate_trs(const sg_scene *s, uint32_t frame, uint32_t *out_idx, float *out)
{
    const uint32_t B = s->cfg.bones_per_unit;
    const int64_t n = (int64_t)s->unit_global.size() * B;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        Trs t;
        animate_one(s, e, frame, t);
        if (out_idx) out_idx[e] = s->bone_nodes[e];
        float *o = out + 10 * (size_t)e;
        o[0] = t.t[0]; o[1] = t.t[1]; o[2] = t.t[2];
        o[3] = t.q[0]; o[4] = t.q[1]; o[5] = t.q[2]; o[6] = t.q[3];
        o[7] = t.s[0]; o[8] = t.s[1]; o[9] = t.s[2];
    }
    return (uint32_t)n;
}
