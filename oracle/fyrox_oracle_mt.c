/*
 * fyrox_oracle_mt.c — "soa-omp-NT": the same arithmetic as fyrox_oracle.c on flat arrays, parallelised with OpenMP
 * over hierarchy levels / nodes / skinned surfaces.
 *
 * TEST INFRASTRUCTURE ONLY (see fyrox_oracle.h).  This is NOT how the reference runs — Fyrox's path is a
 * single-threaded recursive walk over a pool of heap nodes, which fyrox_oracle.c restates — it is
 * the best-effort multi-core CPU baseline BASELINE.md §3 asks to be timed next to it, labelled as such wherever it
 * is reported.  Every arithmetic step calls the functions of fyrox_oracle.c (orc_mat4_mul, orc_aabb_transform,
 * orc_aabb_add_point, orc_frustum_is_intersects_aabb, orc_skin_vertices), so its results are bit-identical to the
 * single-threaded restatement (tests/test_oracle_kat.py::test_mt_oracle_equals_the_single_threaded_one) for
 * scenes whose skeletons precede their meshes in DFS order (all generated scenes).
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "fyrox_oracle.h"

typedef struct {
    uint32_t mesh, n_bones, n_verts;
    uint32_t *bones;
    unsigned char *verts; /* owned, interleaved */
    orc_vertex_layout layout;
} mt_surface;

struct orc_mt {
    uint32_t n, root, n_levels;
    uint32_t *parent, *flags, *mask;
    uint32_t *order, *level_off; /* nodes sorted by depth */
    float *L, *G, *IB;           /* 16 floats per node */
    orc_aabb *la, *wa;
    uint8_t *gvis, *gen, *reach;
    mt_surface *surf;
    uint32_t n_surf, cap_surf;
    uint8_t *skinned;            /* per node: has a surface with bones */
    int threads;
};

void orc_mt_set_threads(orc_mt *m, int t) { m->threads = t > 0 ? t : 1; }

orc_mt *orc_mt_new(uint32_t n, uint32_t root, const uint32_t *parent, const uint32_t *flags, const uint32_t *mask,
                   const float *local_m16, const float *local_aabb6)
{
    orc_mt *m = (orc_mt *)calloc(1, sizeof *m);
    m->n = n;
    m->root = root;
    m->threads = 1;
#ifdef _OPENMP
    m->threads = omp_get_max_threads();
#endif
    size_t n1 = n ? n : 1;
    m->parent = (uint32_t *)malloc(4 * n1);
    m->flags = (uint32_t *)malloc(4 * n1);
    m->mask = (uint32_t *)malloc(4 * n1);
    m->order = (uint32_t *)malloc(4 * n1);
    m->L = (float *)malloc(64 * n1);
    m->G = (float *)malloc(64 * n1);
    m->IB = (float *)malloc(64 * n1);
    m->la = (orc_aabb *)malloc(sizeof(orc_aabb) * n1);
    m->wa = (orc_aabb *)malloc(sizeof(orc_aabb) * n1);
    m->gvis = (uint8_t *)calloc(n1, 1);
    m->gen = (uint8_t *)calloc(n1, 1);
    m->reach = (uint8_t *)calloc(n1, 1);
    m->skinned = (uint8_t *)calloc(n1, 1);
    memcpy(m->parent, parent, 4 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i) {
        m->flags[i] = flags ? flags[i] : (ORC_FLAG_VISIBILITY | ORC_FLAG_ENABLED | ORC_FLAG_FRUSTUM_CULLING | ORC_FLAG_CAST_SHADOWS | ORC_FLAG_ALIVE);
        m->mask[i] = mask ? mask[i] : 0xFFFFFFFFu;
        if (local_m16) memcpy(m->L + 16 * (size_t)i, local_m16 + 16 * (size_t)i, 64);
        else orc_mat4_identity(m->L + 16 * (size_t)i);
        orc_mat4_identity(m->G + 16 * (size_t)i);
        orc_mat4_identity(m->IB + 16 * (size_t)i);
        if (local_aabb6) {
            memcpy(m->la[i].min, local_aabb6 + 6 * (size_t)i, 12);
            memcpy(m->la[i].max, local_aabb6 + 6 * (size_t)i + 3, 12);
        } else {
            orc_aabb_unit(&m->la[i]);
        }
        orc_aabb_default(&m->wa[i]);
    }
    /* depth of every alive node (parents may have larger indices), then a counting sort by depth */
    int32_t *depth = (int32_t *)malloc(sizeof(int32_t) * n1);
    for (uint32_t i = 0; i < n; ++i) depth[i] = -1;
    uint32_t *stack = (uint32_t *)malloc(4 * n1);
    uint32_t maxd = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (!(m->flags[i] & ORC_FLAG_ALIVE) || depth[i] >= 0) continue;
        uint32_t sp = 0, x = i;
        int32_t base = -1;
        for (;;) {
            if (depth[x] >= 0) { base = depth[x]; break; }
            stack[sp++] = x;
            uint32_t p = m->parent[x];
            if (p == ORC_NONE || p >= n || !(m->flags[p] & ORC_FLAG_ALIVE)) break;
            x = p;
        }
        while (sp) {
            depth[stack[--sp]] = ++base;
            if ((uint32_t)base > maxd) maxd = (uint32_t)base;
        }
    }
    m->n_levels = n ? maxd + 1 : 0;
    m->level_off = (uint32_t *)calloc((size_t)m->n_levels + 2, 4);
    for (uint32_t i = 0; i < n; ++i)
        if (depth[i] >= 0) m->level_off[depth[i] + 1]++;
    for (uint32_t l = 0; l < m->n_levels; ++l) m->level_off[l + 1] += m->level_off[l];
    uint32_t *cur = (uint32_t *)malloc(4 * ((size_t)m->n_levels + 1));
    memcpy(cur, m->level_off, 4 * ((size_t)m->n_levels + 1));
    for (uint32_t i = 0; i < n; ++i)
        if (depth[i] >= 0) m->order[cur[depth[i]]++] = i;
    free(cur);
    free(stack);
    free(depth);
    return m;
}

void orc_mt_free(orc_mt *m)
{
    if (!m) return;
    for (uint32_t s = 0; s < m->n_surf; ++s) {
        free(m->surf[s].bones);
        free(m->surf[s].verts);
    }
    free(m->surf);
    free(m->parent); free(m->flags); free(m->mask); free(m->order); free(m->level_off);
    free(m->L); free(m->G); free(m->IB); free(m->la); free(m->wa);
    free(m->gvis); free(m->gen); free(m->reach); free(m->skinned);
    free(m);
}

void orc_mt_set_local_matrices(orc_mt *m, uint32_t count, const uint32_t *idx, const float *m16)
{
#pragma omp parallel for num_threads(m->threads) schedule(static)
    for (uint32_t e = 0; e < count; ++e) {
        uint32_t i = idx ? idx[e] : e;
        if (i < m->n) memcpy(m->L + 16 * (size_t)i, m16 + 16 * (size_t)e, 64);
    }
}

void orc_mt_set_inv_bind(orc_mt *m, uint32_t node, const float m16[16])
{
    if (node < m->n) memcpy(m->IB + 16 * (size_t)node, m16, 64);
}

uint32_t orc_mt_add_surface(orc_mt *m, uint32_t mesh, uint32_t n_bones, const uint32_t *bones, uint32_t n_verts,
                            const void *verts, const orc_vertex_layout *layout)
{
    if (m->n_surf == m->cap_surf) {
        m->cap_surf = m->cap_surf ? m->cap_surf * 2 : 64;
        m->surf = (mt_surface *)realloc(m->surf, sizeof(mt_surface) * m->cap_surf);
    }
    mt_surface *s = &m->surf[m->n_surf];
    s->mesh = mesh;
    s->n_bones = n_bones;
    s->n_verts = n_verts;
    s->bones = (uint32_t *)malloc(4 * (size_t)(n_bones ? n_bones : 1));
    memcpy(s->bones, bones, 4 * (size_t)n_bones);
    s->layout = *layout;
    s->verts = (unsigned char *)malloc((size_t)n_verts * layout->stride + 1);
    if (n_verts) memcpy(s->verts, verts, (size_t)n_verts * layout->stride);
    if (n_bones && mesh < m->n) m->skinned[mesh] = 1;
    /* Mesh::local_bounding_box = bounds of the vertex positions (orc_mesh_recalc_local_aabb; one surface per mesh here) */
    if (n_verts && mesh < m->n) {
        orc_aabb bb;
        orc_aabb_default(&bb);
        for (uint32_t v = 0; v < n_verts; ++v) {
            float p[3];
            memcpy(p, s->verts + (size_t)v * layout->stride + layout->position_offset, 12);
            orc_aabb_add_point(&bb, p);
        }
        m->la[mesh] = bb;
    }
    return m->n_surf++;
}

/* Graph::update_hierarchical_data (all nodes): level by level — parents of level l are all in level l-1.
 * Per node exactly what update_global_transform_recursively / update_visibility_recursively /
 * update_enabled_flag_recursively compute (scene/graph/mod.rs:1166-1241); world boxes for renderable nodes
 * (Mesh::on_global_transform_changed, scene/mesh/mod.rs:667-689), skinned meshes in a second pass over the bones. */
void orc_mt_update(orc_mt *m)
{
    for (uint32_t l = 0; l < m->n_levels; ++l) {
        const uint32_t lo = m->level_off[l], hi = m->level_off[l + 1];
#pragma omp parallel for num_threads(m->threads) schedule(static)
        for (uint32_t k = lo; k < hi; ++k) {
            const uint32_t i = m->order[k];
            const uint32_t p = m->parent[i];
            const int has_parent = (l > 0);
            float pg[16];
            if (has_parent) memcpy(pg, m->G + 16 * (size_t)p, 64);
            else orc_mat4_identity(pg);
            orc_mat4_mul(pg, m->L + 16 * (size_t)i, m->G + 16 * (size_t)i);
            const uint32_t f = m->flags[i];
            m->gvis[i] = (uint8_t)((has_parent ? m->gvis[p] : 1) && (f & ORC_FLAG_VISIBILITY));
            m->gen[i] = (uint8_t)((has_parent ? m->gen[p] : 1) && (f & ORC_FLAG_ENABLED));
            m->reach[i] = (uint8_t)(has_parent ? m->reach[p] : (i == m->root));
            if (f & ORC_FLAG_RENDERABLE) orc_aabb_transform(&m->la[i], m->G + 16 * (size_t)i, &m->wa[i]);
        }
    }
    /* skinned meshes: add_point(bone.global_position()) in surface / bone order */
#pragma omp parallel for num_threads(m->threads) schedule(dynamic, 64)
    for (uint32_t s = 0; s < m->n_surf; ++s) {
        const mt_surface *sf = &m->surf[s];
        if (!sf->n_bones || sf->mesh >= m->n) continue;
        orc_aabb w = m->wa[sf->mesh];
        for (uint32_t b = 0; b < sf->n_bones; ++b) {
            const uint32_t bn = sf->bones[b];
            if (bn < m->n && (m->flags[bn] & ORC_FLAG_ALIVE)) orc_aabb_add_point(&w, m->G + 16 * (size_t)bn + 12);
        }
        m->wa[sf->mesh] = w; /* one skinned surface per mesh in the generated scenes */
    }
}

/* RenderDataBundleStorage::from_graph reduced to the visible set, as orc_from_graph, order = node index */
size_t orc_mt_cull(const orc_mt *m, const orc_frustum *f, uint32_t render_mask, int shadow_pass, uint32_t *out, size_t cap)
{
    const int T = m->threads;
    size_t *counts = (size_t *)calloc((size_t)T + 1, sizeof(size_t));
    uint32_t **bufs = (uint32_t **)calloc((size_t)T, sizeof(uint32_t *));
#pragma omp parallel num_threads(T)
    {
        int t = 0, nt = 1;
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        const uint32_t lo = (uint32_t)((uint64_t)m->n * (uint64_t)t / (uint64_t)nt), hi = (uint32_t)((uint64_t)m->n * (uint64_t)(t + 1) / (uint64_t)nt);
        uint32_t *buf = (uint32_t *)malloc(4 * (size_t)(hi - lo + 1));
        size_t c = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t fl = m->flags[i];
            if ((fl & (ORC_FLAG_ALIVE | ORC_FLAG_RENDERABLE)) != (ORC_FLAG_ALIVE | ORC_FLAG_RENDERABLE)) continue;
            if (!m->reach[i]) continue;
            if ((m->mask[i] & render_mask) == 0) continue;
            if (!m->gvis[i] || !m->gen[i]) continue;
            if ((fl & ORC_FLAG_FRUSTUM_CULLING) && f && !orc_frustum_is_intersects_aabb(f, &m->wa[i])) continue;
            if (shadow_pass && !(fl & ORC_FLAG_CAST_SHADOWS)) continue;
            buf[c++] = i;
        }
        bufs[t] = buf;
        counts[t + 1] = c;
    }
    size_t total = 0;
    for (int t = 0; t < T; ++t) {
        if (bufs[t]) {
            for (size_t k = 0; k < counts[t + 1]; ++k)
                if (total + k < cap) out[total + k] = bufs[t][k];
            free(bufs[t]);
        }
        total += counts[t + 1];
    }
    free(bufs);
    free(counts);
    return total;
}

/* bone palette + LBS of one surface (scene/mesh/mod.rs:781-793, 501-522) */
void orc_mt_skin_surface(const orc_mt *m, uint32_t s, float *out_pos, float *out_nrm)
{
    const mt_surface *sf = &m->surf[s];
    float *pal = (float *)malloc(64 * (size_t)(sf->n_bones ? sf->n_bones : 1));
    for (uint32_t b = 0; b < sf->n_bones; ++b) {
        const uint32_t bn = sf->bones[b];
        if (bn < m->n && (m->flags[bn] & ORC_FLAG_ALIVE)) orc_mat4_mul(m->G + 16 * (size_t)bn, m->IB + 16 * (size_t)bn, pal + 16 * (size_t)b);
        else orc_mat4_identity(pal + 16 * (size_t)b);
    }
    orc_skin_vertices(pal, sf->n_verts, sf->verts, &sf->layout, out_pos, out_nrm);
    free(pal);
}

/* every surface, one thread per surface at a time, output into per-thread scratch (like the single-threaded
 * baseline, which overwrites one buffer: the skinned streams are produced, not kept) */
void orc_mt_skin_all(const orc_mt *m)
{
    uint32_t maxv = 1;
    for (uint32_t s = 0; s < m->n_surf; ++s)
        if (m->surf[s].n_verts > maxv) maxv = m->surf[s].n_verts;
#pragma omp parallel num_threads(m->threads)
    {
        float *pos = (float *)malloc(12 * (size_t)maxv), *nrm = (float *)malloc(12 * (size_t)maxv);
#pragma omp for schedule(dynamic, 4)
        for (uint32_t s = 0; s < m->n_surf; ++s) orc_mt_skin_surface(m, s, pos, nrm);
        free(pos);
        free(nrm);
    }
}

void orc_mt_get(const orc_mt *m, uint32_t i, float g16[16], orc_aabb *wa, uint32_t *gflags)
{
    memcpy(g16, m->G + 16 * (size_t)i, 64);
    *wa = m->wa[i];
    *gflags = (uint32_t)m->gvis[i] | ((uint32_t)m->gen[i] << 1) | ((uint32_t)m->reach[i] << 2);
}
