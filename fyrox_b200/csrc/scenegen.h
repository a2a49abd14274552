/*
 * scenegen.h — deterministic synthetic Fyrox scenes for tests and bench.py.
 *
 * Host-side input generator only: it produces the arrays a Rust shim would read out of Graph's
 * Pool<Node> (parents, flags, render masks, local matrices, local AABBs, skinned surfaces) and the
 * per-frame changed bone matrices.  It is not part of the hot path and not part of the oracle.
 *
 * Layout of a scene with N nodes and U skinned units (node index = creation order = Handle::index):
 *   0                      root pivot (identity), as Graph::new (scene/graph/mod.rs:408-424)
 *   1 .. S                 "sector" pivots under the root
 *   S+1 .. S+S*S           "group" pivots, S per sector
 *   then static leaf meshes, dealt round-robin to the groups
 *   then U units of (bones_per_unit bone pivots + 1 skinned mesh node), dealt round-robin to the groups;
 *        bones form a root bone + a complete binary tree; they precede their mesh node in DFS order.
 * Every per-node value is a pure function of (seed, global node index), so any sharding of the
 * scene over ranks (sector s belongs to rank s % nranks; the root is replicated) sees the same data.
 */
#ifndef FYROX_SCENEGEN_H
#define FYROX_SCENEGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sg_config {
    uint64_t seed;
    uint32_t n_nodes;          /* total nodes of the whole (unsharded) scene, including the root */
    uint32_t n_units;          /* skinned meshes */
    uint32_t bones_per_unit;   /* 64 in BASELINE.json's configs; 1..255 */
    uint32_t verts_per_unit;   /* 5000 in BASELINE.json's configs */
    int32_t  rank, nranks;     /* shard selector; 0,1 = whole scene */
} sg_config;

typedef struct sg_scene sg_scene;

sg_scene *sg_create(const sg_config *cfg);
void      sg_free(sg_scene *s);
/* OpenMP threads used by the generator (launchers such as torchrun export OMP_NUM_THREADS=1) */
void      sg_set_threads(int n);

/* shard-local node arrays (index = local node index; local 0 is the root) */
uint32_t        sg_capacity(const sg_scene *s);
uint32_t        sg_n_renderable(const sg_scene *s);
const uint32_t *sg_parent(const sg_scene *s);        /* local indices, 0xFFFFFFFF for the root */
const uint32_t *sg_flags(const sg_scene *s);         /* FYX_NODE_* input bits */
const uint32_t *sg_render_mask(const sg_scene *s);
const float    *sg_local_m16(const sg_scene *s);     /* capacity*16, column-major */
const float    *sg_local_aabb(const sg_scene *s);    /* capacity*6: min xyz, max xyz */
const uint32_t *sg_global_index(const sg_scene *s);  /* node index in the unsharded scene */

/* skinned units of this shard */
uint32_t        sg_n_units(const sg_scene *s);
uint32_t        sg_unit_mesh_node(const sg_scene *s, uint32_t u);     /* local index */
const uint32_t *sg_unit_bone_nodes(const sg_scene *s, uint32_t u);    /* bones_per_unit local indices */
const float    *sg_unit_inv_bind(const sg_scene *s, uint32_t u);      /* bones_per_unit*16 */
/* writes verts_per_unit AnimatedVertex records (68 B each, scene/mesh/vertex.rs:140-155) and the
 * min/max of their positions (Mesh::local_bounding_box) */
void            sg_unit_vertices(const sg_scene *s, uint32_t u, void *out_verts, float out_aabb6[6]);
/* same for units u0..u0+count-1 back to back (OpenMP); out_aabb6 = count*6 floats or NULL */
void            sg_units_vertices(const sg_scene *s, uint32_t u0, uint32_t count, void *out_verts, float *out_aabb6);

/* per-frame animation: every bone's local rotation is perturbed; writes n_units*bones_per_unit entries
 * (local node index + Transform::matrix()).  Returns the number of entries. */
uint32_t        sg_animate(const sg_scene *s, uint32_t frame, uint32_t *out_idx, float *out_m16);
/* the same poses as position / rotation (i,j,k,w) / scale records (10 f32 each, fyx_trs): what an animation
 * system hands to Transform; the matrix is then Transform::calculate_local_transform's job */
uint32_t        sg_animate_trs(const sg_scene *s, uint32_t frame, uint32_t *out_idx, float *out_trs10);

#ifdef __cplusplus
}
#endif
#endif
