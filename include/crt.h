/* crt.h — C ABI of the B200 ray-intersection backend (libcrt_b200.so).
 *
 * Drop-in for the `crust_rt` crate API of doubleailes/crust-render
 * (crates/crust-rt/src/lib.rs:45-53): Geometry -> SceneBuilder -> commit() ->
 * Scene::{intersect, occluded}, ID-based hits. The reference has no FFI; these
 * are the entry points a `crust-rt-sys` shim would bind (INTEGRATION.md shows
 * the Rust side). Every `file:line` below is relative to
 * /root/reference/crates/crust-rt/src.
 *
 * Conventions
 *  - every call returns a crt_status (0 = ok) and never aborts or throws;
 *    crt_last_error() gives the message for the calling thread;
 *  - there is NO CPU fallback: queries run on the CUDA device or fail;
 *  - `*_n` calls take DEVICE pointers and a CUDA stream (the hot path);
 *    `*_host_n` calls take HOST pointers and do the copies themselves;
 *  - transforms are 12 floats in glam `Affine3A` column order:
 *    x_axis[3], y_axis[3], z_axis[3], translation[3].
 */
#ifndef CRT_H
#define CRT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ray.rs:8-11 */
#define CRT_MASK_CAMERA 0x1u
#define CRT_MASK_SHADOW 0x2u
#define CRT_MASK_INDIRECT 0x4u
#define CRT_MASK_ALL 0xFFFFFFFFu
/* lib.rs:53 */
#define CRT_INVALID_ID 0xFFFFFFFFu

typedef enum crt_status {
    CRT_OK = 0,
    CRT_ERR_INVALID_ARG = 1,
    CRT_ERR_UNKNOWN_GEOMETRY = 2, /* set_geometry on an id never attached (the reference panics, scene.rs:196-200) */
    CRT_ERR_CUDA = 3,
    CRT_ERR_NO_DEVICE = 4,
    CRT_ERR_NESTING = 5, /* instance nesting deeper than CRT_MAX_INSTANCE_DEPTH */
    CRT_ERR_OOM = 6
} crt_status;

#define CRT_MAX_INSTANCE_DEPTH 8

typedef struct crt_builder crt_builder; /* scene.rs:147 SceneBuilder */
typedef struct crt_scene crt_scene;     /* scene.rs:345 Scene (ref-counted like Arc<Scene>) */

/* ray.rs:17-22 Ray + the (t_min, t_max) the reference passes per query
 * (scene.rs:354). `time` and `mask` travel as optional parallel arrays. 32 B. */
typedef struct crt_ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} crt_ray;

/* scene.rs:134-142 RayHit. `front_face` travels as an optional parallel array.
 * Miss: t = +inf, geom_id = prim_id = CRT_INVALID_ID, normal = 0. 32 B. */
typedef struct crt_hit {
    float t, u, v;
    uint32_t geom_id, prim_id;
    float nx, ny, nz;
} crt_hit;

/* scene.rs:18-23 CurveSegment */
typedef struct crt_curve_segment {
    float p0[3], p1[3], r0, r1;
} crt_curve_segment;
/* scene.rs:76-80 CubicCurveSegment */
typedef struct crt_cubic_curve_segment {
    float cp[4][3], r0, r1;
} crt_cubic_curve_segment;

/* scene.rs:62-68 PrimitiveBreakdown */
typedef struct crt_primitive_breakdown {
    uint64_t triangles, spheres, curve_segments, cubic_curve_spans, instances;
} crt_primitive_breakdown;

/* scene.rs:34-46 MemoryFootprint, restated for this backend's structures
 * (bytes resident on the device; shared instanced scenes counted once). */
typedef struct crt_memory_footprint {
    uint64_t bvh_nodes;    /* 128-byte BVH4 nodes */
    uint64_t leaf_records; /* self-contained leaf primitive records */
    uint64_t prim_table;   /* (geom_id, prim_id) per primitive + vertex normals + motion transforms */
    uint64_t descriptors;  /* per-scene device descriptors */
} crt_memory_footprint;

/* Build statistics of one committed scene (this level only). */
typedef struct crt_build_info {
    uint64_t primitives, binary_nodes, wide_nodes, leaves, leaf_refs;
    uint64_t node_bytes, leaf_bytes;
    double build_seconds, upload_seconds;
    uint32_t instance_depth; /* 0 = no instances below */
    uint32_t built_on_device; /* 1 = CRT_COMMIT_DEVICE_BUILD took effect (build_seconds is device time) */
} crt_build_info;

const char* crt_last_error(void);
/* Bind the calling process to a CUDA device (default 0). Fails with
 * CRT_ERR_NO_DEVICE when there is none: this library has no host path. */
int crt_init(int device);
int crt_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem, uint64_t* l2_bytes);
/* L2 persistence of the tree's top levels: bytes of L2 set aside at crt_init (CRT_L2_PERSIST_MB, default 16; 0 = off) and
 * how many traversal launches had the access-policy window over their scene's leading nodes attached / refused. */
int crt_l2_persist_info(uint64_t* set_aside_bytes, uint64_t* windows_applied, uint64_t* windows_refused);

/* ---- SceneBuilder (scene.rs:151-342) -------------------------------------- */
int crt_builder_new(crt_builder** out);                                   /* SceneBuilder::new */
void crt_builder_free(crt_builder* b);                                    /* drop without commit */
int crt_builder_reserve(crt_builder* b, size_t additional);               /* scene.rs:174 reserve */
int crt_builder_count(const crt_builder* b, size_t* out);                 /* scene.rs:179 count */
/* scene.rs:158-166 attach / attach_masked; one call per Geometry variant (scene.rs:87-122).
 * Inputs are copied. `*geom_id` receives the dense id (attach order). */
int crt_attach_triangle_mesh(crt_builder* b, const float* vertices, size_t n_vertices, const uint32_t* indices,
                             size_t n_triangles, const float* normals /* or NULL */, size_t n_normals, uint32_t mask,
                             uint32_t* geom_id);
int crt_attach_sphere(crt_builder* b, const float center[3], float radius, uint32_t mask, uint32_t* geom_id);
int crt_attach_round_curves(crt_builder* b, const crt_curve_segment* segments, size_t n, uint32_t mask,
                            uint32_t* geom_id);
int crt_attach_cubic_curves(crt_builder* b, const crt_cubic_curve_segment* segments, size_t n, uint32_t mask,
                            uint32_t* geom_id);
/* Retains `scene` (Arc clone). `l2w_end` NULL = no transform motion blur. */
int crt_attach_instance(crt_builder* b, crt_scene* scene, const float l2w[12], const float* l2w_end /* or NULL */,
                        uint32_t mask, uint32_t* geom_id);
/* scene.rs:205-211 empty_geometry placeholder */
int crt_attach_empty(crt_builder* b, uint32_t mask, uint32_t* geom_id);
/* scene.rs:199-201 set_geometry: replace geometry `id`, keeping its mask. */
int crt_set_triangle_mesh(crt_builder* b, uint32_t id, const float* vertices, size_t n_vertices,
                          const uint32_t* indices, size_t n_triangles, const float* normals, size_t n_normals);
int crt_set_sphere(crt_builder* b, uint32_t id, const float center[3], float radius);
int crt_set_round_curves(crt_builder* b, uint32_t id, const crt_curve_segment* segments, size_t n);
int crt_set_cubic_curves(crt_builder* b, uint32_t id, const crt_cubic_curve_segment* segments, size_t n);
int crt_set_instance(crt_builder* b, uint32_t id, crt_scene* scene, const float l2w[12], const float* l2w_end);
int crt_set_empty(crt_builder* b, uint32_t id);

/* scene.rs:226-341 commit(self): expands geometries, builds the deterministic
 * SAH/SBVH BVH4 on the host (bvh.rs:300-327, 1059-1375), flattens it to the
 * device layout and uploads. Consumes the builder (also on failure). */
int crt_commit(crt_builder* b, crt_scene** out);
#define CRT_COMMIT_KEEP_HOST_TREE 0x1u /* keep the host copy for crt_debug_* */
#define CRT_COMMIT_HOST_ONLY 0x2u      /* test support: build the host tree and skip the upload; every query on such
                                          a scene fails with CRT_ERR_NO_DEVICE (there is no host traversal) */
/* Build the BVH on the device instead of the host: Morton-order radix tree collapsed to BVH4 with
 * the reference's largest-area rule, leaves of <= 4 triangles. For scenes made only of triangle meshes without
 * vertex normals; any other scene silently takes the host build (crt_build_info.built_on_device tells which).
 * Milliseconds instead of seconds at 10 M triangles. The tree is NOT the reference's SAH/SBVH tree: every query
 * returns the same closest hit except where two primitives tie at exactly equal t (the reference's "later wins"
 * rule, bvh.rs:537-544, depends on visiting order), and traversal is somewhat slower. Deterministic. */
#define CRT_COMMIT_DEVICE_BUILD 0x4u
int crt_commit_ex(crt_builder* b, uint32_t flags, int host_threads /* 0 = all */, crt_scene** out);

void crt_scene_retain(crt_scene* s);
void crt_scene_release(crt_scene* s);

/* ---- Scene (scene.rs:351-472) --------------------------------------------- */
int crt_bounds(const crt_scene* s, float out_min_max[6], int* non_empty);            /* scene.rs:375 */
int crt_geometry_count(const crt_scene* s, uint32_t* out);                           /* scene.rs:380 */
int crt_has_motion(const crt_scene* s, int* out);                                    /* scene.rs:395 */
int crt_primitive_count(const crt_scene* s, uint64_t* out);                          /* scene.rs:400 */
int crt_primitive_breakdown_of(const crt_scene* s, crt_primitive_breakdown* out);    /* scene.rs:409 */
int crt_unique_primitive_breakdown(const crt_scene* s, crt_primitive_breakdown* out); /* scene.rs:422 */
/* scene.rs:446: (count, scene diagonal, mean prim diagonal, max prim diagonal) */
int crt_primitive_extents(const crt_scene* s, uint64_t* count, float* scene_diag, float* mean_diag, float* max_diag);
int crt_memory_footprint_of(const crt_scene* s, crt_memory_footprint* out);          /* scene.rs:459 */
int crt_build_info_of(const crt_scene* s, crt_build_info* out);

/* ---- queries --------------------------------------------------------------
 * scene.rs:354 intersect / :370 occluded, batched. DEVICE pointers; work is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call
 * returns without synchronising. `time`, `mask`, `front_face` may be NULL
 * (time 0, CRT_MASK_ALL, not written).
 * Occlusion answers do not depend on the order in which the tree is walked, but their cost does: a ray with a finite
 * tmax (a segment towards a light) has the children of a node visited farthest first, an open-ended one nearest first
 * (DESIGN.md §3.1). Give shadow rays their real far end rather than INFINITY. */
int crt_intersect_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                    crt_hit* hits, uint8_t* front_face, void* stream);
int crt_occluded_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                   uint8_t* occluded, void* stream);
/* Same, HOST pointers: chunks are staged through pinned buffers on internal
 * streams so H2D, traversal and D2H overlap. Synchronous. */
int crt_intersect_host_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                         crt_hit* hits, uint8_t* front_face);
int crt_occluded_host_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                        uint8_t* occluded);
/* One ray, reference call shape (Scene::intersect / Scene::occluded, scene.rs:354-372). Runs the
 * same device kernel on a batch of one; `*hit` = 1 when something was hit.
 * This is the reference's SEMANTICS on one ray, for tests and spot checks — not its calling pattern: the reference's
 * queries are `&self` and are called once per ray from every rayon worker concurrently, whereas each call here is one
 * kernel launch plus two copies (tens of microseconds) and calls from several host threads are serialised by the
 * host-buffer pipeline's lock. A host integrator must batch its rays and use crt_*_n / crt_*_host_n (any number of host
 * threads may call those concurrently; the _host_n entries share one staging pipeline), or run on the device through
 * crt_render (include/crt_render.h). */
int crt_intersect1(const crt_scene* s, const crt_ray* ray, float time, uint32_t mask, crt_hit* out, int* front_face,
                   int* hit);
int crt_occluded1(const crt_scene* s, const crt_ray* ray, float time, uint32_t mask, int* occluded);

/* Number of kernel launches this library has issued in this process. */
uint64_t crt_launch_count(void);
/* Sticky device-side error bits; bit 0 = a traversal stack overflowed (an
 * entry was dropped, results of that launch are not trustworthy).
 * Synchronises the device. */
/* Replicating a commit without rebuilding it (multi-GPU: the reference builds once per process, scene.rs:226; with one
 * process per GPU the same tree would be built N times). crt_scene_export writes one committed scene — its flattened
 * device arrays and host-side metadata — as a flat blob into host (dst_on_device = 0) or device memory; crt_scene_import
 * turns such a blob into a committed scene on the calling process' device. Scenes instanced by the exported one are not
 * embedded: the blob names them by their index among the scene's distinct children in first-attachment order, and the
 * importer passes the handles it obtained for them (import or commit) in that order. The imported scene retains them like
 * an instance attach does. A device blob can be broadcast with NCCL and is unpacked with device-to-device copies. */
int crt_scene_export_size(const crt_scene* s, size_t* bytes);
int crt_scene_export(const crt_scene* s, void* dst, size_t bytes, int dst_on_device, void* stream);
int crt_scene_import(const void* src, size_t bytes, int src_on_device, crt_scene* const* children, size_t n_children, crt_scene** out);

/* Device-side traversal counters (the reference's `stats` module, bvh.rs:31-136, counts on the CPU): while enabled, every
 * traversal launch runs the counting build of its kernel (slower; never enable it around a timed region).
 * crt_trace_counters_read synchronises the device; out = {rays, BVH4 nodes fetched, popped entries culled by their entry
 * distance without a fetch, leaves entered, triangle records tested, other primitive records tested, instance descents}. */
int crt_trace_counters_enable(int on);
int crt_trace_counters_read(uint64_t out[7], int reset);

int crt_device_status(uint32_t* bits);
/* Clears the sticky bits (synchronises the device). */
int crt_device_status_reset(void);

/* ---- debug / test support (needs CRT_COMMIT_KEEP_HOST_TREE) ---------------
 * Canonical view of the host tree for builder-equivalence tests:
 * sizes = {wide nodes, leaves, leaf refs, binary nodes}. */
int crt_debug_tree_sizes(const crt_scene* s, uint64_t out4[4]);
/* nodes: 32 words per wide node in bvh.rs:196-215 order (bounds x6, child[4],
 * flags, 3 zero words) with leaf lanes numbered in creation order; per leaf the
 * count of triangle and other references; refs: primitive indices leaf by leaf. */
int crt_debug_tree_export(const crt_scene* s, void* nodes, uint32_t* tri_counts, uint32_t* other_counts,
                          uint32_t* refs);
/* The flattened host arrays as 32-bit words, for tests that compare two commits of the same scene word for word (instance
 * records hold zeros where the upload patches device pointers in): which = 0 wide nodes (32 words each), 1 leaf records
 * (4 words per 16-byte unit, padding included), 2 motion pairs (24 words each), 3 instance record offsets. Returns the
 * number of words through *n_words; copies min(capacity_words, *n_words) words when `out` is not NULL. */
int crt_debug_host_array(const crt_scene* s, int which, uint32_t* out, uint64_t capacity_words, uint64_t* n_words);

#ifdef __cplusplus
}
#endif
#endif /* CRT_H */
