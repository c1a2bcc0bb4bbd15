// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see crt_oracle.hpp header).
// Plain C entry points over the restatement so Python tests (ctypes), smoke()
// and bench.py's CPU arm can drive it. Ray / hit records use the same 32-byte
// layouts as include/crt.h so identical buffers can be fed to both sides.
#include "crt_oracle.hpp"

#include <chrono>
#include <cstdio>

using namespace orc;

namespace {
struct SceneHandle {
    std::shared_ptr<Scene> scene;
};
struct RayRec {  // == crt_ray
    float ox, oy, oz, tmin, dx, dy, dz, tmax;
};
struct HitRec {  // == crt_hit
    float t, u, v;
    uint32_t geom_id, prim_id;
    float nx, ny, nz;
};
Affine affine_from12(const float* m) {  // glam column order: x_axis, y_axis, z_axis, translation
    Affine a;
    a.x_axis = {m[0], m[1], m[2]};
    a.y_axis = {m[3], m[4], m[5]};
    a.z_axis = {m[6], m[7], m[8]};
    a.translation = {m[9], m[10], m[11]};
    return a;
}
Geometry mesh_geom(const float* verts, size_t nverts, const uint32_t* idx, size_t ntris, const float* normals,
                   size_t n_normals) {
    Geometry g;
    g.type = Geometry::Type::TriangleMesh;
    g.vertices.resize(nverts);
    for (size_t i = 0; i < nverts; ++i) g.vertices[i] = {verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]};
    g.indices.assign(idx, idx + 3 * ntris);
    if (normals) {
        g.has_normals = true;
        g.normals.resize(n_normals);
        for (size_t i = 0; i < n_normals; ++i) g.normals[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    }
    return g;
}
Geometry sphere_geom(float cx, float cy, float cz, float r) {
    Geometry g;
    g.type = Geometry::Type::Sphere;
    g.center = {cx, cy, cz};
    g.radius = r;
    return g;
}
Geometry curves_geom(const float* segs, size_t n) {
    Geometry g;
    g.type = Geometry::Type::RoundCurves;
    g.segments.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const float* s = segs + 8 * i;
        g.segments[i] = {{s[0], s[1], s[2]}, {s[3], s[4], s[5]}, s[6], s[7]};
    }
    return g;
}
Geometry cubic_geom(const float* segs, size_t n) {
    Geometry g;
    g.type = Geometry::Type::CubicCurves;
    g.cubic.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const float* s = segs + 14 * i;
        for (int k = 0; k < 4; ++k) g.cubic[i].cp[k] = {s[3 * k], s[3 * k + 1], s[3 * k + 2]};
        g.cubic[i].r0 = s[12];
        g.cubic[i].r1 = s[13];
    }
    return g;
}
Geometry instance_geom(SceneHandle* child, const float* l2w, const float* l2w_end) {
    Geometry g;
    g.type = Geometry::Type::Instance;
    g.scene = child->scene;
    g.transform = affine_from12(l2w);
    if (l2w_end) {
        g.has_end = true;
        g.transform_end = affine_from12(l2w_end);
    }
    return g;
}
int resolve_threads(int threads) {
    if (threads > 0) return threads;
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}
template <class F>
void parallel_ranges(size_t n, int threads, F&& f) {
    threads = resolve_threads(threads);
    if (threads <= 1 || n < 1024) {
        f(0, n, 0);
        return;
    }
    // Contiguous slices, one per thread: the reference's rayon-over-pixels model
    // (crust-core/src/tracer.rs:428-489).
    std::vector<std::thread> th;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t a = std::min(n, chunk * t), b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=, &f] { f(a, b, t); });
    }
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

void* orc_builder_new() { return new SceneBuilder(); }
void orc_builder_free(void* b) { delete (SceneBuilder*)b; }
uint64_t orc_builder_count(void* b) { return ((SceneBuilder*)b)->geoms.size(); }

uint32_t orc_attach_triangle_mesh(void* b, const float* verts, size_t nverts, const uint32_t* idx, size_t ntris,
                                  const float* normals, size_t n_normals, uint32_t mask) {
    return ((SceneBuilder*)b)->attach_masked(mesh_geom(verts, nverts, idx, ntris, normals, n_normals), mask);
}
uint32_t orc_attach_sphere(void* b, float cx, float cy, float cz, float r, uint32_t mask) {
    return ((SceneBuilder*)b)->attach_masked(sphere_geom(cx, cy, cz, r), mask);
}
uint32_t orc_attach_round_curves(void* b, const float* segs, size_t n, uint32_t mask) {
    return ((SceneBuilder*)b)->attach_masked(curves_geom(segs, n), mask);
}
uint32_t orc_attach_cubic_curves(void* b, const float* segs, size_t n, uint32_t mask) {
    return ((SceneBuilder*)b)->attach_masked(cubic_geom(segs, n), mask);
}
uint32_t orc_attach_instance(void* b, void* child, const float* l2w, const float* l2w_end, uint32_t mask) {
    return ((SceneBuilder*)b)->attach_masked(instance_geom((SceneHandle*)child, l2w, l2w_end), mask);
}
uint32_t orc_attach_empty(void* b, uint32_t mask) { return ((SceneBuilder*)b)->attach_masked(Geometry{}, mask); }
int orc_set_triangle_mesh(void* b, uint32_t id, const float* verts, size_t nverts, const uint32_t* idx, size_t ntris,
                          const float* normals, size_t n_normals) {
    return ((SceneBuilder*)b)->set_geometry(id, mesh_geom(verts, nverts, idx, ntris, normals, n_normals)) ? 0 : 1;
}
int orc_set_sphere(void* b, uint32_t id, float cx, float cy, float cz, float r) {
    return ((SceneBuilder*)b)->set_geometry(id, sphere_geom(cx, cy, cz, r)) ? 0 : 1;
}
int orc_set_instance(void* b, uint32_t id, void* child, const float* l2w, const float* l2w_end) {
    return ((SceneBuilder*)b)->set_geometry(id, instance_geom((SceneHandle*)child, l2w, l2w_end)) ? 0 : 1;
}

// Consumes the builder (scene.rs:226 `commit(self)`).
void* orc_commit(void* b, int threads) {
    SceneBuilder* sb = (SceneBuilder*)b;
    auto* h = new SceneHandle{sb->commit(resolve_threads(threads))};
    delete sb;
    return h;
}
void orc_scene_release(void* s) { delete (SceneHandle*)s; }

uint32_t orc_geometry_count(void* s) { return ((SceneHandle*)s)->scene->n_geoms; }
uint64_t orc_primitive_count(void* s) { return ((SceneHandle*)s)->scene->bvh.prims.size(); }
int orc_has_motion(void* s) { return ((SceneHandle*)s)->scene->has_motion ? 1 : 0; }
int orc_bounds(void* s, float* out6) {
    const auto& rb = ((SceneHandle*)s)->scene->bvh.root_bbox;
    if (!rb) return 0;
    out6[0] = rb->minimum.x;
    out6[1] = rb->minimum.y;
    out6[2] = rb->minimum.z;
    out6[3] = rb->maximum.x;
    out6[4] = rb->maximum.y;
    out6[5] = rb->maximum.z;
    return 1;
}
// triangles, spheres, curve_segments, cubic_curve_spans, instances (bvh.rs:347-359)
void orc_primitive_breakdown(void* s, uint64_t* out5) {
    for (int i = 0; i < 5; ++i) out5[i] = 0;
    for (const PrimNode& p : ((SceneHandle*)s)->scene->bvh.prims) out5[(int)p.kind]++;
}
static void unique_into(const Scene& sc, std::unordered_set<const Scene*>& visited, uint64_t* acc) {  // bvh.rs:397-416
    for (const PrimNode& p : sc.bvh.prims) {
        acc[(int)p.kind]++;
        if (p.kind == Kind::Instance && visited.insert(p.inst->scene.get()).second) unique_into(*p.inst->scene, visited, acc);
    }
}
void orc_unique_primitive_breakdown(void* s, uint64_t* out5) {
    for (int i = 0; i < 5; ++i) out5[i] = 0;
    std::unordered_set<const Scene*> visited;
    unique_into(*((SceneHandle*)s)->scene, visited, out5);
}

// ---- tree export (for builder-equivalence tests) ---------------------------
// sizes: wide nodes, leaves, packets, scalar indices, leaf references, binary nodes
void orc_tree_sizes(void* s, uint64_t* out6) {
    const Bvh& b = ((SceneHandle*)s)->scene->bvh;
    out6[0] = b.wide.size();
    out6[1] = b.leaves.size();
    out6[2] = b.packets.size();
    out6[3] = b.indices.size();
    out6[4] = b.leaf_ref_count();
    out6[5] = b.binary_nodes;
}
// Raw 128-byte wide nodes (bvh.rs:196-215 layout; 3 trailing pad words zero).
void orc_tree_nodes(void* s, void* out) {
    const Bvh& b = ((SceneHandle*)s)->scene->bvh;
    std::memcpy(out, b.wide.data(), b.wide.size() * sizeof(WideNode));
}
// Per leaf: number of packed triangle lanes and of scalar prims; then one flat
// list of prim indices, leaf by leaf (packet lanes in order, then scalars).
void orc_tree_leaves(void* s, uint32_t* tri_counts, uint32_t* other_counts, uint32_t* refs) {
    const Bvh& b = ((SceneHandle*)s)->scene->bvh;
    size_t k = 0;
    for (size_t li = 0; li < b.leaves.size(); ++li) {
        const Leaf& l = b.leaves[li];
        uint32_t nt = 0;
        for (uint32_t p = l.pkt_first; p < l.pkt_first + l.pkt_count; ++p)
            for (int lane = 0; lane < 4; ++lane)
                if (b.packets[p].active & (1u << lane)) {
                    refs[k++] = b.packets[p].prim[lane];
                    nt++;
                }
        for (uint32_t i = l.idx_first; i < l.idx_first + l.idx_count; ++i) refs[k++] = b.indices[i];
        tri_counts[li] = nt;
        other_counts[li] = l.idx_count;
    }
}
// Average packet occupancy inputs (bvh.rs:1571-1579)
uint64_t orc_packet_lanes(void* s) {
    const Bvh& b = ((SceneHandle*)s)->scene->bvh;
    uint64_t n = 0;
    for (const Tri4& p : b.packets) n += __builtin_popcount(p.active);
    return n;
}

// ---- queries ---------------------------------------------------------------
// stats_out (optional, 16 u64): queries[2], nodes[2], leaves[2], packets[2],
// prim_tests[2], tri_lanes, sphere, curve, cubic, instance tests, stack high water.
static void write_stats(const Stats& st, uint64_t* o) {
    o[0] = st.queries[0];
    o[1] = st.queries[1];
    o[2] = st.nodes[0];
    o[3] = st.nodes[1];
    o[4] = st.leaves[0];
    o[5] = st.leaves[1];
    o[6] = st.packets[0];
    o[7] = st.packets[1];
    o[8] = st.prim_tests[0];
    o[9] = st.prim_tests[1];
    o[10] = st.tri_lanes;
    o[11] = st.sphere_tests;
    o[12] = st.curve_tests;
    o[13] = st.cubic_tests;
    o[14] = st.instance_tests;
    o[15] = st.stack_high_water;
}

// Closest hit for n rays. time/mask may be NULL (0 / MASK_ALL). A miss writes
// t = +inf, ids = INVALID_ID, zero normal. Returns seconds spent in the loop.
double orc_intersect_n(void* s, uint64_t n, const void* rays_, const float* time, const uint32_t* mask, void* hits_,
                       uint8_t* front_face, int threads, uint64_t* stats_out) {
    const Scene& sc = *((SceneHandle*)s)->scene;
    const RayRec* rays = (const RayRec*)rays_;
    HitRec* hits = (HitRec*)hits_;
    int nt = resolve_threads(threads);
    std::vector<Stats> per(stats_out ? nt : 0);
    auto t0 = std::chrono::steady_clock::now();
    parallel_ranges(n, nt, [&](size_t a, size_t b, int tid) {
        Ctx cx;
        if (stats_out) cx.stats = &per[tid];
        for (size_t i = a; i < b; ++i) {
            const RayRec& r = rays[i];
            Ray ray{{r.ox, r.oy, r.oz}, {r.dx, r.dy, r.dz}, time ? time[i] : 0.0f, mask ? mask[i] : MASK_ALL};
            auto h = scene_intersect(sc, ray, r.tmin, r.tmax, cx);
            if (h) {
                hits[i] = {h->t, h->u, h->v, h->geom_id, h->prim_id, h->normal.x, h->normal.y, h->normal.z};
                if (front_face) front_face[i] = h->front_face ? 1 : 0;
            } else {
                hits[i] = {INF, 0, 0, INVALID_ID, INVALID_ID, 0, 0, 0};
                if (front_face) front_face[i] = 0;
            }
        }
    });
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats_out) {
        Stats tot;
        for (auto& p : per) tot.add(p);
        write_stats(tot, stats_out);
    }
    return secs;
}

double orc_occluded_n(void* s, uint64_t n, const void* rays_, const float* time, const uint32_t* mask, uint8_t* out,
                      int threads, uint64_t* stats_out) {
    const Scene& sc = *((SceneHandle*)s)->scene;
    const RayRec* rays = (const RayRec*)rays_;
    int nt = resolve_threads(threads);
    std::vector<Stats> per(stats_out ? nt : 0);
    auto t0 = std::chrono::steady_clock::now();
    parallel_ranges(n, nt, [&](size_t a, size_t b, int tid) {
        Ctx cx;
        if (stats_out) cx.stats = &per[tid];
        for (size_t i = a; i < b; ++i) {
            const RayRec& r = rays[i];
            Ray ray{{r.ox, r.oy, r.oz}, {r.dx, r.dy, r.dz}, time ? time[i] : 0.0f, mask ? mask[i] : MASK_ALL};
            out[i] = sc.occluded(ray, r.tmin, r.tmax, cx) ? 1 : 0;
        }
    });
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats_out) {
        Stats tot;
        for (auto& p : per) tot.add(p);
        write_stats(tot, stats_out);
    }
    return secs;
}

// Linear scan over the scene's primitives in input order (bvh.rs:1460-1470),
// the reference's own cross-check for the BVH.
int orc_linear_scan(void* s, const void* ray_, float time, uint32_t mask, void* hit_) {
    const Scene& sc = *((SceneHandle*)s)->scene;
    const RayRec& r = *(const RayRec*)ray_;
    Ray ray{{r.ox, r.oy, r.oz}, {r.dx, r.dy, r.dz}, time, mask};
    Ctx cx;
    float closest = r.tmax;
    std::optional<PrimHit> best;
    for (const PrimNode& p : sc.bvh.prims) {
        auto h = p.hit(ray, r.tmin, closest, cx);
        if (h) {
            closest = h->t;
            best = h;
        }
    }
    HitRec* out = (HitRec*)hit_;
    if (!best) {
        *out = {INF, 0, 0, INVALID_ID, INVALID_ID, 0, 0, 0};
        return 0;
    }
    *out = {best->t, best->u, best->v, best->geom_id, best->prim_id, best->outward.x, best->outward.y, best->outward.z};
    return 1;
}

// ---- unit hooks for the known-answer tests ---------------------------------
int orc_test_triangle_intersect(const float* o, const float* d, const float* v9, float t_min, float t_max, float* tuv) {
    Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    auto r = triangle_intersect(ray, {v9[0], v9[1], v9[2]}, {v9[3], v9[4], v9[5]}, {v9[6], v9[7], v9[8]}, t_min, t_max);
    if (!r) return 0;
    tuv[0] = r->t;
    tuv[1] = r->u;
    tuv[2] = r->v;
    return 1;
}
// tris: n x (9 floats), masks: n; out: hits, fallback, t[4], u[4], v[4], active
void orc_test_tri4_intersect(const float* o, const float* d, const float* tris, const uint32_t* masks, int n,
                             uint32_t ray_mask, float t_min, float t_max, uint32_t* bits3, float* tuv12) {
    TriIn in[4];
    for (int i = 0; i < n; ++i) {
        const float* v = tris + 9 * i;
        in[i] = TriIn{{v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}, (uint32_t)i, masks[i]};
    }
    Tri4 pk(in, n);
    Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    Hit4 h = pk.intersect(RayShear(ray), ray_mask, t_min, t_max);
    bits3[0] = h.hits;
    bits3[1] = h.fallback;
    bits3[2] = pk.active;
    for (int l = 0; l < 4; ++l) {
        tuv12[l] = h.t[l];
        tuv12[4 + l] = h.u[l];
        tuv12[8 + l] = h.v[l];
    }
}
int orc_test_rounded_cone(const float* o, const float* d, const float* p0, const float* p1, float r0, float r1,
                          float t_min, float t_max, float* tn4) {
    Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    auto r = rounded_cone_intersect(ray, {p0[0], p0[1], p0[2]}, {p1[0], p1[1], p1[2]}, r0, r1, t_min, t_max);
    if (!r) return 0;
    tn4[0] = r->t;
    tn4[1] = r->n.x;
    tn4[2] = r->n.y;
    tn4[3] = r->n.z;
    return 1;
}
int orc_test_cubic_curve(const float* o, const float* d, const float* cp12, float r0, float r1, float t_min, float t_max,
                         float* tn4) {
    Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    V3 cp[4];
    for (int k = 0; k < 4; ++k) cp[k] = {cp12[3 * k], cp12[3 * k + 1], cp12[3 * k + 2]};
    auto r = cubic_curve_intersect(ray, cp, r0, r1, t_min, t_max);
    if (!r) return 0;
    tn4[0] = r->t;
    tn4[1] = r->n.x;
    tn4[2] = r->n.y;
    tn4[3] = r->n.z;
    return 1;
}
uint32_t orc_test_flatness_depth(const float* cp12, float max_width) {
    V3 cp[4];
    for (int k = 0; k < 4; ++k) cp[k] = {cp12[3 * k], cp12[3 * k + 1], cp12[3 * k + 2]};
    return flatness_depth(cp, max_width);
}
int orc_test_clip_triangle_aabb(const float* v9, int axis, float mn, float mx, float* out6) {
    auto r = clip_triangle_aabb({v9[0], v9[1], v9[2]}, {v9[3], v9[4], v9[5]}, {v9[6], v9[7], v9[8]}, axis, mn, mx);
    if (!r) return 0;
    out6[0] = r->minimum.x;
    out6[1] = r->minimum.y;
    out6[2] = r->minimum.z;
    out6[3] = r->maximum.x;
    out6[4] = r->maximum.y;
    out6[5] = r->maximum.z;
    return 1;
}
uint32_t orc_sizeof_wide_node() { return (uint32_t)sizeof(WideNode); }

}  // extern "C"

// ---- integrator (crt_oracle_render.hpp) -----------------------------------------------------
#include "../crust-render_b200/csrc/shade/abi_convert.h"
#include "crt_oracle_render.hpp"

namespace {
struct WorldHandle {
    orcr::World world;
};
shade::V3 sv(const float* p) { return shade::v3(p[0], p[1], p[2]); }
void put3(float* o, shade::V3 v) {
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}
shade::SurfaceHit surface(const float* p, const float* n, int front_face) {
    shade::SurfaceHit r;
    r.p = sv(p);
    r.normal = sv(n);
    r.t = 1.0f;
    r.front_face = front_face != 0;
    return r;
}
}  // namespace

extern "C" {
void* orc_world_new(void* scene, const crt_material* materials, size_t n_materials, const uint32_t* geom_material, size_t n_geoms,
                    const crt_light* lights, size_t n_lights) {
    auto* h = new WorldHandle();
    h->world.scene = ((SceneHandle*)scene)->scene;
    for (size_t i = 0; i < n_materials; ++i) h->world.materials.push_back(shade::material_from_abi(materials[i]));
    h->world.geom_material.assign(geom_material, geom_material + n_geoms);
    for (size_t i = 0; i < n_lights; ++i) h->world.lights.push_back(shade::light_from_abi(lights[i]));
    return h;
}
void orc_world_free(void* w) { delete (WorldHandle*)w; }
int orc_world_set_environment(void* w, uint32_t light_index, uint32_t width, uint32_t height, const float* rgb, const float* l2w9) {
    orcr::World& W = ((WorldHandle*)w)->world;
    if (light_index >= W.lights.size() || W.lights[light_index].kind != shade::LIGHT_DOME) return 1;
    W.env_tables.emplace_back();
    if (!W.env_tables.back().build((int32_t)width, (int32_t)height, rgb)) {
        W.env_tables.pop_back();
        return 1;
    }
    shade::light_set_environment(W.lights[light_index], (int32_t)W.env_tables.size() - 1, l2w9);
    W.envs.clear();
    for (const auto& t : W.env_tables) W.envs.push_back(t.view());
    return 0;
}
// Scene::with_volumes (tracer.rs:167-170)
namespace {
void set_volumes(std::vector<shade::VolumeRegion>& regions, std::vector<std::vector<float>>& grids, const crt_volume* v, size_t n) {
    regions.clear();
    grids.clear();
    grids.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const float* data = nullptr;
        if (v[i].field == CRT_FIELD_GRID) {
            size_t cells = (size_t)v[i].grid_dims[0] * v[i].grid_dims[1] * v[i].grid_dims[2];
            grids.emplace_back(v[i].grid_data, v[i].grid_data + cells);
            data = grids.back().data();
        }
        regions.push_back(shade::volume_from_abi(v[i], data, data));
    }
}
struct VolumesHandle {
    std::vector<shade::VolumeRegion> regions;
    std::vector<std::vector<float>> grids;
};
}  // namespace
int orc_world_set_volumes(void* w, const crt_volume* v, size_t n) {
    orcr::World& W = ((WorldHandle*)w)->world;
    for (size_t i = 0; i < n; ++i)
        if (!shade::volume_abi_valid(v[i])) return 1;
    set_volumes(W.volumes, W.grids, v, n);
    return 0;
}
// volume.rs unit-test surface: a bare Volumes aggregate and an Rng whose state the caller keeps between calls
void* orc_volumes_new(const crt_volume* v, size_t n) {
    auto* h = new VolumesHandle();
    set_volumes(h->regions, h->grids, v, n);
    return h;
}
void orc_volumes_free(void* h) { delete (VolumesHandle*)h; }
uint32_t orc_rng_new(uint32_t seed) { return shade::rng_new(seed).state; }
float orc_rng_next(uint32_t* state) {
    shade::Rng r{*state};
    float f = shade::rng_next_f32(r);
    *state = r.state;
    return f;
}
void orc_volumes_region_info(void* h, uint32_t i, float* out) {  // sigma_s3, sigma_a3, g, majorant, aabb6, homogeneous
    const shade::VolumeRegion& r = ((VolumesHandle*)h)->regions[i];
    put3(out, r.sigma_s);
    put3(out + 3, r.sigma_a);
    out[6] = r.g;
    out[7] = r.majorant;
    put3(out + 8, r.aabb_min);
    put3(out + 11, r.aabb_max);
    out[14] = r.field == shade::FIELD_HOMOGENEOUS ? 1.0f : 0.0f;
}
// n independent Volumes::transmittance calls on one ray, consuming one Rng stream (out: 3 floats each)
void orc_volumes_transmittance_n(void* h, const float* o, const float* d, float t_eps, float t_max, uint32_t* state, size_t n, float* out) {
    auto* H = (VolumesHandle*)h;
    shade::Rng rng{*state};
    for (size_t i = 0; i < n; ++i) put3(out + 3 * i, shade::volumes_transmittance(H->regions.data(), (int)H->regions.size(), sv(o), sv(d), t_eps, t_max, rng));
    *state = rng.state;
}
// n Volumes::sample_interaction calls; out per call: scatter, t, p3, weight-or-transmittance3, emitted3, n_lobes, (w, g) of the first two lobes
void orc_volumes_sample_n(void* h, const float* o, const float* d, float t_eps, float t_max, uint32_t* state, size_t n, float* out) {
    auto* H = (VolumesHandle*)h;
    shade::Rng rng{*state};
    for (size_t i = 0; i < n; ++i) {
        shade::VolumeEvent ev;
        ev.t = 0.0f;
        ev.p = shade::splat(0.0f);
        shade::volumes_sample_interaction(H->regions.data(), (int)H->regions.size(), sv(o), sv(d), t_eps, t_max, rng, ev);
        float* q = out + 16 * i;
        q[0] = ev.scatter ? 1.0f : 0.0f;
        q[1] = ev.t;
        put3(q + 2, ev.p);
        put3(q + 5, ev.weight);
        put3(q + 8, ev.emitted);
        q[11] = (float)ev.phase.n;
        for (int k = 0; k < 2; ++k) {
            q[12 + 2 * k] = k < ev.phase.n ? ev.phase.w[k] : 0.0f;
            q[13 + 2 * k] = k < ev.phase.n ? ev.phase.g[k] : 0.0f;
        }
    }
    *state = rng.state;
}
float orc_volumes_field_density(void* h, uint32_t i, const float* u) { return shade::field_density(((VolumesHandle*)h)->regions[i], sv(u)); }
float orc_volumes_density(void* h, uint32_t i, const float* p_world) { return shade::region_density(((VolumesHandle*)h)->regions[i], sv(p_world)); }
int orc_volumes_intersect(void* h, uint32_t i, const float* o, const float* d, float* out2) {
    return shade::region_intersect(((VolumesHandle*)h)->regions[i], sv(o), sv(d), out2[0], out2[1]) ? 1 : 0;
}
float orc_vol_hash3(int32_t x, int32_t y, int32_t z, uint32_t seed) { return shade::vol_hash3(x, y, z, seed); }
float orc_hg_phase(float cos_theta, float g) { return shade::hg_phase(cos_theta, g); }
float orc_phase_single_pdf(float g, float cos_theta) {  // PhaseMix::single(g).pdf(cos_theta)
    shade::PhaseMix m;
    m.n = 1;
    m.w[0] = 1.0f;
    m.g[0] = g;
    return shade::phase_pdf(m, cos_theta);
}
void orc_mat4_inverse(const float* m, float* out) { shade::mat4_inverse(m, out); }
// A light of a world with its environment table (usd_scene.rs:873-922 probes lights through the scene).
int orc_world_light_sample(void* w, uint32_t index, const float* from, float u, float v, float* out8) {
    orcr::World& W = ((WorldHandle*)w)->world;
    shade::LightSample ls;
    if (index >= W.lights.size() || !shade::light_sample_li(W.lights[index], sv(from), u, v, ls, W.envs.data())) return 0;
    put3(out8, ls.direction);
    put3(out8 + 3, ls.radiance);
    out8[6] = ls.distance;
    out8[7] = ls.pdf;
    return 1;
}
int orc_world_light_escaped(void* w, uint32_t index, const float* dir, float* radiance3, float* pdf) {
    orcr::World& W = ((WorldHandle*)w)->world;
    shade::V3 r;
    if (index >= W.lights.size() || !shade::light_escaped(W.lights[index], sv(dir), r, *pdf, W.envs.data())) return 0;
    put3(radiance3, r);
    return 1;
}
// environment.rs hooks for tests/test_shade_kats.py: op 0 = sample(u1, u2) -> dir3, radiance3, pdf (returns 0 when
// it declines), 1 = pdf(dir), 2 = radiance(dir), 3 = direction_to_uv, 4 = uv_to_direction (ops 3, 4 take a null map)
void* orc_env_new(uint32_t width, uint32_t height, const float* rgb) {
    auto* t = new shade::EnvMapTables();
    if (!t->build((int32_t)width, (int32_t)height, rgb)) {
        delete t;
        return nullptr;
    }
    return t;
}
void orc_env_free(void* e) { delete (shade::EnvMapTables*)e; }
int orc_env_call(void* e, int op, const float* a, float* out) {
    if (op == 3) { shade::env_direction_to_uv(sv(a), out[0], out[1]); return 1; }
    if (op == 4) { put3(out, shade::env_uv_to_direction(a[0], a[1])); return 1; }
    if (!e) return -1;
    shade::EnvMap m = ((shade::EnvMapTables*)e)->view();
    if (op == 0) {
        shade::V3 d, r;
        float pdf;
        if (!shade::env_sample(m, a[0], a[1], d, r, pdf)) return 0;
        put3(out, d);
        put3(out + 3, r);
        out[6] = pdf;
        return 1;
    }
    if (op == 1) { out[0] = shade::env_pdf(m, sv(a)); return 1; }
    if (op == 2) { put3(out, shade::env_radiance(m, sv(a))); return 1; }
    return -1;
}
// returns seconds
double orc_render(void* w, const crt_camera* cam, const crt_render_settings* st, float* fb, double* var_map, uint64_t* stats8, int threads) {
    shade::Camera c = shade::camera_from_abi(*cam);
    shade::Filter f = shade::filter_new(st->pixel_filter, st->filter_radius);
    shade::RenderSettings s = shade::settings_from_abi(*st);
    auto t0 = std::chrono::steady_clock::now();
    orcr::RayStats rs = orcr::render(((WorldHandle*)w)->world, c, f, s, fb, var_map, threads, st->tile_size, st->tile_rank, st->tile_world);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats8) {
        uint64_t v[8] = {rs.camera_rays, rs.closest_hit, rs.shadow_rays, rs.vertices, rs.rr_tested, rs.rr_killed, rs.ended_escaped, rs.ended_depth};
        for (int i = 0; i < 8; ++i) stats8[i] = v[i];
    }
    return dt;
}

// ---- shading-math hooks for the known-answer tests (tests/test_shade_kats.py) ----------------
void orc_material_default(crt_material* out) {
    shade::Material m = shade::material_default();
    shade::material_to_abi(m, out);
}
void orc_material_pmf(const crt_material* m, float* out5) {
    shade::Material mm = shade::material_from_abi(*m);
    out5[0] = mm.pmf.p_diffuse;
    out5[1] = mm.pmf.p_specular;
    out5[2] = mm.pmf.p_coat;
    out5[3] = mm.pmf.p_fuzz;
    out5[4] = mm.pmf.p_transmission;
}
int orc_material_medium(const crt_material* m, float* sigma_a3, float* sigma_s3, float* g) {
    shade::Material mm = shade::material_from_abi(*m);
    put3(sigma_a3, mm.medium.sigma_a);
    put3(sigma_s3, mm.medium.sigma_s);
    *g = mm.medium.g;
    return (int)mm.has_medium;
}
int orc_material_eval(const crt_material* m, const float* ray_dir, const float* p, const float* normal, int front_face, const float* wi, float* value3,
                      float* pdf) {
    shade::Material mm = shade::material_from_abi(*m);
    shade::V3 v;
    bool ok = shade::material_eval(mm, sv(ray_dir), surface(p, normal, front_face), sv(wi), v, *pdf);
    if (ok) put3(value3, v);
    return ok ? 1 : 0;
}
// out: origin3, dir3, value3, pdf, delta, enters_medium (11 floats)
int orc_material_scatter(const crt_material* m, const float* ray_dir, const float* p, const float* normal, int front_face, const float* s4, float* out11) {
    shade::Material mm = shade::material_from_abi(*m);
    shade::ScatterSample ss;
    if (!shade::material_scatter(mm, sv(ray_dir), surface(p, normal, front_face), s4, ss)) return 0;
    put3(out11, ss.origin);
    put3(out11 + 3, ss.dir);
    put3(out11 + 6, ss.value);
    out11[9] = ss.pdf;
    out11[10] = (ss.delta ? 1.0f : 0.0f) + (ss.enters_medium ? 2.0f : 0.0f);
    return 1;
}
void orc_material_emitted_directional(const crt_material* m, float cos_o, float* out3) {
    put3(out3, shade::material_emitted_directional(shade::material_from_abi(*m), cos_o));
}
// out: direction3, radiance3, distance, pdf
int orc_light_sample_li(const crt_light* l, const float* from, float u, float v, float* out8) {
    shade::LightSample ls;
    if (!shade::light_sample_li(shade::light_from_abi(*l), sv(from), u, v, ls)) return 0;
    put3(out8, ls.direction);
    put3(out8 + 3, ls.radiance);
    out8[6] = ls.distance;
    out8[7] = ls.pdf;
    return 1;
}
float orc_light_pdf_at_point(const crt_light* l, const float* from, const float* point) {
    return shade::light_pdf_at_point(shade::light_from_abi(*l), sv(from), sv(point));
}
int orc_light_escaped(const crt_light* l, const float* dir, float* radiance3, float* pdf) {
    shade::V3 r;
    if (!shade::light_escaped(shade::light_from_abi(*l), sv(dir), r, *pdf)) return 0;
    put3(radiance3, r);
    return 1;
}
int orc_light_pick(int n, float u) { return shade::light_pick(n, u); }
void orc_camera_ray(const crt_camera* cam, float s, float t, float lu, float lv, float* origin3, float* dir3) {
    shade::V3 o, d;
    shade::camera_get_ray(shade::camera_from_abi(*cam), s, t, lu, lv, o, d);
    put3(origin3, o);
    put3(dir3, d);
}
void orc_filter_sample(uint32_t kind, float radius, float u, float* offset, float* weight) {
    shade::Filter f = shade::filter_new(kind, radius);
    shade::filter_sample(f, u, *offset, *weight);
}
float orc_filter_eval(uint32_t kind, float radius, float x) {
    shade::Filter f = shade::filter_new(kind, radius);
    return shade::filter_eval(f, x);
}
// sampler: root = new(x, y, frame, index); then new_domain for each key; draws one 4-D block + one rnd.
void orc_sampler_draw(int32_t x, int32_t y, int32_t frame, int32_t index, const int32_t* keys, int n_keys, float* out4, float* rnd) {
    shade::PathSampler s = shade::sampler_new(x, y, frame, index);
    for (int i = 0; i < n_keys; ++i) s = shade::new_domain(s, keys[i]);
    shade::draw_sample4(s, out4);
    *rnd = shade::draw_rnd1(s);
}
float orc_mis_weight(uint32_t strategy, int light_side, float a, float b) {
    return light_side ? shade::strategy_light_weight(strategy, a, b) : shade::strategy_bounce_weight(strategy, a, b);
}
}  // extern "C"

// Generic hook over the leaf shading functions (shade/brdf.h, material.h), addressed by name so
// tests/test_shade_kats.py can restate crust-core's unit tests. Returns the number of floats written,
// -1 for an unknown name.
extern "C" int orc_shade_fn(const char* name, const crt_material* mat, const float* a, float* out) {
    using namespace shade;
    std::string n(name);
    Material m = mat ? material_from_abi(*mat) : material_default();
    auto put = [&](shade::V3 v) { put3(out, v); return 3; };
    if (n == "fresnel_dielectric") { out[0] = fresnel_dielectric(a[0], a[1], a[2]); return 1; }
    if (n == "f0_from_ior") { out[0] = f0_from_ior(a[0]); return 1; }
    if (n == "roughness_to_alpha_aniso") { roughness_to_alpha_aniso(a[0], a[1], out[0], out[1]); return 2; }
    if (n == "sheen_charlie") { out[0] = sheen_charlie(a[0], a[1], a[2], a[3]); return 1; }
    if (n == "coat_darkening_factor") return put(coat_darkening_factor(sv(a), a[3], a[4]));
    if (n == "thin_film_fresnel") return put(thin_film_fresnel(a[0], a[1], a[2], a[3], a[4]));
    if (n == "thin_film_fresnel_metal") return put(thin_film_fresnel_metal(a[0], a[1], a[2], sv(a + 3), a[6]));
    if (n == "dispersive_ior") return put(dispersive_ior(a[0], a[1], a[2]));
    if (n == "cauchy_ior") { out[0] = cauchy_ior(a[0], a[1], a[2]); return 1; }
    if (n == "eon_diffuse") return put(eon_diffuse(sv(a), a[3], sv(a + 4), sv(a + 7)));
    if (n == "eon_albedo_exact") { out[0] = eon_albedo_exact(a[0], a[1]); return 1; }
    if (n == "eon_albedo_approx") { out[0] = eon_albedo_approx(a[0], a[1]); return 1; }
    if (n == "fresnel_f82_tint") return put(fresnel_f82_tint(a[0], sv(a + 1), sv(a + 4)));
    if (n == "fresnel_schlick") return put(fresnel_schlick(a[0], sv(a + 1)));
    if (n == "eval_all") return put(eval_all(m, sv(a), sv(a + 3), a[6] != 0.0f));
    if (n == "pdf_all") { out[0] = pdf_all(m, sv(a), sv(a + 3), a[6] != 0.0f); return 1; }
    if (n == "eval_coat") return put(eval_coat(m, sv(a), sv(a + 3), sv(a + 6), a[9], a[10]));
    if (n == "coat_attenuation") return put(coat_attenuation(m, a[0], a[1]));
    if (n == "coat_passage") return put(coat_passage(m, a[0]));
    if (n == "transmission_iors") return put(transmission_iors(m));
    if (n == "material_emitted") return put(material_emitted(m));
    if (n == "medium_from_transmission") {
        Medium md = medium_from_transmission(sv(a), a[3], sv(a + 4), a[7]);
        put3(out, md.sigma_a); put3(out + 3, md.sigma_s); out[6] = md.g; return 7;
    }
    if (n == "medium_from_subsurface") {
        Medium md = medium_from_subsurface(sv(a), a[3], sv(a + 4), a[7]);
        put3(out, md.sigma_a); put3(out + 3, md.sigma_s); out[6] = md.g; return 7;
    }
    if (n == "medium_transmittance") {
        Medium md{sv(a), sv(a + 3), a[6]};
        return put(medium_transmittance(md, a[7]));
    }
    if (n == "sample_henyey_greenstein") return put(sample_henyey_greenstein(sv(a), a[3], a[4], a[5]));
    if (n == "ggx_d_aniso") { out[0] = ggx_d_aniso(a[0], a[1], a[2], a[3], a[4]); return 1; }
    if (n == "ggx_g2_smith_aniso") { out[0] = ggx_g2_smith_aniso(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); return 1; }
    if (n == "sample_vndf_ggx_aniso_local") return put(sample_vndf_ggx_aniso_local(sv(a), a[3], a[4], a[5], a[6]));
    if (n == "pdf_vndf_ggx_aniso_local") { out[0] = pdf_vndf_ggx_aniso_local(sv(a), sv(a + 3), a[6], a[7]); return 1; }
    if (n == "cosine_hemisphere") return put(cosine_hemisphere(a[0], a[1]));
    if (n == "concentric_disk") return put(concentric_disk(a[0], a[1]));
    if (n == "align_to_normal") return put(align_to_normal(sv(a), sv(a + 3)));
    if (n == "power_heuristic") { out[0] = power_heuristic(a[0], a[1]); return 1; }
    if (n == "balance_heuristic") { out[0] = balance_heuristic(a[0], a[1]); return 1; }
    return -1;
}
extern "C" void orc_filter_sample_n(uint32_t kind, float radius, size_t n, const float* u, float* offset, float* weight) {
    shade::Filter f = shade::filter_new(kind, radius);
    for (size_t i = 0; i < n; ++i) shade::filter_sample(f, u[i], offset[i], weight[i]);
}
