// C ABI of libcrt_b200.so (include/crt.h): builder, commit + upload, scene
// queries and the batched / host-buffer query entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "crt.h"
#include "host_build.h"
#include "launch.h"
#include "internal.h"

using namespace crt;

// ---------------------------------------------------------------------------
// process-wide state
// ---------------------------------------------------------------------------
namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? CRT_ERR_NO_DEVICE : CRT_ERR_CUDA;
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t _e = (call);                          \
        if (_e != cudaSuccess) return fail_cuda(_e, #call); \
    } while (0)

struct Device {
    std::mutex mu;
    bool ready = false;
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t total_mem = 0, l2_bytes = 0;
    unsigned int* counters = nullptr;  // ring of ray counters, one per in-flight launch
    unsigned int* status = nullptr;
    std::atomic<uint64_t> next_counter{0};
    std::atomic<uint64_t> launches{0};
    unsigned long long* trace_stats = nullptr;  // 7 counters of the counting traversal build
    std::atomic<bool> count_traversal{false};
    size_t l2_persist_bytes = 0;  // L2 set aside for the top levels of the BVH (0 = feature off)
    std::atomic<uint64_t> l2_applied{0}, l2_refused{0};
} g_dev;
constexpr int kCounterRing = 4096;

int ensure_device() {
    std::lock_guard<std::mutex> lk(g_dev.mu);
    if (g_dev.ready) return CRT_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CRT_ERR_NO_DEVICE, "no CUDA device: libcrt_b200 has no host execution path");
    }
    CK(cudaGetDevice(&g_dev.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_dev.device));
    g_dev.sm_count = prop.multiProcessorCount;
    g_dev.cc_major = prop.major;
    g_dev.cc_minor = prop.minor;
    g_dev.total_mem = prop.totalGlobalMem;
    g_dev.l2_bytes = (size_t)prop.l2CacheSize;
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device is sm_%d%d; this library carries sm_100a code only", prop.major, prop.minor);
        return fail(CRT_ERR_NO_DEVICE, buf);
    }
    {
        // L2 persistence for the top levels of the tree (BASELINE north_star): a slice of the 126 MB L2 is set aside and
        // every traversal launch carries an access-policy window over the leading, breadth-first-ordered nodes of its scene,
        // so the lines every ray touches are not evicted by the streaming ray / hit / leaf-record traffic.
        // CRT_L2_PERSIST_MB overrides the size (0 switches the feature off).
        size_t want = (size_t)16 << 20;
        if (const char* e = getenv("CRT_L2_PERSIST_MB")) want = (size_t)std::max(0L, atol(e)) << 20;
        want = std::min(want, (size_t)prop.persistingL2CacheMaxSize);
        if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
            g_dev.l2_persist_bytes = std::min(want, (size_t)prop.accessPolicyMaxWindowSize);
        cudaGetLastError();
    }
    CK(cudaMalloc(&g_dev.counters, sizeof(unsigned int) * (kCounterRing + 1)));
    CK(cudaMemset(g_dev.counters, 0, sizeof(unsigned int) * (kCounterRing + 1)));
    g_dev.status = g_dev.counters + kCounterRing;
    CK(cudaMalloc(&g_dev.trace_stats, 8 * sizeof(unsigned long long)));
    CK(cudaMemset(g_dev.trace_stats, 0, 8 * sizeof(unsigned long long)));
    g_dev.ready = true;
    return CRT_OK;
}

// ---- L2 persistence of the top levels ------------------------------------------
// The builder emits nodes in the reference's DFS pre-order, which scatters the top levels over the array. For the device
// copy the first `budget` bytes' worth of nodes are re-ordered breadth-first from the root (the rest keep their DFS order
// behind them) and child words are renumbered: same tree, same child lanes, same visiting order — only addresses change.
// Returns the size of the breadth-first prefix in bytes (0 when the tree is smaller than two nodes or the feature is off).
template <class NodeVec>
size_t top_levels_first(const NodeVec& nodes, size_t budget, std::vector<DevNode>& out) {
    const size_t n = nodes.size();
    size_t n_top = std::min(n, budget / sizeof(DevNode));
    if (n_top < 2) return 0;
    if (n_top == n) return n * sizeof(DevNode);  // the whole tree fits the window: no need to move anything
    std::vector<uint32_t> order;  // new index -> old index
    order.reserve(n);
    std::vector<uint32_t> new_of(n, 0xFFFFFFFFu);
    order.push_back(0);
    new_of[0] = 0;
    for (size_t head = 0; head < order.size() && order.size() < n_top; ++head) {
        const DevNode& nd = nodes[order[head]];
        for (int k = 0; k < 4 && order.size() < n_top; ++k) {
            uint32_t c = nd.child[k];
            if (c == CRT_EMPTY_LANE || (c & CRT_LEAF_BIT) || new_of[c] != 0xFFFFFFFFu) continue;
            new_of[c] = (uint32_t)order.size();
            order.push_back(c);
        }
    }
    n_top = order.size();
    for (uint32_t old = 0; old < n; ++old)
        if (new_of[old] == 0xFFFFFFFFu) {
            new_of[old] = (uint32_t)order.size();
            order.push_back(old);
        }
    out.resize(n);
    for (size_t i = 0; i < n; ++i) {
        DevNode nd = nodes[order[i]];
        for (int k = 0; k < 4; ++k) {
            uint32_t c = nd.child[k];
            if (c != CRT_EMPTY_LANE && !(c & CRT_LEAF_BIT)) nd.child[k] = new_of[c];
        }
        out[i] = nd;
    }
    return n_top * sizeof(DevNode);
}
// Attaches the window to `stream` (a stream attribute: it applies to the launches that follow) and returns whether one was
// set. Failure is not an error — the hint is an optimisation only. The caller clears it again with clear_top_levels_hint
// once its launches are enqueued: the stream is the caller's (a user or torch stream) and must not keep a window over
// memory this library may free, nor pass it on to the caller's own kernels.
bool hint_top_levels(const crt_scene* s, cudaStream_t stream) {
    if (!s->top_bytes || !g_dev.l2_persist_bytes) return false;
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = s->d_nodes;
    attr.accessPolicyWindow.num_bytes = std::min(s->top_bytes, g_dev.l2_persist_bytes);
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) {
        cudaGetLastError();
        g_dev.l2_refused.fetch_add(1);
        return false;
    }
    g_dev.l2_applied.fetch_add(1);
    return true;
}
void clear_top_levels_hint(cudaStream_t stream) {
    cudaStreamAttrValue attr{};  // num_bytes = 0 disables the window
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

}  // namespace

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
struct crt_builder {
    std::vector<Geom> geoms;
};


namespace {
ChildInfo child_info_of(crt_scene* s) { return ChildInfo{s->non_empty, s->root, s->has_motion, s->instance_depth}; }

Xform xform_from12(const float* m) {
    return Xform{{m[0], m[1], m[2]}, {m[3], m[4], m[5]}, {m[6], m[7], m[8]}, {m[9], m[10], m[11]}};
}
void release_geom(Geom& g) {
    if (g.type == G_INSTANCE && g.child) crt_scene_release(g.child);
    g.child = nullptr;
}
Geom mesh_geom(const float* v, size_t nv, const uint32_t* idx, size_t nt, const float* nrm, size_t nn) {
    Geom g;
    g.type = G_MESH;
    if (nv) g.verts.assign(v, v + 3 * nv);
    if (nt) g.indices.assign(idx, idx + 3 * nt);
    if (nrm) {
        g.has_normals = true;
        g.normals.assign(nrm, nrm + 3 * nn);
    }
    return g;
}
Geom sphere_geom(const float c[3], float r) {
    Geom g;
    g.type = G_SPHERE;
    g.sphere[0] = c[0], g.sphere[1] = c[1], g.sphere[2] = c[2], g.sphere[3] = r;
    return g;
}
Geom curves_geom(const crt_curve_segment* s, size_t n) {
    Geom g;
    g.type = G_CURVES;
    g.segs.resize(8 * n);
    for (size_t i = 0; i < n; ++i) {
        float* o = &g.segs[8 * i];
        o[0] = s[i].p0[0], o[1] = s[i].p0[1], o[2] = s[i].p0[2];
        o[3] = s[i].p1[0], o[4] = s[i].p1[1], o[5] = s[i].p1[2];
        o[6] = s[i].r0, o[7] = s[i].r1;
    }
    return g;
}
Geom cubics_geom(const crt_cubic_curve_segment* s, size_t n) {
    Geom g;
    g.type = G_CUBICS;
    g.segs.resize(14 * n);
    for (size_t i = 0; i < n; ++i) {
        float* o = &g.segs[14 * i];
        for (int k = 0; k < 4; ++k)
            for (int c = 0; c < 3; ++c) o[3 * k + c] = s[i].cp[k][c];
        o[12] = s[i].r0, o[13] = s[i].r1;
    }
    return g;
}
Geom instance_geom(crt_scene* child, const float* l2w, const float* l2w_end) {
    Geom g;
    g.type = G_INSTANCE;
    g.child = child;
    crt_scene_retain(child);
    g.l2w = xform_from12(l2w);
    if (l2w_end) {
        g.has_end = true;
        g.l2w_end = xform_from12(l2w_end);
    }
    return g;
}
int attach(crt_builder* b, Geom&& g, uint32_t mask, uint32_t* id) {
    g.mask = mask;
    b->geoms.push_back(std::move(g));
    if (id) *id = (uint32_t)b->geoms.size() - 1;
    return CRT_OK;
}
int replace(crt_builder* b, uint32_t id, Geom&& g) {
    if (!b) return fail(CRT_ERR_INVALID_ARG, "null builder");
    if (id >= b->geoms.size()) {
        release_geom(g);
        return fail(CRT_ERR_UNKNOWN_GEOMETRY, "set_geometry: id was never attached");
    }
    g.mask = b->geoms[id].mask;  // scene.rs:199-201 keeps the mask
    release_geom(b->geoms[id]);
    b->geoms[id] = std::move(g);
    return CRT_OK;
}
template <class T>
int upload(void** dst, const T* src, size_t count) {
    *dst = nullptr;
    if (count == 0) return CRT_OK;
    cudaError_t e = cudaMalloc(dst, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CRT_ERR_OOM : CRT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    CK(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return CRT_OK;
}
}  // namespace

extern "C" {

const char* crt_last_error(void) { return g_err.c_str(); }

int crt_init(int device) {
    {
        std::lock_guard<std::mutex> lk(g_dev.mu);
        if (!g_dev.ready) {
            int n = 0;
            cudaError_t e = cudaGetDeviceCount(&n);
            if (e != cudaSuccess || n == 0) {
                cudaGetLastError();
                return fail(CRT_ERR_NO_DEVICE, "no CUDA device: libcrt_b200 has no host execution path");
            }
            if (device < 0 || device >= n) return fail(CRT_ERR_INVALID_ARG, "device index out of range");
            CK(cudaSetDevice(device));
        }
    }
    return ensure_device();
}

int crt_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem, uint64_t* l2_bytes) {
    int rc = ensure_device();
    if (rc) return rc;
    if (sm_count) *sm_count = g_dev.sm_count;
    if (cc_major) *cc_major = g_dev.cc_major;
    if (cc_minor) *cc_minor = g_dev.cc_minor;
    if (total_mem) *total_mem = g_dev.total_mem;
    if (l2_bytes) *l2_bytes = g_dev.l2_bytes;
    return CRT_OK;
}

// ---- builder ---------------------------------------------------------------
int crt_builder_new(crt_builder** out) {
    if (!out) return fail(CRT_ERR_INVALID_ARG, "null out");
    *out = new crt_builder();
    return CRT_OK;
}
void crt_builder_free(crt_builder* b) {
    if (!b) return;
    for (Geom& g : b->geoms) release_geom(g);
    delete b;
}
int crt_builder_reserve(crt_builder* b, size_t additional) {
    if (!b) return fail(CRT_ERR_INVALID_ARG, "null builder");
    b->geoms.reserve(b->geoms.size() + additional);
    return CRT_OK;
}
int crt_builder_count(const crt_builder* b, size_t* out) {
    if (!b || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = b->geoms.size();
    return CRT_OK;
}
int crt_attach_triangle_mesh(crt_builder* b, const float* v, size_t nv, const uint32_t* idx, size_t nt, const float* nrm,
                             size_t nn, uint32_t mask, uint32_t* id) {
    if (!b || (nv && !v) || (nt && !idx)) return fail(CRT_ERR_INVALID_ARG, "attach_triangle_mesh: null argument");
    return attach(b, mesh_geom(v, nv, idx, nt, nrm, nn), mask, id);
}
int crt_attach_sphere(crt_builder* b, const float c[3], float r, uint32_t mask, uint32_t* id) {
    if (!b || !c) return fail(CRT_ERR_INVALID_ARG, "attach_sphere: null argument");
    return attach(b, sphere_geom(c, r), mask, id);
}
int crt_attach_round_curves(crt_builder* b, const crt_curve_segment* s, size_t n, uint32_t mask, uint32_t* id) {
    if (!b || (n && !s)) return fail(CRT_ERR_INVALID_ARG, "attach_round_curves: null argument");
    return attach(b, curves_geom(s, n), mask, id);
}
int crt_attach_cubic_curves(crt_builder* b, const crt_cubic_curve_segment* s, size_t n, uint32_t mask, uint32_t* id) {
    if (!b || (n && !s)) return fail(CRT_ERR_INVALID_ARG, "attach_cubic_curves: null argument");
    return attach(b, cubics_geom(s, n), mask, id);
}
int crt_attach_instance(crt_builder* b, crt_scene* scene, const float l2w[12], const float* l2w_end, uint32_t mask,
                        uint32_t* id) {
    if (!b || !scene || !l2w) return fail(CRT_ERR_INVALID_ARG, "attach_instance: null argument");
    return attach(b, instance_geom(scene, l2w, l2w_end), mask, id);
}
int crt_attach_empty(crt_builder* b, uint32_t mask, uint32_t* id) {
    if (!b) return fail(CRT_ERR_INVALID_ARG, "null builder");
    return attach(b, Geom{}, mask, id);
}
int crt_set_triangle_mesh(crt_builder* b, uint32_t id, const float* v, size_t nv, const uint32_t* idx, size_t nt,
                          const float* nrm, size_t nn) {
    if ((nv && !v) || (nt && !idx)) return fail(CRT_ERR_INVALID_ARG, "set_triangle_mesh: null argument");
    return replace(b, id, mesh_geom(v, nv, idx, nt, nrm, nn));
}
int crt_set_sphere(crt_builder* b, uint32_t id, const float c[3], float r) {
    if (!c) return fail(CRT_ERR_INVALID_ARG, "set_sphere: null argument");
    return replace(b, id, sphere_geom(c, r));
}
int crt_set_round_curves(crt_builder* b, uint32_t id, const crt_curve_segment* s, size_t n) {
    if (n && !s) return fail(CRT_ERR_INVALID_ARG, "set_round_curves: null argument");
    return replace(b, id, curves_geom(s, n));
}
int crt_set_cubic_curves(crt_builder* b, uint32_t id, const crt_cubic_curve_segment* s, size_t n) {
    if (n && !s) return fail(CRT_ERR_INVALID_ARG, "set_cubic_curves: null argument");
    return replace(b, id, cubics_geom(s, n));
}
int crt_set_instance(crt_builder* b, uint32_t id, crt_scene* scene, const float l2w[12], const float* l2w_end) {
    if (!scene || !l2w) return fail(CRT_ERR_INVALID_ARG, "set_instance: null argument");
    return replace(b, id, instance_geom(scene, l2w, l2w_end));
}
int crt_set_empty(crt_builder* b, uint32_t id) { return replace(b, id, Geom{}); }

// ---- commit ----------------------------------------------------------------
int crt_commit_ex(crt_builder* b, uint32_t flags, int host_threads, crt_scene** out) {
    if (!b || !out) {
        crt_builder_free(b);
        return fail(CRT_ERR_INVALID_ARG, "commit: null argument");
    }
    *out = nullptr;
    const bool host_only = (flags & CRT_COMMIT_HOST_ONLY) != 0;
    if (!host_only) {
        int rc = ensure_device();
        if (rc) {
            crt_builder_free(b);
            return rc;
        }
    }
    crt_scene* s = new crt_scene();
    s->host_only = host_only;
    s->n_geoms = (uint32_t)b->geoms.size();
    // distinct children (retained by the scene; the builder's own references are dropped below)
    {
        std::unordered_set<crt_scene*> seen;
        for (Geom& g : b->geoms)
            if (g.type == G_INSTANCE && g.child && seen.insert(g.child).second) {
                crt_scene_retain(g.child);
                s->children.push_back(g.child);
            }
    }
    for (Geom& g : b->geoms)
        if (g.type == G_INSTANCE && g.child && g.child->host_only != host_only) {
            crt_builder_free(b);
            crt_scene_release(s);
            return fail(CRT_ERR_INVALID_ARG, "commit: instanced scene was committed in a different mode (host-only vs device)");
        }
    if ((flags & CRT_COMMIT_DEVICE_BUILD) && !host_only && !(flags & CRT_COMMIT_KEEP_HOST_TREE) && device_build_supported(b->geoms)) {
        int rc = device_build(b->geoms, child_info_of, s);
        crt_builder_free(b);
        if (rc) {
            crt_scene_release(s);
            return rc;
        }
        s->info.built_on_device = 1;
        *out = s;
        return CRT_OK;
    }
    HostTree* t = new HostTree();
    const char* err = nullptr;
    bool ok = build_host_tree(b->geoms, child_info_of, host_threads, *t, &err);
    if (!ok) {
        delete t;
        crt_builder_free(b);
        crt_scene_release(s);
        return fail(CRT_ERR_NESTING, err ? err : "commit failed");
    }
    s->non_empty = t->non_empty;
    s->root = t->root;
    s->has_motion = t->has_motion;
    s->instance_depth = t->instance_depth;
    s->n_prims = t->prims.size();
    uint32_t feat = 0;
    for (size_t i = 0; i < t->prims.size(); ++i) {
        switch (t->prims.kind[i]) {
            case P_TRI: s->breakdown.triangles++; break;
            case P_SPHERE: s->breakdown.spheres++, feat |= FEAT_SPHERE; break;
            case P_CURVE: s->breakdown.curve_segments++, feat |= FEAT_CURVE; break;
            case P_CUBIC: s->breakdown.cubic_curve_spans++, feat |= FEAT_CURVE; break;
            default: s->breakdown.instances++, feat |= FEAT_INSTANCE; break;
        }
    }
    if (t->prims.any_normals) feat |= FEAT_NORMALS;
    for (crt_scene* c : s->children) feat |= c->features;
    s->features = feat;
    s->extent_sum = t->extent_sum;  // scene.rs:446-455 / bvh.rs:335-345
    s->extent_max = t->extent_max;
    s->info.primitives = t->prims.size();
    s->info.binary_nodes = t->n_binary;
    s->info.wide_nodes = t->nodes.size();
    s->info.leaves = t->n_leaves;
    s->info.leaf_refs = t->n_leaf_refs;
    s->info.node_bytes = t->nodes.size() * sizeof(DevNode);
    s->info.leaf_bytes = t->leafs.size() * sizeof(uint32_t);
    s->info.build_seconds = t->build_seconds;
    s->info.instance_depth = t->instance_depth;

    s->n_nodes = t->nodes.size();
    s->n_leaf_quads = t->leafs.size() / 4;
    s->n_prim_ids = t->prim_ids.size() / 2;
    s->n_normal_floats = t->normals.size();
    s->n_motion = t->motion.size();
    // Patch child descriptor pointers into the instance records.
    for (uint32_t q : t->instance_record_q) {
        uint32_t slot = t->leafs[4 * (size_t)(q + 4) + 3];
        crt_scene* child = t->prims.inst[slot].child;
        s->inst_q.push_back(q);
        s->inst_child.push_back((uint32_t)(std::find(s->children.begin(), s->children.end(), child) - s->children.begin()));
        uint64_t ptr = (uint64_t)(uintptr_t)child->d_desc;
        t->leafs[4 * (size_t)(q + 4) + 0] = (uint32_t)(ptr & 0xFFFFFFFFu);
        t->leafs[4 * (size_t)(q + 4) + 1] = (uint32_t)(ptr >> 32);
        const uint64_t cn = (uint64_t)(uintptr_t)child->d_nodes, cl = (uint64_t)(uintptr_t)child->d_leafs;
        t->leafs[4 * (size_t)(q + 5) + 0] = (uint32_t)(cn & 0xFFFFFFFFu);
        t->leafs[4 * (size_t)(q + 5) + 1] = (uint32_t)(cn >> 32);
        t->leafs[4 * (size_t)(q + 5) + 2] = (uint32_t)(cl & 0xFFFFFFFFu);
        t->leafs[4 * (size_t)(q + 5) + 3] = (uint32_t)(cl >> 32);
        if (child->desc.has_triangles) t->leafs[4 * (size_t)q + 3] |= REC_CHILD_TRIS;
        if (child->desc.has_triangles && rec_is_translated_copy(&t->leafs[4 * (size_t)q])) s->desc.shear_at_top = 1;
    }
    // The builder's per-geometry references are no longer needed.
    crt_builder_free(b);

    if (!host_only) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = CRT_OK;
        std::vector<DevNode> reordered;  // device copy only: the host tree keeps the reference's DFS order
        s->top_bytes = top_levels_first(t->nodes, g_dev.l2_persist_bytes, reordered);
        const DevNode* node_src = reordered.empty() ? t->nodes.data() : reordered.data();
        if ((rc = upload(&s->d_nodes, node_src, t->nodes.size())) ||
            (rc = upload(&s->d_leafs, t->leafs.data(), t->leafs.size())) ||
            (rc = upload(&s->d_prim_ids, t->prim_ids.data(), t->prim_ids.size())) ||
            (rc = upload(&s->d_normals, t->normals.data(), t->normals.size())) ||
            (rc = upload(&s->d_motion, t->motion.data(), t->motion.size()))) {
            delete t;
            crt_scene_release(s);
            return rc;
        }
        if (s->d_nodes) CK(mark_empty_lanes((DevNode*)s->d_nodes, t->nodes.size(), 0));
        s->desc.nodes = (const DevNode*)s->d_nodes;
        s->desc.leafs = (const U4*)s->d_leafs;
        s->desc.prim_ids = (const U2*)s->d_prim_ids;
        s->desc.normals = (const float*)s->d_normals;
        s->desc.motion = (const MotionXform*)s->d_motion;
        s->desc.n_nodes = (uint32_t)t->nodes.size();
        s->desc.has_triangles = s->breakdown.triangles != 0;
        s->desc.shear_at_top |= s->desc.has_triangles;
        if ((rc = upload((void**)&s->d_desc, &s->desc, 1))) {
            delete t;
            crt_scene_release(s);
            return rc;
        }
        s->info.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    s->footprint.bvh_nodes = t->nodes.size() * sizeof(DevNode);
    s->footprint.leaf_records = t->leafs.size() * sizeof(uint32_t);
    s->footprint.prim_table = t->prim_ids.size() * 4 + t->normals.size() * 4 + t->motion.size() * sizeof(MotionXform);
    s->footprint.descriptors = sizeof(DevScene);

    if (flags & CRT_COMMIT_KEEP_HOST_TREE) {
        s->host = t;
    } else {
        delete t;
    }
    *out = s;
    return CRT_OK;
}
int crt_commit(crt_builder* b, crt_scene** out) { return crt_commit_ex(b, 0, 0, out); }

void crt_scene_retain(crt_scene* s) {
    if (s) s->refs.fetch_add(1);
}
void crt_scene_release(crt_scene* s) {
    if (s && s->refs.fetch_sub(1) == 1) delete s;
}

// ---- scene queries ---------------------------------------------------------
int crt_bounds(const crt_scene* s, float o[6], int* non_empty) {
    if (!s || !o || !non_empty) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *non_empty = s->non_empty ? 1 : 0;
    o[0] = s->root.lo.x, o[1] = s->root.lo.y, o[2] = s->root.lo.z;
    o[3] = s->root.hi.x, o[4] = s->root.hi.y, o[5] = s->root.hi.z;
    return CRT_OK;
}
int crt_geometry_count(const crt_scene* s, uint32_t* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = s->n_geoms;
    return CRT_OK;
}
int crt_has_motion(const crt_scene* s, int* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = s->has_motion ? 1 : 0;
    return CRT_OK;
}
int crt_primitive_count(const crt_scene* s, uint64_t* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = s->n_prims;
    return CRT_OK;
}
int crt_primitive_breakdown_of(const crt_scene* s, crt_primitive_breakdown* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = s->breakdown;
    return CRT_OK;
}
static void unique_into(const crt_scene* s, std::unordered_set<const crt_scene*>& seen, crt_primitive_breakdown& acc) {
    acc.triangles += s->breakdown.triangles;
    acc.spheres += s->breakdown.spheres;
    acc.curve_segments += s->breakdown.curve_segments;
    acc.cubic_curve_spans += s->breakdown.cubic_curve_spans;
    acc.instances += s->breakdown.instances;
    for (const crt_scene* c : s->children)
        if (seen.insert(c).second) unique_into(c, seen, acc);
}
int crt_unique_primitive_breakdown(const crt_scene* s, crt_primitive_breakdown* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = crt_primitive_breakdown{};
    std::unordered_set<const crt_scene*> seen;
    unique_into(s, seen, *out);
    return CRT_OK;
}
int crt_primitive_extents(const crt_scene* s, uint64_t* count, float* scene_diag, float* mean_diag, float* max_diag) {
    if (!s) return fail(CRT_ERR_INVALID_ARG, "null argument");
    float dx = s->root.hi.x - s->root.lo.x, dy = s->root.hi.y - s->root.lo.y, dz = s->root.hi.z - s->root.lo.z;
    float diag = s->non_empty ? sqrtf((dx * dx + dy * dy) + dz * dz) : 0.0f;
    if (count) *count = s->n_prims;
    if (scene_diag) *scene_diag = diag;
    if (mean_diag) *mean_diag = s->n_prims ? s->extent_sum / (float)s->n_prims : 0.0f;
    if (max_diag) *max_diag = s->extent_max;
    return CRT_OK;
}
static void footprint_into(const crt_scene* s, std::unordered_set<const crt_scene*>& seen, crt_memory_footprint& acc) {
    acc.bvh_nodes += s->footprint.bvh_nodes;
    acc.leaf_records += s->footprint.leaf_records;
    acc.prim_table += s->footprint.prim_table;
    acc.descriptors += s->footprint.descriptors;
    for (const crt_scene* c : s->children)
        if (seen.insert(c).second) footprint_into(c, seen, acc);
}
int crt_memory_footprint_of(const crt_scene* s, crt_memory_footprint* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = crt_memory_footprint{};
    std::unordered_set<const crt_scene*> seen;
    footprint_into(s, seen, *out);
    return CRT_OK;
}
int crt_build_info_of(const crt_scene* s, crt_build_info* out) {
    if (!s || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = s->info;
    return CRT_OK;
}

// ---- replicating a commit (multi-GPU: build once, share) --------------------
// A committed scene as one flat blob: header (host-side metadata + array sizes) | nodes | leaf records | prim ids | normals
// | motion | instance-record offsets | instance-record child indices. Instanced scenes are referenced by their index in the
// scene's list of distinct children (first-appearance order over the attached geometries — the same on every process that
// attaches the same geometries), so each scene of an instancing hierarchy travels on its own and the importer hands in the
// handles of the children it imported before. The blob lives in host or device memory; a device blob goes straight through
// an NCCL broadcast and is unpacked with device-to-device copies.
namespace {
struct BlobHeader {
    uint32_t magic, version;
    uint64_t total_bytes;
    uint64_t n_nodes, n_leaf_quads, n_prim_ids, n_normal_floats, n_motion, n_children, n_inst;
    uint64_t top_bytes;
    // host-side metadata of crt_scene
    uint32_t non_empty, has_motion, n_geoms, instance_depth, features, has_triangles, shear_at_top, pad_;
    uint64_t n_prims;
    float root[6];
    float extent_sum, extent_max;
    crt_primitive_breakdown breakdown;
    crt_build_info info;
    crt_memory_footprint footprint;
};
constexpr uint32_t kBlobMagic = 0x42545243u;  // "CRTB"
constexpr uint32_t kBlobVersion = 2;
size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }
struct BlobLayout {
    size_t nodes, leafs, prim_ids, normals, motion, inst_q, inst_child, total;
    explicit BlobLayout(const BlobHeader& h) {
        size_t o = pad16(sizeof(BlobHeader));
        nodes = o, o += pad16(h.n_nodes * sizeof(DevNode));
        // nodes are 128-byte records; keep them 128-byte aligned inside the blob as well (cheap, and lets a consumer map them)
        leafs = o, o += pad16(h.n_leaf_quads * 16);
        prim_ids = o, o += pad16(h.n_prim_ids * 8);
        normals = o, o += pad16(h.n_normal_floats * 4);
        motion = o, o += pad16(h.n_motion * sizeof(MotionXform));
        inst_q = o, o += pad16(h.n_inst * 4);
        inst_child = o, o += pad16(h.n_inst * 4);
        total = o;
    }
};
BlobHeader header_of(const crt_scene* s) {
    BlobHeader h{};
    h.magic = kBlobMagic, h.version = kBlobVersion;
    h.n_nodes = s->n_nodes, h.n_leaf_quads = s->n_leaf_quads, h.n_prim_ids = s->n_prim_ids, h.n_normal_floats = s->n_normal_floats;
    h.n_motion = s->n_motion, h.n_children = s->children.size(), h.n_inst = s->inst_q.size(), h.top_bytes = s->top_bytes;
    h.non_empty = s->non_empty, h.has_motion = s->has_motion, h.n_geoms = s->n_geoms, h.instance_depth = s->instance_depth;
    h.features = s->features, h.has_triangles = s->desc.has_triangles, h.shear_at_top = s->desc.shear_at_top, h.pad_ = 0, h.n_prims = s->n_prims;
    h.root[0] = s->root.lo.x, h.root[1] = s->root.lo.y, h.root[2] = s->root.lo.z, h.root[3] = s->root.hi.x, h.root[4] = s->root.hi.y, h.root[5] = s->root.hi.z;
    h.extent_sum = s->extent_sum, h.extent_max = s->extent_max;
    h.breakdown = s->breakdown, h.info = s->info, h.footprint = s->footprint;
    h.total_bytes = BlobLayout(h).total;
    return h;
}
__global__ void patch_instances_kernel(U4* leafs, const uint32_t* __restrict__ q, const uint32_t* __restrict__ child, const uint64_t* __restrict__ ptrs,
                                       uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* c = ptrs + 3 * (size_t)child[i];  // descriptor, nodes, leafs of the child in THIS process
    U4& q4 = leafs[q[i] + 4];
    q4.x = (uint32_t)(c[0] & 0xFFFFFFFFu), q4.y = (uint32_t)(c[0] >> 32);
    U4& q5 = leafs[q[i] + 5];
    q5.x = (uint32_t)(c[1] & 0xFFFFFFFFu), q5.y = (uint32_t)(c[1] >> 32), q5.z = (uint32_t)(c[2] & 0xFFFFFFFFu), q5.w = (uint32_t)(c[2] >> 32);
}
}  // namespace

int crt_scene_export_size(const crt_scene* s, size_t* bytes) {
    if (!s || !bytes) return fail(CRT_ERR_INVALID_ARG, "null argument");
    if (s->host_only) return fail(CRT_ERR_NO_DEVICE, "scene was committed host-only: nothing to export");
    *bytes = header_of(s).total_bytes;
    return CRT_OK;
}

int crt_scene_export(const crt_scene* s, void* dst, size_t bytes, int dst_on_device, void* stream_) {
    if (!s || !dst) return fail(CRT_ERR_INVALID_ARG, "null argument");
    if (s->host_only) return fail(CRT_ERR_NO_DEVICE, "scene was committed host-only: nothing to export");
    const BlobHeader h = header_of(s);
    if (bytes < h.total_bytes) return fail(CRT_ERR_INVALID_ARG, "crt_scene_export: buffer smaller than crt_scene_export_size");
    const BlobLayout L(h);
    cudaStream_t st = (cudaStream_t)stream_;
    const cudaMemcpyKind from_dev = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const cudaMemcpyKind from_host = dst_on_device ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost;
    char* d = (char*)dst;
    CK(cudaMemcpyAsync(d, &h, sizeof h, from_host, st));
    if (h.n_nodes) CK(cudaMemcpyAsync(d + L.nodes, s->d_nodes, h.n_nodes * sizeof(DevNode), from_dev, st));
    if (h.n_leaf_quads) CK(cudaMemcpyAsync(d + L.leafs, s->d_leafs, h.n_leaf_quads * 16, from_dev, st));
    if (h.n_prim_ids) CK(cudaMemcpyAsync(d + L.prim_ids, s->d_prim_ids, h.n_prim_ids * 8, from_dev, st));
    if (h.n_normal_floats) CK(cudaMemcpyAsync(d + L.normals, s->d_normals, h.n_normal_floats * 4, from_dev, st));
    if (h.n_motion) CK(cudaMemcpyAsync(d + L.motion, s->d_motion, h.n_motion * sizeof(MotionXform), from_dev, st));
    if (h.n_inst) {
        CK(cudaMemcpyAsync(d + L.inst_q, s->inst_q.data(), h.n_inst * 4, from_host, st));
        CK(cudaMemcpyAsync(d + L.inst_child, s->inst_child.data(), h.n_inst * 4, from_host, st));
    }
    CK(cudaStreamSynchronize(st));  // the header and the two index lists are read from this call's stack / the scene's vectors
    return CRT_OK;
}

int crt_scene_import(const void* src, size_t bytes, int src_on_device, crt_scene* const* children, size_t n_children, crt_scene** out) {
    if (!src || !out || (n_children && !children)) return fail(CRT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    int rc = ensure_device();
    if (rc) return rc;
    if (bytes < sizeof(BlobHeader)) return fail(CRT_ERR_INVALID_ARG, "crt_scene_import: truncated blob");
    BlobHeader h;
    CK(cudaMemcpy(&h, src, sizeof h, src_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    if (h.magic != kBlobMagic || h.version != kBlobVersion) return fail(CRT_ERR_INVALID_ARG, "crt_scene_import: not a scene blob of this library version");
    const BlobLayout L(h);
    if (h.total_bytes != L.total || bytes < L.total) return fail(CRT_ERR_INVALID_ARG, "crt_scene_import: truncated blob");
    if (h.n_children != n_children) return fail(CRT_ERR_INVALID_ARG, "crt_scene_import: the blob instances a different number of distinct scenes than were handed in");
    for (size_t c = 0; c < n_children; ++c)
        if (!children[c] || children[c]->host_only) return fail(CRT_ERR_INVALID_ARG, "crt_scene_import: child scene missing or host-only");
    const char* b = (const char*)src;
    const cudaMemcpyKind to_dev = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    crt_scene* s = new crt_scene();
    for (size_t c = 0; c < n_children; ++c) {
        crt_scene_retain(children[c]);
        s->children.push_back(children[c]);
    }
    s->non_empty = h.non_empty != 0, s->has_motion = h.has_motion != 0, s->n_geoms = h.n_geoms, s->instance_depth = h.instance_depth;
    s->features = h.features, s->n_prims = h.n_prims, s->top_bytes = (size_t)h.top_bytes;
    s->root.lo.x = h.root[0], s->root.lo.y = h.root[1], s->root.lo.z = h.root[2], s->root.hi.x = h.root[3], s->root.hi.y = h.root[4], s->root.hi.z = h.root[5];
    s->extent_sum = h.extent_sum, s->extent_max = h.extent_max;
    s->breakdown = h.breakdown, s->info = h.info, s->footprint = h.footprint;
    s->n_nodes = h.n_nodes, s->n_leaf_quads = h.n_leaf_quads, s->n_prim_ids = h.n_prim_ids, s->n_normal_floats = h.n_normal_floats, s->n_motion = h.n_motion;
    auto fail_free = [&](int code) {
        crt_scene_release(s);
        return code;
    };
    auto take = [&](void** dst, size_t off, size_t nbytes) -> int {
        *dst = nullptr;
        if (!nbytes) return CRT_OK;
        cudaError_t e = cudaMalloc(dst, nbytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? CRT_ERR_OOM : CRT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
        e = cudaMemcpy(*dst, b + off, nbytes, to_dev);
        return e == cudaSuccess ? CRT_OK : fail_cuda(e, "crt_scene_import copy");
    };
    if ((rc = take(&s->d_nodes, L.nodes, h.n_nodes * sizeof(DevNode))) || (rc = take(&s->d_leafs, L.leafs, h.n_leaf_quads * 16)) ||
        (rc = take(&s->d_prim_ids, L.prim_ids, h.n_prim_ids * 8)) || (rc = take(&s->d_normals, L.normals, h.n_normal_floats * 4)) ||
        (rc = take(&s->d_motion, L.motion, h.n_motion * sizeof(MotionXform))))
        return fail_free(rc);
    if (h.n_inst) {
        s->inst_q.resize(h.n_inst);
        s->inst_child.resize(h.n_inst);
        const cudaMemcpyKind to_host = src_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
        cudaError_t e = cudaMemcpy(s->inst_q.data(), b + L.inst_q, h.n_inst * 4, to_host);
        if (e == cudaSuccess) e = cudaMemcpy(s->inst_child.data(), b + L.inst_child, h.n_inst * 4, to_host);
        if (e != cudaSuccess) return fail_free(fail_cuda(e, "crt_scene_import copy"));
        for (uint32_t c : s->inst_child)
            if (c >= n_children) return fail_free(fail(CRT_ERR_INVALID_ARG, "crt_scene_import: corrupt child index"));
        // the instance records carry addresses of the exporting process: rewrite them with this process' children
        std::vector<uint64_t> ptrs(3 * n_children);
        for (size_t c = 0; c < n_children; ++c) {
            ptrs[3 * c] = (uint64_t)(uintptr_t)children[c]->d_desc;
            ptrs[3 * c + 1] = (uint64_t)(uintptr_t)children[c]->d_nodes;
            ptrs[3 * c + 2] = (uint64_t)(uintptr_t)children[c]->d_leafs;
        }
        uint32_t *d_q = nullptr, *d_c = nullptr;
        uint64_t* d_p = nullptr;
        e = cudaMalloc(&d_q, h.n_inst * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_c, h.n_inst * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_p, ptrs.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(d_q, s->inst_q.data(), h.n_inst * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_c, s->inst_child.data(), h.n_inst * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_p, ptrs.data(), ptrs.size() * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            patch_instances_kernel<<<(unsigned)((h.n_inst + 255) / 256), 256>>>((U4*)s->d_leafs, d_q, d_c, d_p, (uint32_t)h.n_inst);
            e = cudaDeviceSynchronize();
        }
        cudaFree(d_q), cudaFree(d_c), cudaFree(d_p);
        if (e != cudaSuccess) return fail_free(fail_cuda(e, "crt_scene_import patch"));
    }
    s->desc.nodes = (const DevNode*)s->d_nodes;
    s->desc.leafs = (const U4*)s->d_leafs;
    s->desc.prim_ids = (const U2*)s->d_prim_ids;
    s->desc.normals = (const float*)s->d_normals;
    s->desc.motion = (const MotionXform*)s->d_motion;
    s->desc.n_nodes = (uint32_t)h.n_nodes;
    s->desc.has_triangles = h.has_triangles;
    s->desc.shear_at_top = h.shear_at_top;
    if ((rc = upload((void**)&s->d_desc, &s->desc, 1))) return fail_free(rc);
    *out = s;
    return CRT_OK;
}

// ---- batched queries -------------------------------------------------------
static int launch_query(const crt_scene* s, bool any_hit, size_t n, const crt_ray* rays, const float* time,
                        const uint32_t* mask, crt_hit* hits, uint8_t* flag, cudaStream_t stream) {
    if (!s) return fail(CRT_ERR_INVALID_ARG, "null scene");
    if (s->host_only) return fail(CRT_ERR_NO_DEVICE, "scene was committed host-only (test mode): no device data to query");
    if (n == 0) return CRT_OK;
    if (!rays || (any_hit ? !flag : !hits)) return fail(CRT_ERR_INVALID_ARG, "null ray or output buffer");
    int rc = ensure_device();
    if (rc) return rc;
    const size_t kMaxPerLaunch = (size_t)1 << 30;  // 32-bit ray counter with head-room for overshoot
    int per_sm = trace_blocks_per_sm(any_hit, s->features, s->instance_depth);
    const bool hinted = hint_top_levels(s, stream);
    struct Unhint {
        bool on;
        cudaStream_t st;
        ~Unhint() {
            if (on) clear_top_levels_hint(st);
        }
    } unhint{hinted, stream};
    for (size_t off = 0; off < n; off += kMaxPerLaunch) {
        size_t cnt = std::min(kMaxPerLaunch, n - off);
        TraceParams p;
        p.scene = s->desc;
        p.rays = (const float4*)(rays + off);
        p.time = time ? time + off : nullptr;
        p.mask = mask ? mask + off : nullptr;
        p.hits = hits ? (float4*)(hits + off) : nullptr;
        p.flag = flag ? flag + off : nullptr;
        p.counter = g_dev.counters + (g_dev.next_counter.fetch_add(1) % kCounterRing);
        p.status = g_dev.status;
        p.n = (uint32_t)cnt;
        p.n_dev = nullptr;
        p.mask_uniform = CRT_MASK_ALL;
        p.stats = g_dev.count_traversal.load() ? g_dev.trace_stats : nullptr;
        int grid = g_dev.sm_count * per_sm;
        size_t need = (cnt + CRT_BLOCK - 1) / CRT_BLOCK;
        if ((size_t)grid > need) grid = (int)need;
        CK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream));
        CK(launch_trace(any_hit, s->features, s->instance_depth, p, grid, stream));
        g_dev.launches.fetch_add(1);
    }
    return CRT_OK;
}

int crt_intersect_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                    crt_hit* hits, uint8_t* front_face, void* stream) {
    return launch_query(s, false, n, rays, time, mask, hits, front_face, (cudaStream_t)stream);
}
int crt_occluded_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                   uint8_t* occluded, void* stream) {
    return launch_query(s, true, n, rays, time, mask, nullptr, occluded, (cudaStream_t)stream);
}

uint64_t crt_launch_count(void) { return g_dev.launches.load(); }
int crt_l2_persist_info(uint64_t* set_aside_bytes, uint64_t* windows_applied, uint64_t* windows_refused) {
    if (set_aside_bytes) *set_aside_bytes = g_dev.l2_persist_bytes;
    if (windows_applied) *windows_applied = g_dev.l2_applied.load();
    if (windows_refused) *windows_refused = g_dev.l2_refused.load();
    return CRT_OK;
}

// Counting traversal (bvh.rs:31-136 `stats`): while enabled, every traversal launch runs the counting
// build of its kernel. crt_trace_counters_read synchronises the device and returns {rays, nodes fetched, popped entries
// culled by entry distance, leaves entered, triangle records tested, other primitive records tested, instance descents}.
int crt_trace_counters_enable(int on) {
    int rc = ensure_device();
    if (rc) return rc;
    g_dev.count_traversal.store(on != 0);
    return CRT_OK;
}
int crt_trace_counters_read(uint64_t out[7], int reset) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    CK(cudaDeviceSynchronize());
    unsigned long long v[8];
    CK(cudaMemcpy(v, g_dev.trace_stats, sizeof v, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 7; ++k) out[k] = v[k];
    if (reset) CK(cudaMemset(g_dev.trace_stats, 0, sizeof v));
    return CRT_OK;
}

// Sticky device-side error bits (bit 0: a traversal stack overflowed and an
// entry was dropped). Synchronises the device.
int crt_device_status(uint32_t* bits) {
    int rc = ensure_device();
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    unsigned int v = 0;
    CK(cudaMemcpy(&v, g_dev.status, sizeof v, cudaMemcpyDeviceToHost));
    if (bits) *bits = v;
    return CRT_OK;
}

int crt_device_status_reset(void) {
    int rc = ensure_device();
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(g_dev.status, 0, sizeof(unsigned int)));
    return CRT_OK;
}

// ---- host-buffer path ------------------------------------------------------
namespace {
constexpr int kSlots = 3;
// rays per pipeline stage; CRT_HOST_CHUNK_LOG2 overrides (tuning only)
static size_t host_chunk() {
    static size_t v = [] {
        const char* e = getenv("CRT_HOST_CHUNK_LOG2");
        int l = e ? atoi(e) : 20;  // 2^20 measured best of 2^19..2^23 on B200 (PCIe-bound: 52 GB/s H2D)
        if (l < 12 || l > 26) l = 20;
        return (size_t)1 << l;
    }();
    return v;
}
#define kChunk host_chunk()
struct HostPipe {
    std::mutex mu;
    bool ready = false;
    cudaStream_t stream[kSlots];
    cudaEvent_t done[kSlots];
    // device scratch per slot
    crt_ray* d_rays[kSlots];
    crt_hit* d_hits[kSlots];
    float* d_time[kSlots];
    uint32_t* d_mask[kSlots];
    uint8_t* d_flag[kSlots];
    // pinned staging per slot (only for pageable user memory)
    crt_ray* h_rays[kSlots] = {};
    crt_hit* h_hits[kSlots];
    float* h_time[kSlots];
    uint32_t* h_mask[kSlots];
    uint8_t* h_flag[kSlots];
} g_pipe;

int ensure_pipe() {
    if (g_pipe.ready) return CRT_OK;
    for (int i = 0; i < kSlots; ++i) {
        CK(cudaStreamCreateWithFlags(&g_pipe.stream[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&g_pipe.done[i], cudaEventDisableTiming));
        CK(cudaMalloc(&g_pipe.d_rays[i], kChunk * sizeof(crt_ray)));
        CK(cudaMalloc(&g_pipe.d_hits[i], kChunk * sizeof(crt_hit)));
        CK(cudaMalloc(&g_pipe.d_time[i], kChunk * sizeof(float)));
        CK(cudaMalloc(&g_pipe.d_mask[i], kChunk * sizeof(uint32_t)));
        CK(cudaMalloc(&g_pipe.d_flag[i], kChunk));
    }
    g_pipe.ready = true;
    return CRT_OK;
}
// Pinned staging is only needed when the caller's buffers are pageable: allocate it on first such call.
int ensure_staging() {
    if (g_pipe.h_rays[0]) return CRT_OK;
    for (int i = 0; i < kSlots; ++i) {
        CK(cudaMallocHost(&g_pipe.h_rays[i], kChunk * sizeof(crt_ray)));
        CK(cudaMallocHost(&g_pipe.h_hits[i], kChunk * sizeof(crt_hit)));
        CK(cudaMallocHost(&g_pipe.h_time[i], kChunk * sizeof(float)));
        CK(cudaMallocHost(&g_pipe.h_mask[i], kChunk * sizeof(uint32_t)));
        CK(cudaMallocHost(&g_pipe.h_flag[i], kChunk));
    }
    return CRT_OK;
}
bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct Pending {  // staged outputs still to be copied out of pinned memory
    bool used = false;
    size_t off = 0, cnt = 0;
};

int host_query(const crt_scene* s, bool any_hit, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
               crt_hit* hits, uint8_t* flag) {
    if (!s) return fail(CRT_ERR_INVALID_ARG, "null scene");
    if (s->host_only) return fail(CRT_ERR_NO_DEVICE, "scene was committed host-only (test mode): no device data to query");
    if (n == 0) return CRT_OK;
    if (!rays || (any_hit ? !flag : !hits)) return fail(CRT_ERR_INVALID_ARG, "null ray or output buffer");
    int rc = ensure_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_pipe.mu);
    if ((rc = ensure_pipe())) return rc;
    const bool in_pinned = is_pinned(rays) && is_pinned(time) && is_pinned(mask);
    const bool out_pinned = is_pinned(hits) && is_pinned(flag);
    if ((!in_pinned || !out_pinned) && (rc = ensure_staging())) return rc;
    Pending pend[kSlots];
    auto drain = [&](int sl) -> int {
        if (!pend[sl].used) return CRT_OK;
        CK(cudaEventSynchronize(g_pipe.done[sl]));
        if (!out_pinned) {
            if (hits) memcpy(hits + pend[sl].off, g_pipe.h_hits[sl], pend[sl].cnt * sizeof(crt_hit));
            if (flag) memcpy(flag + pend[sl].off, g_pipe.h_flag[sl], pend[sl].cnt);
        }
        pend[sl].used = false;
        return CRT_OK;
    };
    // The copies are the bottleneck (PCIe), so what a call adds to the pure copy time is its drain: the last chunk's kernel and
    // download run after the last upload has finished. The tail is therefore tapered — once less than two chunks remain, each
    // chunk takes half of what is left (in units of kTail rays) — which shrinks the drain from a 2^20-ray kernel + 32 MB
    // download to a 2^16-ray one.
    constexpr size_t kTail = (size_t)1 << 16;
    static const bool taper = getenv("CRT_HOST_NO_TAPER") == nullptr;  // tuning only
    size_t chunk_no = 0;
    for (size_t off = 0, cnt = 0; off < n; off += cnt, ++chunk_no) {
        int sl = (int)(chunk_no % kSlots);
        const size_t rem = n - off;
        cnt = std::min(kChunk, rem);
        if (taper && rem < 2 * kChunk && rem > kTail) cnt = std::min(cnt, std::max((rem / 2 + kTail - 1) / kTail * kTail, kTail));
        if ((rc = drain(sl))) return rc;
        cudaStream_t st = g_pipe.stream[sl];
        const crt_ray* src_r = rays + off;
        const float* src_t = time ? time + off : nullptr;
        const uint32_t* src_m = mask ? mask + off : nullptr;
        if (!in_pinned) {
            memcpy(g_pipe.h_rays[sl], src_r, cnt * sizeof(crt_ray));
            src_r = g_pipe.h_rays[sl];
            if (time) memcpy(g_pipe.h_time[sl], src_t, cnt * sizeof(float)), src_t = g_pipe.h_time[sl];
            if (mask) memcpy(g_pipe.h_mask[sl], src_m, cnt * sizeof(uint32_t)), src_m = g_pipe.h_mask[sl];
        }
        CK(cudaMemcpyAsync(g_pipe.d_rays[sl], src_r, cnt * sizeof(crt_ray), cudaMemcpyHostToDevice, st));
        if (time) CK(cudaMemcpyAsync(g_pipe.d_time[sl], src_t, cnt * sizeof(float), cudaMemcpyHostToDevice, st));
        if (mask) CK(cudaMemcpyAsync(g_pipe.d_mask[sl], src_m, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        rc = launch_query(s, any_hit, cnt, g_pipe.d_rays[sl], time ? g_pipe.d_time[sl] : nullptr,
                          mask ? g_pipe.d_mask[sl] : nullptr, any_hit ? nullptr : g_pipe.d_hits[sl],
                          (any_hit || flag) ? g_pipe.d_flag[sl] : nullptr, st);
        if (rc) return rc;
        if (hits)
            CK(cudaMemcpyAsync(out_pinned ? hits + off : g_pipe.h_hits[sl], g_pipe.d_hits[sl], cnt * sizeof(crt_hit),
                               cudaMemcpyDeviceToHost, st));
        if (flag)
            CK(cudaMemcpyAsync(out_pinned ? flag + off : g_pipe.h_flag[sl], g_pipe.d_flag[sl], cnt, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(g_pipe.done[sl], st));
        pend[sl] = Pending{true, off, cnt};
    }
    for (int sl = 0; sl < kSlots; ++sl)
        if ((rc = drain(sl))) return rc;
    return CRT_OK;
}
}  // namespace

int crt_intersect_host_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                         crt_hit* hits, uint8_t* front_face) {
    return host_query(s, false, n, rays, time, mask, hits, front_face);
}
int crt_occluded_host_n(const crt_scene* s, size_t n, const crt_ray* rays, const float* time, const uint32_t* mask,
                        uint8_t* occluded) {
    return host_query(s, true, n, rays, time, mask, nullptr, occluded);
}
int crt_intersect1(const crt_scene* s, const crt_ray* ray, float time, uint32_t mask, crt_hit* out, int* front_face,
                   int* hit) {
    if (!ray || !out) return fail(CRT_ERR_INVALID_ARG, "null argument");
    uint8_t ff = 0;
    int rc = host_query(s, false, 1, ray, &time, &mask, out, &ff);
    if (rc) return rc;
    if (front_face) *front_face = ff;
    if (hit) *hit = out->geom_id != CRT_INVALID_ID;
    return CRT_OK;
}
int crt_occluded1(const crt_scene* s, const crt_ray* ray, float time, uint32_t mask, int* occluded) {
    if (!ray || !occluded) return fail(CRT_ERR_INVALID_ARG, "null argument");
    uint8_t o = 0;
    int rc = host_query(s, true, 1, ray, &time, &mask, nullptr, &o);
    if (rc) return rc;
    *occluded = o;
    return CRT_OK;
}

// ---- debug export ----------------------------------------------------------
int crt_debug_tree_sizes(const crt_scene* s, uint64_t out4[4]) {
    if (!s || !out4) return fail(CRT_ERR_INVALID_ARG, "null argument");
    if (!s->host) return fail(CRT_ERR_INVALID_ARG, "scene was not committed with CRT_COMMIT_KEEP_HOST_TREE");
    export_sizes(*s->host, out4);
    return CRT_OK;
}
int crt_debug_host_array(const crt_scene* s, int which, uint32_t* out, uint64_t capacity_words, uint64_t* n_words) {
    if (!s || !n_words) return fail(CRT_ERR_INVALID_ARG, "null argument");
    if (!s->host) return fail(CRT_ERR_INVALID_ARG, "scene was not committed with CRT_COMMIT_KEEP_HOST_TREE");
    const HostTree& t = *s->host;
    const void* src = nullptr;
    uint64_t words = 0;
    switch (which) {
        case 0: src = t.nodes.data(), words = t.nodes.size() * (sizeof(DevNode) / 4); break;
        case 1: src = t.leafs.data(), words = t.leafs.size(); break;
        case 2: src = t.motion.data(), words = t.motion.size() * (sizeof(MotionXform) / 4); break;
        case 3: src = t.instance_record_q.data(), words = t.instance_record_q.size(); break;
        default: return fail(CRT_ERR_INVALID_ARG, "crt_debug_host_array: which must be 0..3");
    }
    *n_words = words;
    if (out && words) memcpy(out, src, (size_t)std::min(capacity_words, words) * 4);
    return CRT_OK;
}
int crt_debug_tree_export(const crt_scene* s, void* nodes, uint32_t* tri_counts, uint32_t* other_counts, uint32_t* refs) {
    if (!s || !nodes || !tri_counts || !other_counts || !refs) return fail(CRT_ERR_INVALID_ARG, "null argument");
    if (!s->host) return fail(CRT_ERR_INVALID_ARG, "scene was not committed with CRT_COMMIT_KEEP_HOST_TREE");
    export_tree(*s->host, nodes, tri_counts, other_counts, refs);
    return CRT_OK;
}

}  // extern "C"

// ---- library-internal interface (internal.h) -----------------------------------------------
namespace crt {
int set_error(int code, const std::string& msg) { return fail(code, msg); }
int set_cuda_error(cudaError_t e, const char* what) { return fail_cuda(e, what); }
int require_device() { return ensure_device(); }
int device_sm_count() { return g_dev.sm_count; }
unsigned int* device_status_word() { return g_dev.status; }
void count_launches(uint64_t n) { g_dev.launches.fetch_add(n); }

bool hint_scene_top_levels(const crt_scene* s, cudaStream_t stream) { return s && !s->host_only && hint_top_levels(s, stream); }
void clear_scene_hint(cudaStream_t stream) { clear_top_levels_hint(stream); }
void scene_device_memory_released(bool had_window) {
    // persisting lines of a freed tree must not outlive it
    if (had_window && g_dev.ready && g_dev.l2_persist_bytes) {
        if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
    }
}

int enqueue_queue_trace(const crt_scene* s, bool any_hit, const QueueTrace& q, cudaStream_t stream) {
    if (!s || s->host_only) return fail(CRT_ERR_NO_DEVICE, "scene has no device data");
    TraceParams p;
    p.scene = s->desc;
    p.rays = q.rays;
    p.time = q.time;
    p.mask = nullptr;
    p.mask_uniform = q.mask_uniform;
    p.hits = q.hits;
    p.flag = q.flag;
    p.counter = q.counter;
    p.status = g_dev.status;
    p.n = q.capacity;
    p.n_dev = q.n_dev;
    p.stats = g_dev.count_traversal.load() ? g_dev.trace_stats : nullptr;
    int grid = g_dev.sm_count * trace_blocks_per_sm(any_hit, s->features, s->instance_depth);
    size_t need = ((size_t)q.capacity + CRT_BLOCK - 1) / CRT_BLOCK;
    if ((size_t)grid > need) grid = (int)(need ? need : 1);
    CK(launch_trace(any_hit, s->features, s->instance_depth, p, grid, stream));
    g_dev.launches.fetch_add(1);
    return CRT_OK;
}
}  // namespace crt
