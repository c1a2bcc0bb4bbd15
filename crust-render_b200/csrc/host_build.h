// Host side of commit(): geometry expansion, the deterministic SAH/SBVH binary
// build, the BVH4 collapse and the flattening into the device layout
// (crt_layout.h). No CUDA in this translation unit.
//
// Behavioural contract = crates/crust-rt/src/{scene.rs:226-341, bvh.rs:300-327,
// 811-1375, triangle.rs:367-429, aabb.rs:47-58, prim.rs:285-319} of the
// reference: same split decisions, same order-preserving partitions, same DFS
// numbering, so the tree (and therefore every tie-break during traversal) is
// the reference's tree. Arithmetic is plain IEEE f32 without FMA contraction
// (this file is compiled with -ffp-contract=off).
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "crt.h"
#include "crt_layout.h"

struct crt_scene;

namespace crt {

// std::allocator whose value-initialisation is default-initialisation: resize() of the big flattened arrays leaves the new
// memory untouched, so its pages are first touched by the threads that fill them (in parallel) instead of being zeroed by one.
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = DefaultInitAlloc<U>;
    };
    using std::allocator<T>::allocator;
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0)
            ::new ((void*)p) U;
        else
            ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
using WordVec = std::vector<uint32_t, DefaultInitAlloc<uint32_t>>;

struct Vec3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct Box {
    Vec3 lo, hi;
};
struct Xform {  // glam Affine3A column order
    Vec3 x, y, z, t;
};

enum GeomType : uint8_t { G_MESH, G_SPHERE, G_CURVES, G_CUBICS, G_INSTANCE };

struct Geom {
    GeomType type = G_MESH;
    uint32_t mask = 0xFFFFFFFFu;
    std::vector<float> verts;       // 3 per vertex
    std::vector<uint32_t> indices;  // 3 per triangle
    std::vector<float> normals;     // 3 per vertex (optional)
    bool has_normals = false;
    float sphere[4] = {0, 0, 0, 0};  // center, radius
    std::vector<float> segs;         // curves: 8 floats each; cubics: 14 floats each
    crt_scene* child = nullptr;      // retained
    Xform l2w{}, l2w_end{};
    bool has_end = false;
};

enum PrimKind : uint8_t { P_TRI = 0, P_SPHERE = 1, P_CURVE = 2, P_CUBIC = 3, P_INSTANCE = 4 };

struct InstancePrim {
    crt_scene* child;
    Xform l2w, w2l, l2w_end;
    bool has_end;
    Box bounds;
};

// Expanded primitives, stored once each in input order (bvh.rs:272-273).
struct PrimStore {
    std::vector<uint8_t> kind;
    std::vector<uint32_t> geom_id, prim_id, mask, data;  // data: index into the per-kind arrays
    std::vector<float> tri;                              // 9 per triangle
    std::vector<float> tri_normals;                      // 9 per triangle that has normals, else absent
    std::vector<uint32_t> tri_normal_slot;               // per triangle: index into tri_normals/9 or ~0
    std::vector<float> sphere;                           // 4 per sphere
    std::vector<float> curve;                            // 8 per segment
    std::vector<float> cubic;                            // 14 per span
    std::vector<InstancePrim> inst;
    size_t size() const { return kind.size(); }
    bool any_normals = false;
};

struct BinNode {  // bvh.rs:171-178
    Box bbox;
    uint32_t first_or_right, count;
};

struct HostTree {
    PrimStore prims;
    bool non_empty = false;
    Box root{};
    bool has_motion = false;
    uint32_t instance_depth = 0;  // deepest nesting below this scene
    // device-layout arrays (host copies; freed after upload unless kept)
    std::vector<DevNode, DefaultInitAlloc<DevNode>> nodes;
    WordVec leafs;  // 4 words per 16-byte unit
    std::vector<uint32_t> prim_ids;  // 2 words per primitive
    std::vector<float> normals;      // 9 per primitive index when any mesh has normals
    std::vector<MotionXform> motion;
    std::vector<uint32_t> instance_record_q;  // offsets (16-B units) of instance records: child pointers are patched at upload
    uint64_t n_binary = 0, n_leaves = 0, n_leaf_refs = 0;
    double build_seconds = 0;
    float extent_sum = 0, extent_max = 0;  // bvh.rs:335-345 primitive_extent_sum
};

// scene.rs:240-335: geometries -> primitives. Child scenes must expose bounds /
// motion / depth through the callbacks (kept abstract so this file stays CUDA-free).
struct ChildInfo {
    bool non_empty;
    Box bounds;
    bool has_motion;
    uint32_t instance_depth;
};
typedef ChildInfo (*ChildInfoFn)(crt_scene*);

// Builds everything up to (not including) the upload. Returns false with `err`
// set when the scene cannot be represented (nesting too deep).
bool build_host_tree(std::vector<Geom>& geoms, ChildInfoFn child_info, int threads, HostTree& out, const char** err);

// The non-mesh primitives of a scene (spheres, curve segments, cubic spans, instances) expanded exactly like
// build_host_tree expands them — same ids, same silent skips, same bounds — each with its padded box and its leaf record
// (crt_layout.h, without REC_LAST; instance records still without child addresses), in input order. For the device build,
// which expands the triangle meshes itself: `prim_index_base` = number of triangle slots that precede these primitives in
// the scene's id table. Mesh inputs of `geoms` are left intact; the others are consumed.
struct OtherPrims {
    size_t n = 0;
    std::vector<float> lo, hi;         // 4 floats per primitive (lo.w = 1: valid)
    std::vector<uint32_t> ids;         // geom_id, prim_id per primitive
    std::vector<uint32_t> rec_off;     // n + 1 offsets into `recs`, in 16-byte units
    WordVec recs;        // 4 words per 16-byte unit
    std::vector<uint32_t> inst_prim;   // which of the primitives are instances ...
    std::vector<crt_scene*> inst_child;  // ... and of which scene
    std::vector<MotionXform> motion;
    crt_primitive_breakdown counts{};
    bool has_motion = false;
    uint32_t instance_depth = 0;
    float extent_sum = 0, extent_max = 0;
};
bool expand_other_prims(std::vector<Geom>& geoms, ChildInfoFn child_info, uint32_t prim_index_base, OtherPrims& out, const char** err);

// Canonical export for builder-equivalence tests (see crt.h crt_debug_*).
void export_sizes(const HostTree& t, uint64_t out4[4]);
void export_tree(const HostTree& t, void* nodes, uint32_t* tri_counts, uint32_t* other_counts, uint32_t* refs);

}  // namespace crt
