// Device-side BVH4 build for sm_100a: an alternative commit() path for
// triangle-only static scenes that constructs the whole acceleration structure on the GPU.
//
// The reference builds its tree on the host (crates/crust-rt/src/bvh.rs:300-327, 1059-1375: binned SAH +
// SBVH, then a collapse to BVH4); csrc/host_build.cpp restates that build exactly and stays the default,
// because hit ties ("later wins", bvh.rs:537-544) depend on the tree. This file is the fast path for when
// build time matters more than reproducing the reference's tree: seconds on the host become milliseconds.
// It emits the same device layout (crt_layout.h), so the traversal kernels are unchanged.
//
//   tri_setup      gather each triangle from its mesh (scene.rs:240-264: out-of-range indices skip the
//                  triangle but still consume a prim_id), padded AABB (aabb.rs:47-58), scene bounds
//   morton         63-bit Morton code of the AABB centre; invalid triangles sort to the end
//   radix sort     cub::DeviceRadixSort (library sort, like a library GEMM)
//   karras         binary radix tree over the sorted codes (Karras 2012), ties broken by position
//   refit          bottom-up AABBs, one atomic counter per internal node
//   mark / scan    which binary nodes become BVH4 nodes: each BVH4 node absorbs up to two of its
//                  descendants, always the child with the largest surface area first (the rule of
//                  bvh.rs:1328-1375); subtrees of <= 4 triangles become leaves; level-synchronous
//   emit           128-byte nodes + self-contained leaf records (triangles in sorted order), flags as the
//                  host flattener sets them (host_build.cpp emit_leaf)
// Every step is deterministic: the tree depends only on the input.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <chrono>
#include <vector>

#include "internal.h"

namespace crt {
namespace {

#define CKB(call)                                              \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return set_cuda_error(_e, #call); \
    } while (0)

constexpr int kB = 256;
constexpr float kPadDev = 1e-4f;  // aabb.rs:50
constexpr uint32_t kLeafMax = 4;  // bvh.rs:142-167 MIN_LEAF_PACKED: one Tri4 packet per leaf

struct MeshTable {            // per triangle mesh, device arrays of length n_meshes (+1 for the offsets)
    const uint32_t* tri_offset;   // first flat triangle index of mesh m; [n_meshes] = total
    const uint32_t* vert_offset;  // first vertex of mesh m in the concatenated vertex array
    const uint32_t* n_verts;
    const uint32_t* geom_id;
    const uint32_t* mask;
    uint32_t n_meshes;
};

struct BuildArrays {
    // per flat triangle
    float4 *v0, *v1, *v2;        // .w of v0 = mask bits
    float4 *lo, *hi;             // padded AABB
    uint2* prim_ids;             // (geom_id, prim_id)
    // sort
    unsigned long long *key_a, *key_b;
    uint32_t *idx_a, *idx_b;
    // binary radix tree over n sorted triangles: internal i in [0, n-1), leaf j = (n-1) + j
    int *left, *right, *parent;  // child / parent ids in the combined numbering; parent of the root = -1
    int *first, *last;           // sorted range of each internal node
    float4 *nlo, *nhi;           // AABB of every node in the combined numbering
    unsigned int* visits;
    // BVH4
    uint32_t *is_root4, *id4;
    int *frontier_a, *frontier_b;
    unsigned int* counters;      // [0] valid triangles, [1] next frontier size, [2] leaves, [8..13] bounds as ordered ints
};

__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned int u) {
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

__global__ void __launch_bounds__(kB) tri_setup_kernel(uint32_t n, MeshTable mt, const float* __restrict__ verts, const uint32_t* __restrict__ idx, BuildArrays A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    bool valid = false;
    if (i < n) {
        // mesh of this triangle: last m with tri_offset[m] <= i
        uint32_t a = 0, b = mt.n_meshes;
        while (b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if (__ldg(mt.tri_offset + m) <= i) a = m; else b = m;
        }
        const uint32_t m = a, nv = __ldg(mt.n_verts + m);
        const uint32_t i0 = __ldg(idx + 3 * (size_t)i), i1 = __ldg(idx + 3 * (size_t)i + 1), i2 = __ldg(idx + 3 * (size_t)i + 2);
        A.prim_ids[i] = make_uint2(__ldg(mt.geom_id + m), i - __ldg(mt.tri_offset + m));
        valid = i0 < nv && i1 < nv && i2 < nv;  // scene.rs:251-253
        if (valid) {
            const float* vb = verts + 3 * (size_t)__ldg(mt.vert_offset + m);
            float3 p0 = make_float3(vb[3 * (size_t)i0], vb[3 * (size_t)i0 + 1], vb[3 * (size_t)i0 + 2]);
            float3 p1 = make_float3(vb[3 * (size_t)i1], vb[3 * (size_t)i1 + 1], vb[3 * (size_t)i1 + 2]);
            float3 p2 = make_float3(vb[3 * (size_t)i2], vb[3 * (size_t)i2 + 1], vb[3 * (size_t)i2 + 2]);
            A.v0[i] = make_float4(p0.x, p0.y, p0.z, __uint_as_float(__ldg(mt.mask + m)));
            A.v1[i] = make_float4(p1.x, p1.y, p1.z, 0.0f);
            A.v2[i] = make_float4(p2.x, p2.y, p2.z, 0.0f);
            lo = make_float3(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z));
            hi = make_float3(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z));
            if (hi.x - lo.x < kPadDev) lo.x -= kPadDev, hi.x += kPadDev;  // aabb.rs:47-58
            if (hi.y - lo.y < kPadDev) lo.y -= kPadDev, hi.y += kPadDev;
            if (hi.z - lo.z < kPadDev) lo.z -= kPadDev, hi.z += kPadDev;
        }
        A.lo[i] = make_float4(lo.x, lo.y, lo.z, valid ? 1.0f : 0.0f);
        A.hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    // block reduction of the bounds and the valid count
    typedef cub::BlockReduce<float, kB> BR;
    __shared__ typename BR::TempStorage tmp;
    float r;
    r = BR(tmp).Reduce(lo.x, cub::Min()); __syncthreads(); if (threadIdx.x == 0) atomicMin(A.counters + 8, f2ord(r));
    r = BR(tmp).Reduce(lo.y, cub::Min()); __syncthreads(); if (threadIdx.x == 0) atomicMin(A.counters + 9, f2ord(r));
    r = BR(tmp).Reduce(lo.z, cub::Min()); __syncthreads(); if (threadIdx.x == 0) atomicMin(A.counters + 10, f2ord(r));
    r = BR(tmp).Reduce(hi.x, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(A.counters + 11, f2ord(r));
    r = BR(tmp).Reduce(hi.y, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(A.counters + 12, f2ord(r));
    r = BR(tmp).Reduce(hi.z, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(A.counters + 13, f2ord(r));
    unsigned int v = __syncthreads_count(valid);
    if (threadIdx.x == 0 && v) atomicAdd(A.counters + 0, v);
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(kB) morton_kernel(uint32_t n, BuildArrays A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = A.lo[i], hi = A.hi[i];
    unsigned long long key = ~0ull;  // invalid triangles go to the end
    if (lo.w != 0.0f) {
        float3 smin = make_float3(ord2f(A.counters[8]), ord2f(A.counters[9]), ord2f(A.counters[10]));
        float3 smax = make_float3(ord2f(A.counters[11]), ord2f(A.counters[12]), ord2f(A.counters[13]));
        float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
        float ex = fmaxf(smax.x - smin.x, 1e-30f), ey = fmaxf(smax.y - smin.y, 1e-30f), ez = fmaxf(smax.z - smin.z, 1e-30f);
        unsigned long long qx = (unsigned long long)fminf(fmaxf((cx - smin.x) / ex * 2097152.0f, 0.0f), 2097151.0f);
        unsigned long long qy = (unsigned long long)fminf(fmaxf((cy - smin.y) / ey * 2097152.0f, 0.0f), 2097151.0f);
        unsigned long long qz = (unsigned long long)fminf(fmaxf((cz - smin.z) / ez * 2097152.0f, 0.0f), 2097151.0f);
        key = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    }
    A.key_a[i] = key;
    A.idx_a[i] = i;
}

// Karras 2012. delta(i, j): common-prefix length of the keys at sorted positions i and j, positions breaking ties.
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(kB) karras_kernel(int n, const unsigned long long* __restrict__ keys, BuildArrays A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? (n - 1) + gamma : gamma;
    int right = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    A.left[i] = left;
    A.right[i] = right;
    A.first[i] = lo;
    A.last[i] = hi;
    A.parent[left] = i;
    A.parent[right] = i;
    if (i == 0) A.parent[0] = -1;
}

__global__ void __launch_bounds__(kB) refit_kernel(int n, const uint32_t* __restrict__ sorted, BuildArrays A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t t = sorted[j];
    int node = (n - 1) + j;
    float4 lo = A.lo[t], hi = A.hi[t];
    A.nlo[node] = lo;
    A.nhi[node] = hi;
    if (n == 1) return;
    int p = A.parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(A.visits + p, 1u) == 0u) return;  // the sibling subtree is not finished: its thread continues
        int l = A.left[p], r = A.right[p];
        // the sibling's box was written by another thread (published by its fence + atomic): read past L1
        float4 llo = __ldcg(A.nlo + l), lhi = __ldcg(A.nhi + l), rlo = __ldcg(A.nlo + r), rhi = __ldcg(A.nhi + r);
        A.nlo[p] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
        A.nhi[p] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
        p = A.parent[p];
    }
}

__device__ __forceinline__ int node_count(const BuildArrays& A, int n, int node) { return node >= n - 1 ? 1 : A.last[node] - A.first[node] + 1; }
__device__ __forceinline__ bool expandable(const BuildArrays& A, int n, int node) { return node < n - 1 && node_count(A, n, node) > (int)kLeafMax; }
__device__ __forceinline__ float node_area(const BuildArrays& A, int node) {  // bvh.rs:811-814
    float4 lo = A.nlo[node], hi = A.nhi[node];
    float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return 2.0f * (ex * ey + ey * ez + ez * ex);
}
// The up-to-four lanes of the BVH4 node rooted at binary node b (bvh.rs:1328-1375).
__device__ __forceinline__ int lanes_of(const BuildArrays& A, int n, int b, int kids[4]) {
    kids[0] = A.left[b];
    kids[1] = A.right[b];
    int nk = 2;
    while (nk < 4) {
        int best = -1;
        float best_a = 0.0f;
        for (int i = 0; i < nk; ++i)
            if (expandable(A, n, kids[i])) {
                float a = node_area(A, kids[i]);
                if (best < 0 || a > best_a) best = i, best_a = a;
            }
        if (best < 0) break;
        int k = kids[best];
        kids[best] = A.left[k];
        kids[nk++] = A.right[k];
    }
    return nk;
}

__global__ void __launch_bounds__(kB) mark_kernel(int n, const int* __restrict__ frontier, uint32_t count, int* __restrict__ next, BuildArrays A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int b = frontier[i];
    A.is_root4[b] = 1u;
    int kids[4];
    int nk = lanes_of(A, n, b, kids);
    for (int l = 0; l < nk; ++l)
        if (expandable(A, n, kids[l])) next[atomicAdd(A.counters + 1, 1u)] = kids[l];
}

__device__ __forceinline__ void write_leaf_records(const BuildArrays& A, const uint32_t* __restrict__ sorted, float4* __restrict__ leafs, int first, int last) {
    for (int j = first; j <= last; ++j) {
        uint32_t t = sorted[j];
        float4 a = A.v0[t], b = A.v1[t], c = A.v2[t];
        int ti = j - first;
        uint32_t flags = REC_TRI | ((uint32_t)(ti & 3) << REC_LANE_SHIFT);
        if ((ti & 3) == 3 || j == last) flags |= REC_PACKET_END;
        if (j == last) flags |= REC_LAST;
        // prim.rs:80-83: geometric normal exactly zero (glam cross: a.yzx * b.zxy - a.zxy * b.yzx, no FMA)
        float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z, e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
        float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e2y, e1z)), ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e2z, e1x)),
              nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e2x, e1y));
        if (nx == 0.0f && ny == 0.0f && nz == 0.0f) flags |= REC_DEGENERATE;
        float4* out = leafs + 3 * (size_t)j;
        out[0] = make_float4(a.x, a.y, a.z, __uint_as_float(flags));
        out[1] = make_float4(b.x, b.y, b.z, __uint_as_float(t));
        out[2] = make_float4(c.x, c.y, c.z, a.w);
    }
}

__global__ void __launch_bounds__(kB) emit_kernel(int n, const uint32_t* __restrict__ sorted, DevNode* __restrict__ nodes, float4* __restrict__ leafs, BuildArrays A) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n - 1 || !A.is_root4[b]) return;
    int kids[4];
    int nk = lanes_of(A, n, b, kids);
    DevNode w;
    for (int l = 0; l < 4; ++l) {
        w.bmin_x[l] = w.bmin_y[l] = w.bmin_z[l] = w.bmax_x[l] = w.bmax_y[l] = w.bmax_z[l] = INFINITY;
        w.child[l] = CRT_EMPTY_LANE;
    }
    w.flags = 0;
    w.pad[0] = w.pad[1] = w.pad[2] = 0;
    unsigned int n_leaves = 0;
    for (int l = 0; l < nk; ++l) {
        int k = kids[l];
        float4 lo = A.nlo[k], hi = A.nhi[k];
        w.bmin_x[l] = lo.x, w.bmin_y[l] = lo.y, w.bmin_z[l] = lo.z;
        w.bmax_x[l] = hi.x, w.bmax_y[l] = hi.y, w.bmax_z[l] = hi.z;
        w.flags |= 1u << l;
        if (expandable(A, n, k)) {
            w.child[l] = A.id4[k];
        } else {
            int first = k >= n - 1 ? k - (n - 1) : A.first[k], last = k >= n - 1 ? first : A.last[k];
            write_leaf_records(A, sorted, leafs, first, last);
            w.child[l] = CRT_LEAF_BIT | (uint32_t)(3 * first);
            w.flags |= 1u << (4 + l);
            ++n_leaves;
        }
    }
    nodes[A.id4[b]] = w;
    if (n_leaves) atomicAdd(A.counters + 2, n_leaves);
}

// n <= kLeafMax triangles: one node with one leaf lane (host_build.cpp run(): bin[0].count > 0)
__global__ void tiny_kernel(int n, const uint32_t* __restrict__ sorted, DevNode* __restrict__ nodes, float4* __restrict__ leafs, BuildArrays A) {
    DevNode w;
    for (int l = 0; l < 4; ++l) {
        w.bmin_x[l] = w.bmin_y[l] = w.bmin_z[l] = w.bmax_x[l] = w.bmax_y[l] = w.bmax_z[l] = INFINITY;
        w.child[l] = CRT_EMPTY_LANE;
    }
    w.pad[0] = w.pad[1] = w.pad[2] = 0;
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    for (int j = 0; j < n; ++j) {
        float4 a = A.lo[sorted[j]], b = A.hi[sorted[j]];
        lo = make_float3(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z));
        hi = make_float3(fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z));
    }
    w.bmin_x[0] = lo.x, w.bmin_y[0] = lo.y, w.bmin_z[0] = lo.z, w.bmax_x[0] = hi.x, w.bmax_y[0] = hi.y, w.bmax_z[0] = hi.z;
    w.child[0] = CRT_LEAF_BIT | 0u;
    w.flags = 1u | (1u << 4);
    write_leaf_records(A, sorted, leafs, 0, n - 1);
    nodes[0] = w;
    A.counters[2] = 1;
}

__global__ void pad_kernel(float4* leafs, size_t first) {  // kernels may over-read one triangle record past the end
    for (int i = 0; i < 4; ++i) leafs[first + i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(REC_LAST | REC_SPHERE));
}

template <typename T>
cudaError_t dmalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

}  // namespace

bool device_build_supported(const std::vector<Geom>& geoms) {
    for (const Geom& g : geoms)
        if (g.type != G_MESH || g.has_normals) return false;
    return true;
}

int device_build(const std::vector<Geom>& geoms, crt_scene* s) {
    auto t0 = std::chrono::steady_clock::now();
    // ---- concatenate the meshes (host) -------------------------------------------------------------------
    std::vector<uint32_t> tri_offset, vert_offset, n_verts, geom_id, mask;
    size_t n_tris = 0, n_v = 0;
    for (size_t g = 0; g < geoms.size(); ++g) {
        const Geom& G = geoms[g];
        if (G.indices.empty()) continue;
        tri_offset.push_back((uint32_t)n_tris);
        vert_offset.push_back((uint32_t)n_v);
        n_verts.push_back((uint32_t)(G.verts.size() / 3));
        geom_id.push_back((uint32_t)g);
        mask.push_back(G.mask);
        n_tris += G.indices.size() / 3;
        n_v += G.verts.size() / 3;
    }
    tri_offset.push_back((uint32_t)n_tris);
    if (n_tris > 0x0FFFFFFFull || n_v > 0xFFFFFFFFull) return set_error(CRT_ERR_INVALID_ARG, "device build: more than 2^28 triangles");
    const uint32_t n_meshes = (uint32_t)geom_id.size();
    s->n_prims = 0;
    s->non_empty = false;
    s->features = 0;
    s->desc = DevScene{};
    if (n_tris == 0) {  // nothing to intersect: the kernels treat n_nodes == 0 as an empty scene
        DevScene d{};
        CKB(cudaMalloc((void**)&s->d_desc, sizeof(DevScene)));
        CKB(cudaMemcpy(s->d_desc, &d, sizeof d, cudaMemcpyHostToDevice));
        return CRT_OK;
    }
    float* d_verts = nullptr;
    uint32_t *d_idx = nullptr, *d_tab = nullptr;
    CKB(dmalloc(&d_verts, 3 * n_v));
    CKB(dmalloc(&d_idx, 3 * n_tris));
    CKB(dmalloc(&d_tab, (size_t)5 * n_meshes + 1));
    {
        size_t vo = 0, io = 0;
        for (const Geom& G : geoms) {
            if (G.indices.empty()) continue;
            CKB(cudaMemcpyAsync(d_verts + vo, G.verts.data(), G.verts.size() * 4, cudaMemcpyHostToDevice, 0));
            CKB(cudaMemcpyAsync(d_idx + io, G.indices.data(), G.indices.size() * 4, cudaMemcpyHostToDevice, 0));
            vo += G.verts.size();
            io += G.indices.size();
        }
        std::vector<uint32_t> tab;
        tab.insert(tab.end(), tri_offset.begin(), tri_offset.end());
        tab.insert(tab.end(), vert_offset.begin(), vert_offset.end());
        tab.insert(tab.end(), n_verts.begin(), n_verts.end());
        tab.insert(tab.end(), geom_id.begin(), geom_id.end());
        tab.insert(tab.end(), mask.begin(), mask.end());
        CKB(cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    }
    MeshTable mt{d_tab, d_tab + n_meshes + 1, d_tab + 2 * (size_t)n_meshes + 1, d_tab + 3 * (size_t)n_meshes + 1, d_tab + 4 * (size_t)n_meshes + 1, n_meshes};
    auto t_up = std::chrono::steady_clock::now();

    // ---- device arrays --------------------------------------------------------------------------------------
    const uint32_t N = (uint32_t)n_tris;
    BuildArrays A{};
    CKB(dmalloc(&A.v0, N)); CKB(dmalloc(&A.v1, N)); CKB(dmalloc(&A.v2, N));
    CKB(dmalloc(&A.lo, N)); CKB(dmalloc(&A.hi, N));
    uint2* d_prim_ids = nullptr;
    CKB(dmalloc(&d_prim_ids, N));
    A.prim_ids = d_prim_ids;
    CKB(dmalloc(&A.key_a, N)); CKB(dmalloc(&A.key_b, N)); CKB(dmalloc(&A.idx_a, N)); CKB(dmalloc(&A.idx_b, N));
    CKB(dmalloc(&A.counters, 16));
    {
        unsigned int init[16] = {0};
        for (int k = 8; k < 11; ++k) init[k] = 0xFFFFFFFFu;
        CKB(cudaMemcpy(A.counters, init, sizeof init, cudaMemcpyHostToDevice));
    }
    const int grid_t = (int)((N + kB - 1) / kB);
    tri_setup_kernel<<<grid_t, kB>>>(N, mt, d_verts, d_idx, A);
    morton_kernel<<<grid_t, kB>>>(N, A);
    cub::DoubleBuffer<unsigned long long> keys(A.key_a, A.key_b);
    cub::DoubleBuffer<uint32_t> vals(A.idx_a, A.idx_b);
    size_t tmp_bytes = 0;
    CKB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, vals, (int)N, 0, 64));
    void* d_tmp = nullptr;
    CKB(cudaMalloc(&d_tmp, tmp_bytes));
    CKB(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys, vals, (int)N, 0, 64));
    unsigned int h_counters[16];
    CKB(cudaMemcpy(h_counters, A.counters, sizeof h_counters, cudaMemcpyDeviceToHost));
    const int n = (int)h_counters[0];  // valid triangles: the first n sorted positions
    const unsigned long long* d_keys = keys.Current();
    const uint32_t* d_sorted = vals.Current();
    crt::count_launches(4);

    s->n_prims = (uint64_t)n;
    s->breakdown.triangles = (uint64_t)n;
    s->non_empty = n > 0;
    if (n > 0) s->root = Box{{ord2f(h_counters[8]), ord2f(h_counters[9]), ord2f(h_counters[10])}, {ord2f(h_counters[11]), ord2f(h_counters[12]), ord2f(h_counters[13])}};

    DevNode* d_nodes = nullptr;
    float4* d_leafs = nullptr;
    uint32_t n4 = 0;
    const size_t leaf_quads = 3 * (size_t)std::max(n, 0) + 4;
    if (n > 0) {
        CKB(dmalloc(&d_leafs, leaf_quads));
        pad_kernel<<<1, 1>>>(d_leafs, 3 * (size_t)n);
        if (n <= (int)kLeafMax) {
            CKB(dmalloc(&d_nodes, 1));
            tiny_kernel<<<1, 1>>>(n, d_sorted, d_nodes, d_leafs, A);
            n4 = 1;
            crt::count_launches(2);
        } else {
            const int n_bin = 2 * n - 1;
            CKB(dmalloc(&A.left, n)); CKB(dmalloc(&A.right, n)); CKB(dmalloc(&A.first, n)); CKB(dmalloc(&A.last, n));
            CKB(dmalloc(&A.parent, n_bin)); CKB(dmalloc(&A.nlo, n_bin)); CKB(dmalloc(&A.nhi, n_bin));
            CKB(dmalloc(&A.visits, n)); CKB(dmalloc(&A.is_root4, n)); CKB(dmalloc(&A.id4, n));
            CKB(dmalloc(&A.frontier_a, n)); CKB(dmalloc(&A.frontier_b, n));
            CKB(cudaMemset(A.visits, 0, (size_t)n * 4));
            CKB(cudaMemset(A.is_root4, 0, (size_t)n * 4));
            const int grid_n = (n + kB - 1) / kB;
            karras_kernel<<<grid_n, kB>>>(n, d_keys, A);
            refit_kernel<<<grid_n, kB>>>(n, d_sorted, A);
            // level-synchronous marking of the binary nodes that become BVH4 nodes
            int zero = 0;
            CKB(cudaMemcpy(A.frontier_a, &zero, 4, cudaMemcpyHostToDevice));
            uint32_t count = 1;
            int *cur = A.frontier_a, *nxt = A.frontier_b;
            uint64_t launches = 3;
            while (count > 0) {
                CKB(cudaMemset(A.counters + 1, 0, 4));
                mark_kernel<<<(int)((count + kB - 1) / kB), kB>>>(n, cur, count, nxt, A);
                CKB(cudaMemcpy(&count, A.counters + 1, 4, cudaMemcpyDeviceToHost));
                std::swap(cur, nxt);
                ++launches;
            }
            size_t scan_bytes = 0;
            CKB(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, A.is_root4, A.id4, n - 1));
            void* d_scan = nullptr;
            CKB(cudaMalloc(&d_scan, scan_bytes));
            CKB(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, A.is_root4, A.id4, n - 1));
            uint32_t last_flag = 0, last_id = 0;
            CKB(cudaMemcpy(&last_flag, A.is_root4 + (n - 2), 4, cudaMemcpyDeviceToHost));
            CKB(cudaMemcpy(&last_id, A.id4 + (n - 2), 4, cudaMemcpyDeviceToHost));
            n4 = last_id + last_flag;
            CKB(dmalloc(&d_nodes, n4));
            emit_kernel<<<grid_n, kB>>>(n, d_sorted, d_nodes, d_leafs, A);
            crt::count_launches(launches + 2);
            cudaFree(d_scan);
        }
        CKB(cudaDeviceSynchronize());
        CKB(cudaMemcpy(h_counters, A.counters, sizeof h_counters, cudaMemcpyDeviceToHost));
    }
    // ---- hand the arrays to the scene, drop the scratch ---------------------------------------------------
    s->d_nodes = d_nodes;
    s->d_leafs = d_leafs;
    s->d_prim_ids = d_prim_ids;
    s->desc.nodes = d_nodes;
    s->desc.leafs = (const U4*)d_leafs;
    s->desc.prim_ids = (const U2*)d_prim_ids;
    s->desc.normals = nullptr;
    s->desc.motion = nullptr;
    s->desc.n_nodes = n4;
    s->desc.has_triangles = n > 0;
    CKB(cudaMalloc((void**)&s->d_desc, sizeof(DevScene)));
    CKB(cudaMemcpy(s->d_desc, &s->desc, sizeof(DevScene), cudaMemcpyHostToDevice));
    for (void* p : {(void*)d_verts, (void*)d_idx, (void*)d_tab, (void*)A.v0, (void*)A.v1, (void*)A.v2, (void*)A.lo, (void*)A.hi, (void*)A.key_a, (void*)A.key_b,
                    (void*)A.idx_a, (void*)A.idx_b, (void*)A.counters, d_tmp, (void*)A.left, (void*)A.right, (void*)A.first, (void*)A.last, (void*)A.parent,
                    (void*)A.nlo, (void*)A.nhi, (void*)A.visits, (void*)A.is_root4, (void*)A.id4, (void*)A.frontier_a, (void*)A.frontier_b})
        cudaFree(p);
    auto t1 = std::chrono::steady_clock::now();
    s->info.primitives = (uint64_t)n;
    s->info.binary_nodes = n > 0 ? 2 * (uint64_t)n - 1 : 0;
    s->info.wide_nodes = n4;
    s->info.leaves = h_counters[2];
    s->info.leaf_refs = (uint64_t)std::max(n, 0);
    s->info.node_bytes = (uint64_t)n4 * sizeof(DevNode);
    s->info.leaf_bytes = n > 0 ? leaf_quads * 16 : 0;
    s->info.upload_seconds = std::chrono::duration<double>(t_up - t0).count();
    s->info.build_seconds = std::chrono::duration<double>(t1 - t_up).count();
    s->footprint.bvh_nodes = s->info.node_bytes;
    s->footprint.leaf_records = s->info.leaf_bytes;
    s->footprint.prim_table = (uint64_t)N * 8;
    s->footprint.descriptors = sizeof(DevScene);
    return CRT_OK;
}

}  // namespace crt
