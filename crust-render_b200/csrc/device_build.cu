// Device-side BVH4 build for sm_100a: an alternative commit() path that constructs the whole
// acceleration structure on the GPU — over triangle meshes (expanded on the device) and over every other primitive kind
// (spheres, curve segments, cubic spans, instances incl. moving ones: boxes and leaf records prepared on the host exactly as
// the host build prepares them, host_build.cpp expand_other_prims, then sorted and grouped on the device with the triangles).
//
// The reference builds its tree on the host (crates/crust-rt/src/bvh.rs:300-327, 1059-1375: binned SAH +
// SBVH, then a collapse to BVH4); csrc/host_build.cpp restates that build exactly and stays the default,
// because hit ties ("later wins", bvh.rs:537-544) depend on the tree. This file is the fast path for when
// build time matters more than reproducing the reference's tree: seconds on the host become milliseconds.
// It emits the same device layout (crt_layout.h), so the traversal kernels are unchanged.
//
//   tri_setup      gather each triangle from its mesh (scene.rs:240-264: out-of-range indices skip the
//                  triangle but still consume a prim_id), padded AABB (aabb.rs:47-58), scene bounds
//   extent/count   early split clipping: triangles much larger than the mean get a regular grid over their box
//   split          one candidate reference per grid cell: the triangle clipped to the cell (Sutherland-Hodgman),
//                  its box, and the 63-bit Morton code of the box centre; empty cells sort to the end
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

#include <algorithm>
#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>
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
#ifndef CRT_DEV_LEAF_MAX
#define CRT_DEV_LEAF_MAX 4
#endif
constexpr uint32_t kLeafMax = CRT_DEV_LEAF_MAX;  // bvh.rs:142-167 MIN_LEAF_PACKED: one Tri4 packet per leaf

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
    float4 *lo, *hi;             // padded AABB of the triangle; lo.w = valid
    uint32_t *cells, *cell_off;  // pre-split grid of the triangle (nx | ny << 8 | nz << 16) and its first candidate reference
    uint2* prim_ids;             // (geom_id, prim_id)
    // per candidate reference (a triangle, or the part of a large triangle inside one cell of its split grid)
    uint32_t* ref_prim;
    float4 *rlo, *rhi;
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
    unsigned int* counters;      // [0] valid triangles, [1] next frontier size, [2] leaves, [3] valid references,
                                 // [4..5] sum of quantised triangle extents (u64), [8..13] bounds as ordered ints,
                                 // [16..17] sum of the triangle box diagonals (double), [18] largest diagonal (float bits)
    // primitives that are not triangles: indices [n_tris, n_tris + n_other) of the per-primitive arrays
    uint32_t n_tris;
    const uint32_t* other_rec_off;  // n_other + 1 offsets (16-byte units) into other_recs
    const float4* other_recs;       // their leaf records as the host flattener writes them (without REC_LAST)
    uint32_t* other_pos;            // out: where each one's record landed in the leaf stream
    uint32_t* rec_size;             // per sorted reference: size of its record in 16-byte units ...
    uint32_t* rec_at;               // ... and its exclusive prefix sum = where a leaf that starts with it begins
};
constexpr int kCounters = 24;

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

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
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
    // scene.rs:446-455 primitive extents (the host sums f32 diagonals in input order; here a double sum of the same terms)
    float diag = 0.0f;
    if (valid) {
        float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
        diag = sqrtf((ex * ex + ey * ey) + ez * ez);
    }
    typedef cub::BlockReduce<double, kB> BRD;
    __shared__ typename BRD::TempStorage tmpd;
    __syncthreads();
    double dsum = BRD(tmpd).Sum((double)diag);
    if (threadIdx.x == 0 && dsum != 0.0) atomicAdd((double*)(A.counters + 16), dsum);
    r = BR(tmp).Reduce(diag, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(A.counters + 18, __float_as_uint(r));
    unsigned int v = __syncthreads_count(valid);
    if (threadIdx.x == 0 && v) atomicAdd(A.counters + 0, v);
}

// Boxes and ids of the primitives that are not triangles, prepared on the host, into their slots behind the triangles.
__global__ void __launch_bounds__(kB) other_setup_kernel(uint32_t n_other, const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                         const uint2* __restrict__ ids, BuildArrays A) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_other) return;
    const uint32_t i = A.n_tris + j;
    A.lo[i] = lo[j];
    A.hi[i] = hi[j];
    A.prim_ids[i] = ids[j];
}

// Mean triangle size, deterministically: longest AABB side of every valid triangle, quantised against the scene
// extent (24 bits) and summed in integers.
__global__ void __launch_bounds__(kB) extent_kernel(uint32_t n, float scene_extent, BuildArrays A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long q = 0;
    if (i < n) {
        float4 lo = A.lo[i], hi = A.hi[i];
        if (lo.w != 0.0f) q = (unsigned long long)(fminf(fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z) / scene_extent, 1.0f) * 16777216.0f);
    }
    typedef cub::BlockReduce<unsigned long long, kB> BR;
    __shared__ typename BR::TempStorage tmp;
    unsigned long long r = BR(tmp).Sum(q);
    if (threadIdx.x == 0 && r) atomicAdd((unsigned long long*)(A.counters + 4), r);
}

// How many candidate references a triangle gets: triangles much larger than the typical one are cut along a
// regular grid over their box (early split clipping), so one long shard or a ground plane does not inflate every
// node above it. The reference reaches the same end with SBVH spatial splits (bvh.rs:974-1054).
__global__ void __launch_bounds__(kB) count_kernel(uint32_t n, float cell, float split_above, BuildArrays A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = A.lo[i], hi = A.hi[i];
    uint32_t dims = 0, count = 0;
    if (lo.w != 0.0f) {
        float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
        uint32_t nx = 1, ny = 1, nz = 1;
        if (i < A.n_tris && fmaxf(fmaxf(ex, ey), ez) > split_above) {  // only triangles are clipped to cells
            float c = cell;
            for (int it = 0; it < 16; ++it) {  // coarsen until the grid has at most 4096 cells and 64 per axis
                nx = (uint32_t)fminf(fmaxf(ceilf(ex / c), 1.0f), 64.0f);
                ny = (uint32_t)fminf(fmaxf(ceilf(ey / c), 1.0f), 64.0f);
                nz = (uint32_t)fminf(fmaxf(ceilf(ez / c), 1.0f), 64.0f);
                if (nx * ny * nz <= 4096u) break;
                c *= 1.5f;
            }
            if (nx * ny * nz > 4096u) nx = ny = nz = 1;
        }
        dims = nx | (ny << 8) | (nz << 16);
        count = nx * ny * nz;
    }
    A.cells[i] = dims;
    A.cell_off[i] = count;  // scanned in place afterwards
}

// Sutherland-Hodgman against one axis-aligned plane; keeps the side where sign * (p[axis] - value) >= 0.
__device__ int clip_plane(float3* in, int n, float3* out, int axis, float value, float sign) {
    int m = 0;
    for (int k = 0; k < n; ++k) {
        float3 a = in[k], b = in[(k + 1) % n];
        float da = sign * ((axis == 0 ? a.x : axis == 1 ? a.y : a.z) - value), db = sign * ((axis == 0 ? b.x : axis == 1 ? b.y : b.z) - value);
        if (da >= 0.0f) out[m++] = a;
        if ((da > 0.0f && db < 0.0f) || (da < 0.0f && db > 0.0f)) {
            float t = da / (da - db);
            float3 p = make_float3(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z));
            if (axis == 0) p.x = value; else if (axis == 1) p.y = value; else p.z = value;  // snap onto the plane
            out[m++] = p;
        }
    }
    return m;
}

// One candidate reference (triangle x grid cell): clipped box + Morton key. Returns whether the cell holds any of the triangle.
__device__ bool split_candidate(uint32_t r, uint32_t n_tris, float eps, const BuildArrays& A) {
    // triangle of this candidate: last t with cell_off[t] <= r (cell_off is the exclusive scan of the counts)
    uint32_t a = 0, b = n_tris;
    while (b - a > 1) {
        uint32_t m = (a + b) >> 1;
        if (__ldg(A.cell_off + m) <= r) a = m; else b = m;
    }
    const uint32_t t = a, dims = A.cells[t], local = r - A.cell_off[t];
    const uint32_t nx = dims & 0xFFu, ny = (dims >> 8) & 0xFFu, nz = dims >> 16;
    float4 tlo = A.lo[t], thi = A.hi[t];
    float3 lo = make_float3(tlo.x, tlo.y, tlo.z), hi = make_float3(thi.x, thi.y, thi.z);
    bool valid = true;
    if (nx * ny * nz > 1u) {
        uint32_t ix = local % nx, iy = (local / nx) % ny, iz = local / (nx * ny);
        float sx = (thi.x - tlo.x) / (float)nx, sy = (thi.y - tlo.y) / (float)ny, sz = (thi.z - tlo.z) / (float)nz;
        // cell faces are computed the same way from both sides, so neighbouring cells share them exactly
        float3 clo = make_float3(ix == 0 ? tlo.x : tlo.x + (float)ix * sx, iy == 0 ? tlo.y : tlo.y + (float)iy * sy, iz == 0 ? tlo.z : tlo.z + (float)iz * sz);
        float3 chi = make_float3(ix + 1 == nx ? thi.x : tlo.x + (float)(ix + 1) * sx, iy + 1 == ny ? thi.y : tlo.y + (float)(iy + 1) * sy,
                                 iz + 1 == nz ? thi.z : tlo.z + (float)(iz + 1) * sz);
        float3 pa[10], pb[10];
        float4 v0 = A.v0[t], v1 = A.v1[t], v2 = A.v2[t];
        pa[0] = make_float3(v0.x, v0.y, v0.z), pa[1] = make_float3(v1.x, v1.y, v1.z), pa[2] = make_float3(v2.x, v2.y, v2.z);
        int m = 3;
        m = clip_plane(pa, m, pb, 0, clo.x, 1.0f);
        if (m) m = clip_plane(pb, m, pa, 0, chi.x, -1.0f);
        if (m) m = clip_plane(pa, m, pb, 1, clo.y, 1.0f);
        if (m) m = clip_plane(pb, m, pa, 1, chi.y, -1.0f);
        if (m) m = clip_plane(pa, m, pb, 2, clo.z, 1.0f);
        if (m) m = clip_plane(pb, m, pa, 2, chi.z, -1.0f);
        valid = m > 0;
        if (valid) {
            lo = pa[0], hi = pa[0];
            for (int k = 1; k < m; ++k) {
                lo = make_float3(fminf(lo.x, pa[k].x), fminf(lo.y, pa[k].y), fminf(lo.z, pa[k].z));
                hi = make_float3(fmaxf(hi.x, pa[k].x), fmaxf(hi.y, pa[k].y), fmaxf(hi.z, pa[k].z));
            }
            // conservative: interpolated corners carry rounding error; never leave the triangle's own box
            lo = make_float3(fmaxf(lo.x - eps, tlo.x), fmaxf(lo.y - eps, tlo.y), fmaxf(lo.z - eps, tlo.z));
            hi = make_float3(fminf(hi.x + eps, thi.x), fminf(hi.y + eps, thi.y), fminf(hi.z + eps, thi.z));
            // ... and, like every box the reference makes (aabb.rs:47-58; clip_triangle_aabb, triangle.rs:367-429), never
            // thinner than the pad: the slab test of a ray whose origin is hundreds of units away rounds by more than the
            // thickness of an unpadded piece of an axis-aligned triangle, and would step over it
            if (hi.x - lo.x < kPadDev) lo.x -= kPadDev, hi.x += kPadDev;
            if (hi.y - lo.y < kPadDev) lo.y -= kPadDev, hi.y += kPadDev;
            if (hi.z - lo.z < kPadDev) lo.z -= kPadDev, hi.z += kPadDev;
        }
    }
    unsigned long long key = ~0ull;
    if (valid) {
        float3 smin = make_float3(ord2f(A.counters[8]), ord2f(A.counters[9]), ord2f(A.counters[10]));
        float3 smax = make_float3(ord2f(A.counters[11]), ord2f(A.counters[12]), ord2f(A.counters[13]));
        // one scale for the three axes: cubic Morton cells (per-axis normalisation would cut flat scenes into pancakes)
        float e = fmaxf(fmaxf(smax.x - smin.x, smax.y - smin.y), fmaxf(smax.z - smin.z, 1e-30f));
        float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
        unsigned long long qx = (unsigned long long)fminf(fmaxf((cx - smin.x) / e * 2097152.0f, 0.0f), 2097151.0f);
        unsigned long long qy = (unsigned long long)fminf(fmaxf((cy - smin.y) / e * 2097152.0f, 0.0f), 2097151.0f);
        unsigned long long qz = (unsigned long long)fminf(fmaxf((cz - smin.z) / e * 2097152.0f, 0.0f), 2097151.0f);
        key = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    }
    A.ref_prim[r] = t;
    A.rlo[r] = make_float4(lo.x, lo.y, lo.z, 0.0f);
    A.rhi[r] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    A.key_a[r] = key;
    A.idx_a[r] = r;
    return valid;
}

__global__ void __launch_bounds__(kB) split_kernel(uint32_t n_tris, uint32_t n_cand, float eps, BuildArrays A) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    // no early return: every thread of the (possibly partial) last block must reach the block-wide count
    const bool valid = r < n_cand && split_candidate(r, n_tris, eps, A);
    unsigned int v = __syncthreads_count(valid);
    if (threadIdx.x == 0 && v) atomicAdd(A.counters + 3, v);
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
    float4 lo = A.rlo[t], hi = A.rhi[t];
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

// ---- SAH-built top (optional, CRT_DEV_SAH_TOP) -------------------------------------------------------------------------
// The radix tree's upper levels split where the Morton code changes, wherever the geometry is. The binary nodes holding
// more than `cluster` references are therefore rebuilt: their children that hold at most `cluster` references ("clusters",
// a few thousand contiguous runs of the sorted references) go to the host with their boxes, the host builds a binary tree
// over them with a full-sweep surface-area heuristic, and that tree is written into the slots of the nodes it replaces
// (a binary tree over K clusters has K - 1 internal nodes — exactly the nodes above them). Below the clusters nothing
// changes, so leaves still cover contiguous runs; above them only left / right / boxes are read by the later stages.
struct TopCluster {
    float lo[3], hi[3];
    int node, count, first;
};
__global__ void __launch_bounds__(kB) top_select_kernel(int n, int cluster, unsigned int cap, BuildArrays A, int* __restrict__ top_ids,
                                                        TopCluster* __restrict__ clusters, unsigned int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || A.last[i] - A.first[i] + 1 <= cluster) return;
    const unsigned int at = atomicAdd(counts + 0, 1u);
    if (at < cap) top_ids[at] = i;  // a lopsided radix tree can have far more such nodes than n / cluster: the caller then keeps it as it is
    const int kids[2] = {A.left[i], A.right[i]};
    for (int k = 0; k < 2; ++k) {
        const int c = kids[k];
        const int first = c >= n - 1 ? c - (n - 1) : A.first[c];
        const int cnt = c >= n - 1 ? 1 : A.last[c] - A.first[c] + 1;
        if (cnt > cluster) continue;
        TopCluster t;
        const float4 lo = A.nlo[c], hi = A.nhi[c];
        t.lo[0] = lo.x, t.lo[1] = lo.y, t.lo[2] = lo.z, t.hi[0] = hi.x, t.hi[1] = hi.y, t.hi[2] = hi.z;
        t.node = c, t.count = cnt, t.first = first;
        const unsigned int ct = atomicAdd(counts + 1, 1u);
        if (ct < cap) clusters[ct] = t;
    }
}
struct TopNode {
    float lo[3], hi[3];
    int slot, left, right;
};
__global__ void __launch_bounds__(kB) top_apply_kernel(int n, int count, const TopNode* __restrict__ t, BuildArrays A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const TopNode v = t[i];
    A.left[v.slot] = v.left, A.right[v.slot] = v.right;
    A.first[v.slot] = 0, A.last[v.slot] = n - 1;  // only "more than a leaf's worth" is read from these above the clusters
    A.nlo[v.slot] = make_float4(v.lo[0], v.lo[1], v.lo[2], 0.0f);
    A.nhi[v.slot] = make_float4(v.hi[0], v.hi[1], v.hi[2], 0.0f);
    A.parent[v.left] = v.slot, A.parent[v.right] = v.slot;
}

// Host side: full-sweep SAH over the clusters (cost = area x references on each side, the reference's measure, bvh.rs:811-830).
struct TopBuilder {
    const std::vector<TopCluster>& c;
    const std::vector<int>& slots;  // ascending; slots[0] == 0 is the root
    std::vector<TopNode> out;
    size_t next_slot = 0;
    static float area(const float lo[3], const float hi[3]) {
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        return 2.0f * (ex * ey + ey * ez + ez * ex);
    }
    int build(std::vector<int>& idx, size_t b, size_t e) {  // returns the binary node id of the subtree over idx[b, e)
        if (e - b == 1) return c[idx[b]].node;
        const size_t m = e - b;
        int best_axis = -1;
        size_t best_pos = 0;
        double best_cost = 0.0;
        {
        std::vector<float> ra(m);
        std::vector<long long> rn(m);
        for (int axis = 0; axis < 3; ++axis) {
            std::sort(idx.begin() + b, idx.begin() + e, [&](int x, int y) {
                const float cx = c[x].lo[axis] + c[x].hi[axis], cy = c[y].lo[axis] + c[y].hi[axis];
                return cx < cy || (cx == cy && c[x].first < c[y].first);
            });
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            long long cnt = 0;
            for (size_t k = m; k-- > 1;) {  // suffix [k, m)
                const TopCluster& t = c[idx[b + k]];
                for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], t.lo[a]), hi[a] = std::max(hi[a], t.hi[a]);
                cnt += t.count;
                ra[k] = area(lo, hi), rn[k] = cnt;
            }
            for (int a = 0; a < 3; ++a) lo[a] = INFINITY, hi[a] = -INFINITY;
            cnt = 0;
            for (size_t k = 1; k < m; ++k) {  // prefix [0, k)
                const TopCluster& t = c[idx[b + k - 1]];
                for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], t.lo[a]), hi[a] = std::max(hi[a], t.hi[a]);
                cnt += t.count;
                const double cost = (double)area(lo, hi) * (double)cnt + (double)ra[k] * (double)rn[k];
                if (best_axis < 0 || cost < best_cost) best_axis = axis, best_pos = k, best_cost = cost;
            }
        }
        }  // the sweep's scratch is gone before the recursion
        if (best_axis != 2)
            std::sort(idx.begin() + b, idx.begin() + e, [&](int x, int y) {
                const float cx = c[x].lo[best_axis] + c[x].hi[best_axis], cy = c[y].lo[best_axis] + c[y].hi[best_axis];
                return cx < cy || (cx == cy && c[x].first < c[y].first);
            });
        const size_t me = out.size();
        out.push_back(TopNode{});
        out[me].slot = slots[next_slot++];
        const int l = build(idx, b, b + best_pos), r = build(idx, b + best_pos, e);
        TopNode& t = out[me];
        t.left = l, t.right = r;
        for (int a = 0; a < 3; ++a) t.lo[a] = INFINITY, t.hi[a] = -INFINITY;
        for (size_t k = b; k < e; ++k)
            for (int a = 0; a < 3; ++a) t.lo[a] = std::min(t.lo[a], c[idx[k]].lo[a]), t.hi[a] = std::max(t.hi[a], c[idx[k]].hi[a]);
        return t.slot;
    }
};

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

// Record size (16-byte units) of every sorted reference, for the prefix sum that places the leaves.
__global__ void __launch_bounds__(kB) rec_size_kernel(int n, const uint32_t* __restrict__ sorted, BuildArrays A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t t = A.ref_prim[sorted[j]];
    A.rec_size[j] = t < A.n_tris ? REC_TRI_Q : A.other_rec_off[t - A.n_tris + 1] - A.other_rec_off[t - A.n_tris];
}

// The records of the leaf over sorted references [first, last], in the order the traversal expects (host_build.cpp
// emit_leaf, bvh.rs:1256-1290): its triangles first — packet flags on every 4th and on the last one — then its other
// primitives; the last record of the leaf carries REC_LAST. Returns the leaf's offset in the record stream.
__device__ __forceinline__ uint32_t write_leaf_records(const BuildArrays& A, const uint32_t* __restrict__ sorted, float4* __restrict__ leafs, int first, int last) {
    const uint32_t base = A.rec_at[first];
    int n_tri = 0, n_other = 0;
    for (int j = first; j <= last; ++j) {
        if (A.ref_prim[sorted[j]] < A.n_tris) ++n_tri; else ++n_other;
    }
    int ti = 0;
    for (int j = first; j <= last; ++j) {
        uint32_t t = A.ref_prim[sorted[j]];
        if (t >= A.n_tris) continue;
        float4 a = A.v0[t], b = A.v1[t], c = A.v2[t];
        uint32_t flags = REC_TRI | ((uint32_t)(ti & 3) << REC_LANE_SHIFT);
        if ((ti & 3) == 3 || ti + 1 == n_tri) flags |= REC_PACKET_END;
        if (ti + 1 == n_tri && n_other == 0) flags |= REC_LAST;
        // prim.rs:80-83: geometric normal exactly zero (glam cross: a.yzx * b.zxy - a.zxy * b.yzx, no FMA)
        float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z, e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
        float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e2y, e1z)), ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e2z, e1x)),
              nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e2x, e1y));
        if (nx == 0.0f && ny == 0.0f && nz == 0.0f) flags |= REC_DEGENERATE;
        float4* out = leafs + base + REC_TRI_Q * (size_t)ti;
        out[0] = make_float4(a.x, a.y, a.z, __uint_as_float(flags));
        out[1] = make_float4(b.x, b.y, b.z, __uint_as_float(t));
        out[2] = make_float4(c.x, c.y, c.z, a.w);
        ++ti;
    }
    uint32_t pos = base + REC_TRI_Q * (uint32_t)n_tri;
    int oi = 0;
    for (int j = first; j <= last; ++j) {
        uint32_t t = A.ref_prim[sorted[j]];
        if (t < A.n_tris) continue;
        const uint32_t k = t - A.n_tris, r0 = A.other_rec_off[k], r1 = A.other_rec_off[k + 1];
        for (uint32_t q = r0; q < r1; ++q) {
            float4 v = A.other_recs[q];
            if (q == r0 && ++oi == n_other) v.w = __uint_as_float(__float_as_uint(v.w) | REC_LAST);
            leafs[pos + (q - r0)] = v;
        }
        A.other_pos[k] = pos;
        pos += r1 - r0;
    }
    return base;
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
            w.child[l] = CRT_LEAF_BIT | write_leaf_records(A, sorted, leafs, first, last);
            w.flags |= 1u << (4 + l);
            ++n_leaves;
        }
    }
    nodes[A.id4[b]] = w;
    if (n_leaves) atomicAdd(A.counters + 2, n_leaves);
}

// n <= kLeafMax references: one node with one leaf lane (host_build.cpp run(): bin[0].count > 0)
__global__ void tiny_kernel(int n, const uint32_t* __restrict__ sorted, DevNode* __restrict__ nodes, float4* __restrict__ leafs, BuildArrays A) {
    DevNode w;
    for (int l = 0; l < 4; ++l) {
        w.bmin_x[l] = w.bmin_y[l] = w.bmin_z[l] = w.bmax_x[l] = w.bmax_y[l] = w.bmax_z[l] = INFINITY;
        w.child[l] = CRT_EMPTY_LANE;
    }
    w.pad[0] = w.pad[1] = w.pad[2] = 0;
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    for (int j = 0; j < n; ++j) {
        float4 a = A.rlo[sorted[j]], b = A.rhi[sorted[j]];
        lo = make_float3(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z));
        hi = make_float3(fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z));
    }
    w.bmin_x[0] = lo.x, w.bmin_y[0] = lo.y, w.bmin_z[0] = lo.z, w.bmax_x[0] = hi.x, w.bmax_y[0] = hi.y, w.bmax_z[0] = hi.z;
    w.flags = 1u | (1u << 4);
    w.child[0] = CRT_LEAF_BIT | write_leaf_records(A, sorted, leafs, 0, n - 1);
    nodes[0] = w;
    A.counters[2] = 1;
}

__global__ void pad_kernel(float4* leafs, size_t first) {  // kernels may over-read one triangle record past the end
    for (int i = 0; i < 4; ++i) leafs[first + i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(REC_LAST | REC_SPHERE));
}

template <typename T>
cudaError_t dmalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

// Build scratch lives in one grow-only device arena and one pinned staging buffer kept for the life of the
// process: cudaMalloc / cudaFree of a few dozen GB-scale buffers cost more than all the build kernels together.
struct Arena {
    std::mutex mu;
    char* dev = nullptr;
    size_t dev_cap = 0;
    char* dev2 = nullptr;  // per-reference arrays (sized after the split count is known)
    size_t dev2_cap = 0;
    char* pinned = nullptr;
    size_t pinned_cap = 0;
} g_arena;
struct Carver {
    char* p;
    char* end;
    template <typename T>
    T* take(size_t count) {
        T* r = (T*)p;
        p += (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
        return p <= end ? r : nullptr;
    }
};
size_t padded(size_t count, size_t elem) { return (std::max<size_t>(count, 1) * elem + 255) & ~(size_t)255; }

}  // namespace

bool device_build_supported(const std::vector<Geom>& geoms) {
    // every primitive kind; the one thing left to the host build is per-vertex shading normals (prim.rs:77-78: the
    // importer never supplies them, usd_import.rs:904,960,1051)
    for (const Geom& g : geoms)
        if (g.type == G_MESH && g.has_normals) return false;
    return true;
}

int device_build(std::vector<Geom>& geoms, ChildInfoFn child_info, crt_scene* s) {
    auto t0 = std::chrono::steady_clock::now();
    const bool trace = getenv("CRT_DEVICE_BUILD_TRACE") != nullptr;  // per-phase wall times on stderr (synchronises)
    const bool trace_sync = trace && atoi(getenv("CRT_DEVICE_BUILD_TRACE")) != 2;  // 2 = host-side times only, no extra synchronisation
    auto lap = [&, last = t0](const char* what) mutable {
        if (!trace) return;
        if (trace_sync) cudaDeviceSynchronize();
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[device_build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    };
    // ---- concatenate the meshes (host) -------------------------------------------------------------------
    std::vector<uint32_t> tri_offset, vert_offset, n_verts, geom_id, mask;
    size_t n_tris = 0, n_v = 0;
    for (size_t g = 0; g < geoms.size(); ++g) {
        const Geom& G = geoms[g];
        if (G.type != G_MESH || G.indices.empty()) continue;
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
    // ---- everything that is not a triangle mesh: expanded on the host like the host build does it -----------------------
    OtherPrims others;
    {
        const char* err = nullptr;
        if (!expand_other_prims(geoms, child_info, (uint32_t)n_tris, others, &err)) return set_error(CRT_ERR_NESTING, err ? err : "commit failed");
    }
    const uint32_t M = (uint32_t)others.n;
    if ((uint64_t)n_tris + M > 0x0FFFFFFFull) return set_error(CRT_ERR_INVALID_ARG, "device build: more than 2^28 primitives");
    bool translated_copy = false;
    for (size_t k = 0; k < others.inst_prim.size(); ++k) {  // child addresses into the instance records (q4, q5, REC_CHILD_TRIS)
        const crt_scene* child = others.inst_child[k];
        const size_t q = others.rec_off[others.inst_prim[k]];
        const uint64_t d = (uint64_t)(uintptr_t)child->d_desc, cn = (uint64_t)(uintptr_t)child->d_nodes, cl = (uint64_t)(uintptr_t)child->d_leafs;
        others.recs[4 * (q + 4) + 0] = (uint32_t)(d & 0xFFFFFFFFu), others.recs[4 * (q + 4) + 1] = (uint32_t)(d >> 32);
        others.recs[4 * (q + 5) + 0] = (uint32_t)(cn & 0xFFFFFFFFu), others.recs[4 * (q + 5) + 1] = (uint32_t)(cn >> 32);
        others.recs[4 * (q + 5) + 2] = (uint32_t)(cl & 0xFFFFFFFFu), others.recs[4 * (q + 5) + 3] = (uint32_t)(cl >> 32);
        if (child->desc.has_triangles) others.recs[4 * q + 3] |= REC_CHILD_TRIS;
        if (child->desc.has_triangles && rec_is_translated_copy(&others.recs[4 * q])) translated_copy = true;
    }
    s->has_motion = others.has_motion;
    s->instance_depth = others.instance_depth;
    s->breakdown.spheres = others.counts.spheres, s->breakdown.curve_segments = others.counts.curve_segments;
    s->breakdown.cubic_curve_spans = others.counts.cubic_curve_spans, s->breakdown.instances = others.counts.instances;
    uint32_t feat = 0;
    if (others.counts.spheres) feat |= FEAT_SPHERE;
    if (others.counts.curve_segments || others.counts.cubic_curve_spans) feat |= FEAT_CURVE;
    if (others.counts.instances) feat |= FEAT_INSTANCE;
    for (crt_scene* c : s->children) feat |= c->features;
    s->features = feat;
    if (n_tris == 0 && M == 0) {  // nothing to intersect: the kernels treat n_nodes == 0 as an empty scene
        DevScene d{};
        CKB(cudaMalloc((void**)&s->d_desc, sizeof(DevScene)));
        CKB(cudaMemcpy(s->d_desc, &d, sizeof d, cudaMemcpyHostToDevice));
        return CRT_OK;
    }
    // ---- arena 1: the upload and everything that is per primitive; one pinned buffer for the upload --------------
    const uint32_t NT = (uint32_t)n_tris;       // triangle slots (skipped triangles keep theirs)
    const uint32_t N = NT + M;                  // primitive slots: triangles, then the others in input order
    const size_t other_quads = others.recs.size() / 4;
    const size_t tab_words = (size_t)5 * n_meshes + 1;
    size_t scan_tmp = 0;
    CKB(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)N));
    const size_t upload_bytes = padded(3 * n_v, 4) + padded(3 * n_tris, 4) + padded(tab_words, 4)  //
                                + 2 * padded(M, 16) + padded(M, 8) + padded((size_t)M + 1, 4) + padded(other_quads, 16);
    const size_t need1 = upload_bytes + 5 * padded(N, 16) + 2 * padded(N, 4) + padded(kCounters, 4) + padded(scan_tmp, 1) + padded(M, 4);
    std::lock_guard<std::mutex> arena_lock(g_arena.mu);
    auto grow = [](char*& p, size_t& cap, size_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, need);
        if (e == cudaSuccess) cap = need;
        return e;
    };
    CKB(grow(g_arena.dev, g_arena.dev_cap, need1));
    if (upload_bytes > g_arena.pinned_cap) {
        if (g_arena.pinned) cudaFreeHost(g_arena.pinned);
        g_arena.pinned = nullptr;
        g_arena.pinned_cap = 0;
        CKB(cudaMallocHost((void**)&g_arena.pinned, upload_bytes));
        g_arena.pinned_cap = upload_bytes;
    }
    Carver dc{g_arena.dev, g_arena.dev + g_arena.dev_cap};
    float* d_verts = dc.take<float>(3 * n_v);
    uint32_t* d_idx = dc.take<uint32_t>(3 * n_tris);
    uint32_t* d_tab = dc.take<uint32_t>(tab_words);
    float4* d_olo = dc.take<float4>(M);
    float4* d_ohi = dc.take<float4>(M);
    uint2* d_oids = dc.take<uint2>(M);
    uint32_t* d_orec_off = dc.take<uint32_t>((size_t)M + 1);
    float4* d_orecs = dc.take<float4>(other_quads);
    {
        Carver hc{g_arena.pinned, g_arena.pinned + g_arena.pinned_cap};
        float* h_verts = hc.take<float>(3 * n_v);
        uint32_t* h_idx = hc.take<uint32_t>(3 * n_tris);
        uint32_t* h_tab = hc.take<uint32_t>(tab_words);
        float* h_olo = (float*)hc.take<float4>(M);
        float* h_ohi = (float*)hc.take<float4>(M);
        uint32_t* h_oids = (uint32_t*)hc.take<uint2>(M);
        uint32_t* h_orec_off = hc.take<uint32_t>((size_t)M + 1);
        uint32_t* h_orecs = (uint32_t*)hc.take<float4>(other_quads);
        if (M) {
            memcpy(h_olo, others.lo.data(), others.lo.size() * 4);
            memcpy(h_ohi, others.hi.data(), others.hi.size() * 4);
            memcpy(h_oids, others.ids.data(), others.ids.size() * 4);
            memcpy(h_orecs, others.recs.data(), others.recs.size() * 4);
        }
        memcpy(h_orec_off, others.rec_off.data(), others.rec_off.size() * 4);
        size_t vo = 0, io = 0;
        for (const Geom& G : geoms) {
            if (G.type != G_MESH || G.indices.empty()) continue;
            memcpy(h_verts + vo, G.verts.data(), G.verts.size() * 4);
            memcpy(h_idx + io, G.indices.data(), G.indices.size() * 4);
            vo += G.verts.size();
            io += G.indices.size();
        }
        uint32_t* t = h_tab;
        for (auto* vec : {&tri_offset, &vert_offset, &n_verts, &geom_id, &mask}) {
            memcpy(t, vec->data(), vec->size() * 4);
            t += vec->size();
        }
        CKB(cudaMemcpyAsync(g_arena.dev, g_arena.pinned, upload_bytes, cudaMemcpyHostToDevice, 0));  // the three regions are laid out identically
    }
    MeshTable mt{d_tab, d_tab + n_meshes + 1, d_tab + 2 * (size_t)n_meshes + 1, d_tab + 3 * (size_t)n_meshes + 1, d_tab + 4 * (size_t)n_meshes + 1, n_meshes};
    lap("concat + upload");
    auto t_up = std::chrono::steady_clock::now();

    BuildArrays A{};
    A.v0 = dc.take<float4>(N), A.v1 = dc.take<float4>(N), A.v2 = dc.take<float4>(N), A.lo = dc.take<float4>(N), A.hi = dc.take<float4>(N);
    A.cells = dc.take<uint32_t>(N), A.cell_off = dc.take<uint32_t>(N);
    A.counters = dc.take<unsigned int>(kCounters);
    A.n_tris = NT, A.other_rec_off = d_orec_off, A.other_recs = d_orecs, A.other_pos = dc.take<uint32_t>(M);
    void* d_scan = dc.take<char>(scan_tmp);
    uint2* d_prim_ids = nullptr;  // owned by the scene
    CKB(dmalloc(&d_prim_ids, N));
    s->d_prim_ids = d_prim_ids;  // from here on the scene owns it: every error return below frees it with the scene
    A.prim_ids = d_prim_ids;
    {
        unsigned int init[kCounters] = {0};
        for (int k = 8; k < 11; ++k) init[k] = 0xFFFFFFFFu;
        CKB(cudaMemcpy(A.counters, init, sizeof init, cudaMemcpyHostToDevice));
    }
    const int grid_t = (int)((N + kB - 1) / kB);
    uint64_t launches = 0;
    if (NT) {
        tri_setup_kernel<<<(int)((NT + kB - 1) / kB), kB>>>(NT, mt, d_verts, d_idx, A);
        ++launches;
    }
    if (M) {
        other_setup_kernel<<<(int)((M + kB - 1) / kB), kB>>>(M, d_olo, d_ohi, d_oids, A);
        ++launches;
    }
    unsigned int h_counters[kCounters];
    CKB(cudaMemcpy(h_counters, A.counters, sizeof h_counters, cudaMemcpyDeviceToHost));
    const uint32_t n_valid_tris = h_counters[0];
    const uint32_t n_valid = n_valid_tris + M;
    s->n_prims = (uint64_t)n_valid;
    s->breakdown.triangles = (uint64_t)n_valid_tris;
    s->non_empty = n_valid > 0;
    Box root{{ord2f(h_counters[8]), ord2f(h_counters[9]), ord2f(h_counters[10])}, {ord2f(h_counters[11]), ord2f(h_counters[12]), ord2f(h_counters[13])}};
    if (n_valid_tris == 0) root = Box{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};  // the counters still hold their sentinels
    for (uint32_t j = 0; j < M; ++j) {  // the other primitives' boxes join the scene bounds (they are few: on the host)
        root.lo.x = std::min(root.lo.x, others.lo[4 * j]), root.lo.y = std::min(root.lo.y, others.lo[4 * j + 1]), root.lo.z = std::min(root.lo.z, others.lo[4 * j + 2]);
        root.hi.x = std::max(root.hi.x, others.hi[4 * j]), root.hi.y = std::max(root.hi.y, others.hi[4 * j + 1]), root.hi.z = std::max(root.hi.z, others.hi[4 * j + 2]);
    }
    if (M) {  // the Morton normalisation of split_kernel reads the bounds from the counters
        auto f2o = [](float f) {
            unsigned int u;
            memcpy(&u, &f, 4);
            return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        };
        unsigned int b6[6] = {f2o(root.lo.x), f2o(root.lo.y), f2o(root.lo.z), f2o(root.hi.x), f2o(root.hi.y), f2o(root.hi.z)};
        CKB(cudaMemcpy(A.counters + 8, b6, sizeof b6, cudaMemcpyHostToDevice));
    }
    if (n_valid > 0) s->root = root;
    {  // scene.rs:446-455: mean / max primitive diagonal
        double tri_sum;
        float tri_max;
        memcpy(&tri_sum, h_counters + 16, 8);
        memcpy(&tri_max, h_counters + 18, 4);
        s->extent_sum = (float)(tri_sum + (double)others.extent_sum);
        s->extent_max = std::max(tri_max, others.extent_max);
    }
    lap("primitive setup");

    DevNode* d_nodes = nullptr;
    float4* d_leafs = nullptr;
    uint32_t n4 = 0;
    int n = 0;             // references the tree is built over
    size_t leaf_quads = 0;
    if (n_valid > 0) {
        // ---- early split clipping: how large is a typical triangle, which ones are much larger -----------------------
        const float scene_extent = std::max(std::max(root.hi.x - root.lo.x, root.hi.y - root.lo.y), std::max(root.hi.z - root.lo.z, 1e-30f));
        unsigned long long ext_sum = 0;
        if (NT) {
            extent_kernel<<<(int)((NT + kB - 1) / kB), kB>>>(NT, scene_extent, A);
            CKB(cudaMemcpy(&ext_sum, A.counters + 4, 8, cudaMemcpyDeviceToHost));
        }
        const float mean_extent = n_valid_tris ? (float)((double)ext_sum / 16777216.0 / (double)n_valid_tris) * scene_extent : 0.0f;
        const float cell = std::max(2.0f * mean_extent, scene_extent * (1.0f / 4096.0f));
        count_kernel<<<grid_t, kB>>>(N, cell, 2.0f * cell, A);
        uint32_t last_count = 0, last_off = 0;
        CKB(cudaMemcpy(&last_count, A.cell_off + (N - 1), 4, cudaMemcpyDeviceToHost));
        size_t scan_bytes = scan_tmp;
        CKB(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, A.cell_off, A.cell_off, (int)N));
        CKB(cudaMemcpy(&last_off, A.cell_off + (N - 1), 4, cudaMemcpyDeviceToHost));
        const uint64_t T64 = (uint64_t)last_off + last_count;
        if (T64 > 0x2AAAAAA0ull) return set_error(CRT_ERR_OOM, "device build: too many split references");
        const uint32_t T = (uint32_t)T64;
        launches += 3;
        lap("extent + count + scan");

        // ---- arena 2: everything per reference ------------------------------------------------------------------------
        size_t sort_tmp = 0;
        {
            cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
            cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
            CKB(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, k, v, (int)T, 0, 64));
        }
        size_t scan_tmp2 = 0;
        CKB(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp2, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)T));
        const size_t need2 = padded(T, 4) + 2 * padded(T, 16) + 2 * padded(T, 8) + 2 * padded(T, 4) + padded(sort_tmp, 1) + padded(scan_tmp2, 1)  //
                             + 4 * padded(T, 4) + padded(2 * (size_t)T, 4) + 2 * padded(2 * (size_t)T, 16) + 3 * padded(T, 4) + 2 * padded(T, 4)  //
                             + 2 * padded(T, 4);
        CKB(grow(g_arena.dev2, g_arena.dev2_cap, need2));
        Carver rc{g_arena.dev2, g_arena.dev2 + g_arena.dev2_cap};
        A.ref_prim = rc.take<uint32_t>(T), A.rlo = rc.take<float4>(T), A.rhi = rc.take<float4>(T);
        A.key_a = rc.take<unsigned long long>(T), A.key_b = rc.take<unsigned long long>(T), A.idx_a = rc.take<uint32_t>(T), A.idx_b = rc.take<uint32_t>(T);
        void* d_tmp = rc.take<char>(sort_tmp);
        void* d_scan2 = rc.take<char>(scan_tmp2);
        split_kernel<<<(int)((T + kB - 1) / kB), kB>>>(N, T, 1e-6f * scene_extent, A);
        lap("split");
        cub::DoubleBuffer<unsigned long long> keys(A.key_a, A.key_b);
        cub::DoubleBuffer<uint32_t> vals(A.idx_a, A.idx_b);
        size_t tmp_bytes = sort_tmp;
        CKB(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys, vals, (int)T, 0, 64));
        CKB(cudaMemcpy(h_counters, A.counters, sizeof h_counters, cudaMemcpyDeviceToHost));
        lap("radix sort");
        n = (int)h_counters[3];  // valid references: the first n sorted positions
        const unsigned long long* d_keys = keys.Current();
        const uint32_t* d_sorted = vals.Current();
        launches += 2;
        // where every leaf's records go: the record sizes differ by kind, so a prefix sum over the sorted references
        A.rec_size = rc.take<uint32_t>(T), A.rec_at = rc.take<uint32_t>(T);
        if (!A.rec_at) return set_error(CRT_ERR_OOM, "device build: scratch arena too small (internal sizing error)");
        size_t rec_quads = 0;
        if (n > 0) {
            rec_size_kernel<<<(n + kB - 1) / kB, kB>>>(n, d_sorted, A);
            size_t sb = scan_tmp2;
            CKB(cub::DeviceScan::ExclusiveSum(d_scan2, sb, A.rec_size, A.rec_at, n));
            uint32_t last_at = 0, last_size = 0;
            CKB(cudaMemcpy(&last_at, A.rec_at + (n - 1), 4, cudaMemcpyDeviceToHost));
            CKB(cudaMemcpy(&last_size, A.rec_size + (n - 1), 4, cudaMemcpyDeviceToHost));
            rec_quads = (size_t)last_at + last_size;
            launches += 2;
        }
        if (rec_quads > 0x7FFFFFF0ull) return set_error(CRT_ERR_OOM, "device build: leaf record stream exceeds the 31-bit offset of a leaf word");
        leaf_quads = rec_quads + 4;
        CKB(dmalloc(&d_leafs, leaf_quads));
        s->d_leafs = d_leafs;
        pad_kernel<<<1, 1>>>(d_leafs, rec_quads);
        if (n <= (int)kLeafMax) {
            CKB(dmalloc(&d_nodes, 1));
            s->d_nodes = d_nodes;
            tiny_kernel<<<1, 1>>>(n, d_sorted, d_nodes, d_leafs, A);
            n4 = 1;
            launches += 2;
        } else {
            const int n_bin = 2 * n - 1;
            A.left = rc.take<int>(n), A.right = rc.take<int>(n), A.first = rc.take<int>(n), A.last = rc.take<int>(n);
            A.parent = rc.take<int>(n_bin), A.nlo = rc.take<float4>(n_bin), A.nhi = rc.take<float4>(n_bin);
            A.visits = rc.take<unsigned int>(n), A.is_root4 = rc.take<uint32_t>(n), A.id4 = rc.take<uint32_t>(n);
            A.frontier_a = rc.take<int>(n), A.frontier_b = rc.take<int>(n);
            if (!A.frontier_b) return set_error(CRT_ERR_OOM, "device build: scratch arena too small (internal sizing error)");
            CKB(cudaMemsetAsync(A.visits, 0, (size_t)n * 4));
            CKB(cudaMemsetAsync(A.is_root4, 0, (size_t)n * 4));
            const int grid_n = (n + kB - 1) / kB;
            karras_kernel<<<grid_n, kB>>>(n, d_keys, A);
            refit_kernel<<<grid_n, kB>>>(n, d_sorted, A);
            lap("karras + refit");
            static const int top_cluster = [] {
                const char* e = getenv("CRT_DEV_SAH_TOP");
                return e ? atoi(e) : 0;
            }();
            if (top_cluster >= 8 && n > 4 * top_cluster) {
                struct DevBuf {  // freed on every way out of this block
                    void* p = nullptr;
                    ~DevBuf() {
                        if (p) cudaFree(p);
                    }
                };
                const size_t cap = (size_t)(8 * (size_t)n / (size_t)top_cluster) + 64;  // a balanced tree has about 2 n / cluster of them
                DevBuf b_top, b_cl, b_cnt, b_tn;
                CKB(cudaMalloc(&b_top.p, cap * sizeof(int)));
                CKB(cudaMalloc(&b_cl.p, (cap + 1) * sizeof(TopCluster)));
                CKB(cudaMalloc(&b_cnt.p, 8));
                int* d_top = (int*)b_top.p;
                TopCluster* d_cl = (TopCluster*)b_cl.p;
                unsigned int* d_cnt = (unsigned int*)b_cnt.p;
                CKB(cudaMemset(d_cnt, 0, 8));
                top_select_kernel<<<grid_n, kB>>>(n, top_cluster, (unsigned int)cap, A, d_top, d_cl, d_cnt);
                unsigned int cnt[2] = {0, 0};
                CKB(cudaMemcpy(cnt, d_cnt, 8, cudaMemcpyDeviceToHost));
                if (cnt[0] + 1 != cnt[1]) return set_error(CRT_ERR_INVALID_ARG, "device build: SAH top selection is inconsistent");
                launches += 1;
                if (cnt[1] <= cap) {  // else: a lopsided radix tree with far more large nodes than expected keeps its own top
                    std::vector<int> slots(cnt[0]);
                    std::vector<TopCluster> cl(cnt[1]);
                    CKB(cudaMemcpy(slots.data(), d_top, slots.size() * sizeof(int), cudaMemcpyDeviceToHost));
                    CKB(cudaMemcpy(cl.data(), d_cl, cl.size() * sizeof(TopCluster), cudaMemcpyDeviceToHost));
                    std::sort(slots.begin(), slots.end());  // the kernel appended them in no particular order
                    std::sort(cl.begin(), cl.end(), [](const TopCluster& a, const TopCluster& b) { return a.first < b.first; });
                    std::vector<int> idx(cl.size());
                    for (size_t k = 0; k < idx.size(); ++k) idx[k] = (int)k;
                    TopBuilder tb{cl, slots};
                    tb.out.reserve(slots.size());
                    tb.build(idx, 0, idx.size());
                    CKB(cudaMalloc(&b_tn.p, tb.out.size() * sizeof(TopNode)));
                    CKB(cudaMemcpy(b_tn.p, tb.out.data(), tb.out.size() * sizeof(TopNode), cudaMemcpyHostToDevice));
                    top_apply_kernel<<<(int)((tb.out.size() + kB - 1) / kB), kB>>>(n, (int)tb.out.size(), (const TopNode*)b_tn.p, A);
                    CKB(cudaDeviceSynchronize());
                    launches += 1;
                }
                lap("SAH top");
            }
            // level-synchronous marking of the binary nodes that become BVH4 nodes
            int zero = 0;
            CKB(cudaMemcpy(A.frontier_a, &zero, 4, cudaMemcpyHostToDevice));
            uint32_t count = 1;
            int *cur = A.frontier_a, *nxt = A.frontier_b;
            uint64_t levels = 0;
            while (count > 0) {
                CKB(cudaMemsetAsync(A.counters + 1, 0, 4));
                mark_kernel<<<(int)((count + kB - 1) / kB), kB>>>(n, cur, count, nxt, A);
                CKB(cudaMemcpy(&count, A.counters + 1, 4, cudaMemcpyDeviceToHost));
                std::swap(cur, nxt);
                ++levels;
            }
            if (trace) fprintf(stderr, "[device_build] %u triangles, %u candidate / %d valid references, cell %.4g, %llu mark levels\n", n_valid, T, n, cell,
                               (unsigned long long)levels);
            lap("mark (level loop)");
            size_t scan_bytes2 = scan_tmp2;
            CKB(cub::DeviceScan::ExclusiveSum(d_scan2, scan_bytes2, A.is_root4, A.id4, n - 1));
            uint32_t last_flag = 0, last_id = 0;
            CKB(cudaMemcpy(&last_flag, A.is_root4 + (n - 2), 4, cudaMemcpyDeviceToHost));
            CKB(cudaMemcpy(&last_id, A.id4 + (n - 2), 4, cudaMemcpyDeviceToHost));
            n4 = last_id + last_flag;
            CKB(dmalloc(&d_nodes, n4));
            s->d_nodes = d_nodes;
            emit_kernel<<<grid_n, kB>>>(n, d_sorted, d_nodes, d_leafs, A);
            launches += levels + 4;
            lap("scan + emit");
        }
        CKB(cudaDeviceSynchronize());
        CKB(cudaMemcpy(h_counters, A.counters, sizeof h_counters, cudaMemcpyDeviceToHost));
    }
    if (d_nodes) CKB(mark_empty_lanes(d_nodes, n4, 0));
    crt::count_launches(launches);
    // ---- hand the arrays to the scene (the scratch arena stays for the next build) ---------------------------------------------------
    s->d_nodes = d_nodes;
    s->d_leafs = d_leafs;
    s->d_prim_ids = d_prim_ids;
    s->desc.nodes = d_nodes;
    s->desc.leafs = (const U4*)d_leafs;
    s->desc.prim_ids = (const U2*)d_prim_ids;
    s->desc.normals = nullptr;
    s->desc.motion = nullptr;
    if (!others.motion.empty()) {
        CKB(cudaMalloc(&s->d_motion, others.motion.size() * sizeof(MotionXform)));
        CKB(cudaMemcpy(s->d_motion, others.motion.data(), others.motion.size() * sizeof(MotionXform), cudaMemcpyHostToDevice));
        s->desc.motion = (const MotionXform*)s->d_motion;
    }
    s->desc.n_nodes = n4;
    s->desc.has_triangles = n_valid_tris > 0;
    s->desc.shear_at_top = s->desc.has_triangles || translated_copy;
    s->n_nodes = n4, s->n_leaf_quads = n > 0 ? leaf_quads : 0, s->n_prim_ids = N, s->n_normal_floats = 0, s->n_motion = others.motion.size();
    if (!others.inst_prim.empty()) {  // where the instance records landed (crt_scene_export / import re-point them)
        std::vector<uint32_t> pos(M);
        CKB(cudaMemcpy(pos.data(), A.other_pos, (size_t)M * 4, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < others.inst_prim.size(); ++k) {
            s->inst_q.push_back(pos[others.inst_prim[k]]);
            s->inst_child.push_back((uint32_t)(std::find(s->children.begin(), s->children.end(), others.inst_child[k]) - s->children.begin()));
        }
    }
    CKB(cudaMalloc((void**)&s->d_desc, sizeof(DevScene)));
    CKB(cudaMemcpy(s->d_desc, &s->desc, sizeof(DevScene), cudaMemcpyHostToDevice));
    auto t1 = std::chrono::steady_clock::now();
    s->info.primitives = (uint64_t)n_valid;
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
    s->footprint.prim_table = (uint64_t)N * 8 + others.motion.size() * sizeof(MotionXform);
    s->footprint.descriptors = sizeof(DevScene);
    return CRT_OK;
}

}  // namespace crt
