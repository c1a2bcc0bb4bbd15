// Device-side BVH4 traversal for sm_100a: closest-hit and any-hit over the
// layout of crt_layout.h, one ray per thread, persistent warps that refill idle
// lanes from a global ray counter.
//
// Semantics follow crates/crust-rt/src of the reference exactly (same tree, same
// visiting order, same IEEE f32 arithmetic; this TU is compiled with
// -fmad=false because the reference never contracts FMAs, docs/simd.md:241-245):
//   bvh.rs:441-509   Bvh::hit           -> trace<false>
//   bvh.rs:585-611   Bvh::hit_any       -> trace<true>
//   bvh.rs:514-572   intersect_leaf     -> leaf record loop, packet / fallback order
//   bvh.rs:662-668   safe_inv3, :790-808 slab4
//   triangle.rs:41-80, 110-172, 276-348  Woop watertight test
//   prim.rs:132-169 sphere, :345-378 instance, curve.rs:17-217 curves
//
// Differences that do not change results:
//   * a popped internal node whose entry distance already exceeds `closest` is
//     skipped instead of fetched: its lanes would all fail `tnear <= tfar`
//     (child boxes are subsets of the parent box and the slab arithmetic is
//     monotone), so the reference visits it and finds nothing;
//   * leaf lanes are queued on the same stack (above the node's internal
//     lanes, near one on top) instead of being intersected inside the node
//     visit — same order;
//   * (u, v) and the normal are recomputed once at the end from the winning
//     record instead of at every accepted hit;
//   * the slab planes are picked by the ray's direction signs instead of by
//     min / max (tests/test_slab_signselect.py);
//   * during traversal the triangle test works in the shear frame without the
//     reference's kx / ky swap (tri_eval below, tests/test_tri_frame_symmetry.py);
//     the reported hit is re-evaluated in the reference's frame (tri_eval_exact);
//   * for rays with t_min >= +0 the push order comes from unsigned compares of
//     the entry distances' bit patterns (same order: they are non-negative).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crt.h"
#include "crt_layout.h"
#include "launch.h"

namespace crt {

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ F3 operator/(F3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
__device__ __forceinline__ F3 normalize(F3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ float pick(F3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
// a / b / c for k = 0 / 1 / 2, as two selects (never a branch: the callers sit in the divergent leaf phase, where nvcc
// otherwise turns a run of picks by the same axis number into nine divergent branch regions)
__device__ __forceinline__ float sel3(int k, float a, float b, float c) {
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\t.reg .f32 bc;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\t"
        "selp.f32 bc, %2, %3, p1;\n\tselp.f32 %0, %1, bc, p0;\n\t}"
        : "=f"(r)
        : "f"(a), "f"(b), "f"(c), "r"(k));
    return r;
}
__device__ __forceinline__ F3 xyz(float4 q) { return {q.x, q.y, q.z}; }

struct Xf {  // glam Affine3A column order
    F3 x, y, z, t;
};
__device__ __forceinline__ F3 xf_vec(const Xf& m, F3 v) { return (m.x * v.x + m.y * v.y) + m.z * v.z; }
__device__ __forceinline__ F3 xf_point(const Xf& m, F3 p) { return xf_vec(m, p) + m.t; }
__device__ __forceinline__ F3 lerp3(F3 a, F3 b, float s) { return a * (1.0f - s) + b * s; }
__device__ inline Xf xf_inverse(const Xf& a) {  // glam Mat3A::inverse + Affine3A::inverse
    F3 t0 = cross(a.y, a.z), t1 = cross(a.z, a.x), t2 = cross(a.x, a.y);
    float inv_det = 1.0f / dot(a.z, t2);
    F3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    Xf r;
    r.x = {c0.x, c1.x, c2.x};
    r.y = {c0.y, c1.y, c2.y};
    r.z = {c0.z, c1.z, c2.z};
    F3 mt = xf_vec(r, a.t);
    r.t = {-mt.x, -mt.y, -mt.z};
    return r;
}
__device__ __forceinline__ F3 ld3(const float* p) { return {p[0], p[1], p[2]}; }

// World-to-local of an instance record at the ray's shutter time (prim.rs:323-331).
__device__ inline Xf instance_w2l(const DevScene& owner, const float4* rec, float time) {
    float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
    uint32_t flags = __float_as_uint(q0.w);
    if ((flags & REC_MOTION) && time > 0.0f) {
        const MotionXform* m = owner.motion + __float_as_uint(q3.w);
        Xf a{ld3(m->l2w), ld3(m->l2w + 3), ld3(m->l2w + 6), ld3(m->l2w + 9)};
        Xf b{ld3(m->l2w_end), ld3(m->l2w_end + 3), ld3(m->l2w_end + 6), ld3(m->l2w_end + 9)};
        Xf l{lerp3(a.x, b.x, time), lerp3(a.y, b.y, time), lerp3(a.z, b.z, time), lerp3(a.t, b.t, time)};  // prim.rs:285-294
        return xf_inverse(l);
    }
    return Xf{xyz(q0), xyz(q1), xyz(q2), xyz(q3)};
}
__device__ __forceinline__ const DevScene* instance_child(const float4* rec) {
    float4 q4 = __ldg(rec + 4);
    return (const DevScene*)(((uint64_t)__float_as_uint(q4.y) << 32) | (uint64_t)__float_as_uint(q4.x));
}
// The child's leaf-record array, carried by the instance record itself (q5.zw) next to its node array (q5.xy).
__device__ __forceinline__ const float4* instance_child_leafs(const float4* rec) {
    uint4 q5 = __ldg((const uint4*)(rec + 5));
    return (const float4*)(((uint64_t)q5.w << 32) | (uint64_t)q5.z);
}

__device__ __forceinline__ float safe_inv(float d) {  // bvh.rs:662-668
    return fabsf(d) < 1e-20f ? copysignf(1e20f, d) : __frcp_rn(d);  // correctly rounded: same bits as 1.0f / d
}

// Per-level ray: origin/direction in the current scene's space plus the
// derived slab reciprocals and Woop shear (triangle.rs:41-80).
struct LevelRay {
    F3 o, d, inv;
    // Shear constants of triangle.rs:41-80 WITHOUT the reference's swap of kx / ky for a direction that is negative along kz:
    // sx = d[(kz+1)%3] / d[kz], sy = d[(kz+2)%3] / d[kz]. See tri_eval for why the hot path does not need the swap.
    float sx, sy, sz;
    int kz;
#if CRT_SIGNSEL
    // byte offsets, inside a 128-byte node, of the bounds vector holding each axis' near plane: the low faces (0, 16, 32)
    // for a ray travelling in +axis, the high faces (48, 64, 80) otherwise; the far plane is the other one of the pair
    uint32_t near_x, near_y, near_z;
#endif
    __device__ __forceinline__ uint32_t pack_axes() const {
        uint32_t w = (uint32_t)kz & 3u;  // the axis is unset (garbage) for a level without triangles
#if CRT_SIGNSEL
        w |= (near_x << 8) | (near_y << 16) | (near_z << 24);
#endif
        return w;
    }
    __device__ __forceinline__ void unpack_axes(uint32_t w) {
        kz = (int)(w & 3u);
#if CRT_SIGNSEL
        near_x = (w >> 8) & 0xFFu, near_y = (w >> 16) & 0xFFu, near_z = w >> 24;
#endif
    }
    __device__ __forceinline__ void set(F3 org, F3 dir, bool want_shear) {
        o = org;
        d = dir;
        inv = {safe_inv(dir.x), safe_inv(dir.y), safe_inv(dir.z)};
#if CRT_SIGNSEL
        // the sign of `inv` (never zero, never NaN for a finite direction), not of `dir`: -0.0 travels "backwards" like the reference's
        // copysign(1e20, d) (bvh.rs:662-668)
        near_x = inv.x < 0.0f ? 48u : 0u;
        near_y = inv.y < 0.0f ? 64u : 16u;
        near_z = inv.z < 0.0f ? 80u : 32u;
#endif
        if (want_shear) set_shear(dir);
    }
    __device__ __forceinline__ void set_shear(F3 dir) {
        float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
        kz = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
        float dz = sel3(kz, dir.x, dir.y, dir.z);
        sx = sel3(kz, dir.y, dir.z, dir.x) / dz;
        sy = sel3(kz, dir.z, dir.x, dir.y) / dz;
        sz = __frcp_rn(dz);
    }
};

struct TriEval {
    float e0, e1, e2, det, t_scaled;
    bool on_edge;
};
// Shared front half of triangle.rs:123-136 / 285-298, in the frame (kx, ky, kz) = ((kz+1)%3, (kz+2)%3, kz) — the reference's
// frame before it swaps kx and ky for d[kz] < 0. With the swap every product pair of an edge function trades places, so
// e0, e1, e2, det and t_scaled all come out exactly negated (IEEE multiplication commutes, a - b = -(b - a), round-to-nearest
// is symmetric), and everything traversal decides from them is blind to that common sign: the zero tests, mixed_signs, both
// range tests and t = t_scaled * (1 / det). Only the sign of a zero u / v / t could differ, and the reported hit is
// re-evaluated once with the swap in place (tri_eval_exact, finalize). Three selects per vertex component instead of
// picking by three independent axis numbers.
__device__ __forceinline__ TriEval tri_eval(const LevelRay& r, F3 v0, F3 v1, F3 v2) {
    const int z = r.kz;
    F3 a = v0 - r.o, b = v1 - r.o, c = v2 - r.o;
    float akz = sel3(z, a.x, a.y, a.z), bkz = sel3(z, b.x, b.y, b.z), ckz = sel3(z, c.x, c.y, c.z);
    float ax = sel3(z, a.y, a.z, a.x) - r.sx * akz, ay = sel3(z, a.z, a.x, a.y) - r.sy * akz;
    float bx = sel3(z, b.y, b.z, b.x) - r.sx * bkz, by = sel3(z, b.z, b.x, b.y) - r.sy * bkz;
    float cx = sel3(z, c.y, c.z, c.x) - r.sx * ckz, cy = sel3(z, c.z, c.x, c.y) - r.sy * ckz;
    TriEval o;
    o.e0 = bx * cy - by * cx;
    o.e1 = cx * ay - cy * ax;
    o.e2 = ax * by - ay * bx;
    o.on_edge = (o.e0 == 0.0f) | (o.e1 == 0.0f) | (o.e2 == 0.0f);
    o.det = (o.e0 + o.e1) + o.e2;
    o.t_scaled = (o.e0 * (r.sz * akz) + o.e1 * (r.sz * bkz)) + o.e2 * (r.sz * ckz);
    return o;
}
// The f64 re-evaluation of triangle.rs:141-145 (scalar path, rare); same frame as tri_eval.
// (The ray's fields travel by value: a reference would force the whole LevelRay of the traversal loop to keep a
// local-memory image, stored again at every refill and instance descent.)
__device__ __noinline__ TriEval tri_eval_f64(F3 ro, float rsx, float rsy, float rsz, int kz, F3 v0, F3 v1, F3 v2) {
    const int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    struct { float sx, sy, sz; int kz; } r = {rsx, rsy, rsz, kz};
    F3 a = v0 - ro, b = v1 - ro, c = v2 - ro;
    float akz = pick(a, r.kz), bkz = pick(b, r.kz), ckz = pick(c, r.kz);
    float ax = pick(a, kx) - r.sx * akz, ay = pick(a, ky) - r.sy * akz;
    float bx = pick(b, kx) - r.sx * bkz, by = pick(b, ky) - r.sy * bkz;
    float cx = pick(c, kx) - r.sx * ckz, cy = pick(c, ky) - r.sy * ckz;
    TriEval o;
    o.e0 = (float)__dsub_rn(__dmul_rn((double)bx, (double)cy), __dmul_rn((double)by, (double)cx));
    o.e1 = (float)__dsub_rn(__dmul_rn((double)cx, (double)ay), __dmul_rn((double)cy, (double)ax));
    o.e2 = (float)__dsub_rn(__dmul_rn((double)ax, (double)by), __dmul_rn((double)ay, (double)bx));
    o.on_edge = false;
    o.det = (o.e0 + o.e1) + o.e2;
    o.t_scaled = (o.e0 * (r.sz * akz) + o.e1 * (r.sz * bkz)) + o.e2 * (r.sz * ckz);
    return o;
}
// The reference's own frame (triangle.rs:47-60: kx and ky swapped when d[kz] < 0), for the one evaluation per ray whose
// (t, u, v) are reported: bit-for-bit the reference's values, signs of zeros included. Once per ray, not in the hot loop.
__device__ __noinline__ TriEval tri_eval_exact(F3 ro, F3 rd, float sx, float sy, float rsz, int kz, F3 v0, F3 v1, F3 v2) {
    int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    if (pick(rd, kz) < 0.0f) {
        int t = kx; kx = ky; ky = t;
        float ts = sx; sx = sy; sy = ts;
    }
    F3 a = v0 - ro, b = v1 - ro, c = v2 - ro;
    float akz = pick(a, kz), bkz = pick(b, kz), ckz = pick(c, kz);
    float ax = pick(a, kx) - sx * akz, ay = pick(a, ky) - sy * akz;
    float bx = pick(b, kx) - sx * bkz, by = pick(b, ky) - sy * bkz;
    float cx = pick(c, kx) - sx * ckz, cy = pick(c, ky) - sy * ckz;
    TriEval o;
    o.e0 = bx * cy - by * cx;
    o.e1 = cx * ay - cy * ax;
    o.e2 = ax * by - ay * bx;
    o.on_edge = (o.e0 == 0.0f) | (o.e1 == 0.0f) | (o.e2 == 0.0f);
    if (o.on_edge) {
        o.e0 = (float)__dsub_rn(__dmul_rn((double)bx, (double)cy), __dmul_rn((double)by, (double)cx));
        o.e1 = (float)__dsub_rn(__dmul_rn((double)cx, (double)ay), __dmul_rn((double)cy, (double)ax));
        o.e2 = (float)__dsub_rn(__dmul_rn((double)ax, (double)by), __dmul_rn((double)ay, (double)bx));
    }
    o.det = (o.e0 + o.e1) + o.e2;
    o.t_scaled = (o.e0 * (rsz * akz) + o.e1 * (rsz * bkz)) + o.e2 * (rsz * ckz);
    return o;
}
__device__ __forceinline__ bool mixed_signs(const TriEval& e) {
    return ((e.e0 < 0.0f) | (e.e1 < 0.0f) | (e.e2 < 0.0f)) & ((e.e0 > 0.0f) | (e.e1 > 0.0f) | (e.e2 > 0.0f));
}
// Packet range test, triangle.rs:327-331 (inclusive, division-free).
__device__ __forceinline__ bool in_range_packet(const TriEval& e, float t_min, float t_max) {
    float ad = fabsf(e.det);
    float ts = e.det < 0.0f ? -e.t_scaled : e.t_scaled;
    return (ts >= t_min * ad) & (ts <= t_max * ad);
}
// Scalar range test, triangle.rs:163-168.
__device__ __forceinline__ bool in_range_scalar(const TriEval& e, float t_min, float t_max) {
    if (e.det < 0.0f && (e.t_scaled > t_min * e.det || e.t_scaled < t_max * e.det)) return false;
    if (e.det > 0.0f && (e.t_scaled < t_min * e.det || e.t_scaled > t_max * e.det)) return false;
    return true;
}

// prim.rs:132-161. Exclusive range. Returns t or a negative sentinel via ok.
__device__ inline bool sphere_hit(const LevelRay& r, F3 c, float radius, float t_min, float t_max, float& t_out, F3& n_out) {
    F3 oc = r.o - c;
    float a = dot(r.d, r.d);
    float half_b = dot(oc, r.d);
    float cc = dot(oc, oc) - radius * radius;
    float disc = half_b * half_b - a * cc;
    if (disc < 0.0f) return false;
    float sq = sqrtf(disc);
    float root = (-half_b - sq) / a;
    if (root <= t_min || root >= t_max) {
        root = (-half_b + sq) / a;
        if (root <= t_min || root >= t_max) return false;
    }
    t_out = root;
    n_out = ((r.o + r.d * root) - c) / radius;
    return true;
}

// curve.rs:17-93
__device__ __noinline__ bool rounded_cone_hit(F3 o, F3 dir, F3 p0, F3 p1, float r0, float r1, float t_min, float t_max,
                                              float& t_out, F3& n_out) {
    r0 = fmaxf(r0, 1e-6f);
    r1 = fmaxf(r1, 1e-6f);
    float len = sqrtf(dot(dir, dir));
    if (len < 1e-20f) return false;
    F3 rd = dir / len;
    float n_min = t_min * len, n_max = t_max * len;
    F3 ba = p1 - p0, oa = o - p0, ob = o - p1;
    float rr = r0 - r1;
    float m0 = dot(ba, ba), m1 = dot(ba, oa), m2 = dot(ba, rd), m3 = dot(rd, oa), m5 = dot(oa, oa), m6 = dot(ob, rd),
          m7 = dot(ob, ob);
    bool have = false;
    float bt = 0.0f;
    F3 bn = f3(0, 0, 0);
#define CRT_CONSIDER(tt, nn)                                     \
    do {                                                         \
        float _t = (tt);                                         \
        if (_t >= n_min && _t <= n_max && (!have || _t < bt)) {  \
            bt = _t;                                             \
            bn = (nn);                                           \
            have = true;                                         \
        }                                                        \
    } while (0)
    float d2 = m0 - rr * rr;
    if (d2 > 0.0f) {
        float k2 = d2 - m2 * m2;
        float k1 = d2 * m3 - m1 * m2 + m2 * rr * r0;
        float k0 = d2 * m5 - m1 * m1 + m1 * rr * r0 * 2.0f - m0 * r0 * r0;
        float h = k1 * k1 - k0 * k2;
        if (h >= 0.0f && fabsf(k2) > 1e-12f) {
            float sq = sqrtf(h);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float t = i == 0 ? (-k1 - sq) / k2 : (-k1 + sq) / k2;
                float y = m1 - r0 * rr + t * m2;
                if (y > 0.0f && y < d2) CRT_CONSIDER(t, normalize((oa + rd * t) * d2 - ba * y));
            }
        }
    }
    float h0 = m3 * m3 - m5 + r0 * r0;
    if (h0 >= 0.0f) {
        float sq = sqrtf(h0);
        CRT_CONSIDER(-m3 - sq, (oa + rd * (-m3 - sq)) / r0);
        CRT_CONSIDER(-m3 + sq, (oa + rd * (-m3 + sq)) / r0);
    }
    float h1 = m6 * m6 - m7 + r1 * r1;
    if (h1 >= 0.0f) {
        float sq = sqrtf(h1);
        CRT_CONSIDER(-m6 - sq, (ob + rd * (-m6 - sq)) / r1);
        CRT_CONSIDER(-m6 + sq, (ob + rd * (-m6 + sq)) / r1);
    }
#undef CRT_CONSIDER
    if (!have) return false;
    t_out = bt / len;
    n_out = bn;
    return true;
}

// aabb.rs:24-42 scalar slab test (used by the cubic-curve subdivision only).
__device__ inline bool box_hit(F3 lo, F3 hi, F3 o, F3 d, float t_min, float t_max) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float inv_d = 1.0f / pick(d, a);
        float t0 = (pick(lo, a) - pick(o, a)) * inv_d;
        float t1 = (pick(hi, a) - pick(o, a)) * inv_d;
        if (inv_d < 0.0f) {
            float s = t0;
            t0 = t1;
            t1 = s;
        }
        t_min = fmaxf(t_min, t0);
        t_max = fminf(t_max, t1);
        if (t_max <= t_min) return false;
    }
    return true;
}
__device__ __forceinline__ F3 min3(F3 a, F3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
__device__ __forceinline__ F3 max3(F3 a, F3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }

// curve.rs:122-141
__device__ inline uint32_t flatness_depth(const F3 cp[4], float max_width) {
    F3 d0 = cp[0] - cp[1] * 2.0f + cp[2];
    F3 d1 = cp[1] - cp[2] * 2.0f + cp[3];
    float l0 = fmaxf(fmaxf(fmaxf(fabsf(d0.x), fabsf(d1.x)), fmaxf(fabsf(d0.y), fabsf(d1.y))), fmaxf(fabsf(d0.z), fabsf(d1.z)));
    if (l0 <= 0.0f || max_width <= 0.0f) return 0;
    float eps = max_width * 0.05f;
    float r0 = log2f(1.41421356237309504880f * 6.0f * l0 / (8.0f * eps)) * 0.5f;
    if (!isfinite(r0)) return 0;
    float f = fmaxf(floorf(r0), 0.0f);
    uint32_t d = f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f;
    return d < 10u ? d : 10u;
}

// curve.rs:166-217, the recursion unrolled onto an explicit stack. Visiting
// order (first half, then second half against the shrunk t_max) is preserved:
// a frame is (control points, u0, u1, depth) and the "best so far" is global to
// the span exactly as `best` threads through the recursive returns — every
// nested call's own `best` starts empty but its t_max argument is the caller's
// cur_t_max, which is what `tmax_stack` carries.
__device__ __noinline__ bool cubic_curve_hit(F3 o, F3 dir, const F3 cp_in[4], float r0_full, float r1_full, float t_min,
                                             float t_max, float& t_out, F3& n_out) {
    struct Frame {
        F3 cp[4];
        float u0, u1, t_max;
        uint32_t depth;
        int parent;      // frame index of the caller
        uint8_t stage;   // 0 = not expanded, 1 = first half done, 2 = both done
        bool has;        // this call's `best`
        float best_t;
        F3 best_n;
    };
    // Depth-first with at most one live frame per recursion level (<= 11).
    Frame fr[11];
    int top = 0;
    fr[0].cp[0] = cp_in[0], fr[0].cp[1] = cp_in[1], fr[0].cp[2] = cp_in[2], fr[0].cp[3] = cp_in[3];
    fr[0].u0 = 0.0f, fr[0].u1 = 1.0f, fr[0].t_max = t_max;
    fr[0].depth = flatness_depth(cp_in, 2.0f * fmaxf(r0_full, r1_full));
    fr[0].parent = -1, fr[0].stage = 0, fr[0].has = false, fr[0].best_t = 0.0f, fr[0].best_n = f3(0, 0, 0);
    bool result_has = false;
    float result_t = 0.0f;
    F3 result_n = f3(0, 0, 0);
    while (top >= 0) {
        Frame& f = fr[top];
        if (f.depth == 0) {
            float ra = r0_full + (r1_full - r0_full) * f.u0, rb = r0_full + (r1_full - r0_full) * f.u1;
            float t;
            F3 n;
            bool h = rounded_cone_hit(o, dir, f.cp[0], f.cp[3], ra, rb, t_min, f.t_max, t, n);
            // return to the caller
            if (top == 0) {
                result_has = h, result_t = t, result_n = n;
            } else if (h) {
                Frame& p = fr[top - 1];
                p.has = true, p.best_t = t, p.best_n = n;
            }
            --top;
            continue;
        }
        if (f.stage == 2) {  // both halves done: return `best`
            if (top == 0) {
                result_has = f.has, result_t = f.best_t, result_n = f.best_n;
            } else if (f.has) {
                Frame& p = fr[top - 1];
                p.has = true, p.best_t = f.best_t, p.best_n = f.best_n;
            }
            --top;
            continue;
        }
        // expand half `stage`
        int h = f.stage;
        f.stage += 1;
        F3 p01 = (f.cp[0] + f.cp[1]) * 0.5f, p12 = (f.cp[1] + f.cp[2]) * 0.5f, p23 = (f.cp[2] + f.cp[3]) * 0.5f;
        F3 p012 = (p01 + p12) * 0.5f, p123 = (p12 + p23) * 0.5f, p0123 = (p012 + p123) * 0.5f;
        float u_mid = (f.u0 + f.u1) * 0.5f;
        F3 s0, s1, s2, s3;
        float su0, su1;
        if (h == 0) {
            s0 = f.cp[0], s1 = p01, s2 = p012, s3 = p0123, su0 = f.u0, su1 = u_mid;
        } else {
            s0 = p0123, s1 = p123, s2 = p23, s3 = f.cp[3], su0 = u_mid, su1 = f.u1;
        }
        float cur_t_max = f.has ? f.best_t : f.t_max;
        float radius = fmaxf(r0_full + (r1_full - r0_full) * su0, r0_full + (r1_full - r0_full) * su1);
        F3 lo = min3(min3(min3(s0, s1), s2), s3) - f3(radius, radius, radius);
        F3 hi = max3(max3(max3(s0, s1), s2), s3) + f3(radius, radius, radius);
        if (!box_hit(lo, hi, o, dir, t_min, cur_t_max)) continue;
        Frame& c = fr[top + 1];
        c.cp[0] = s0, c.cp[1] = s1, c.cp[2] = s2, c.cp[3] = s3;
        c.u0 = su0, c.u1 = su1, c.t_max = cur_t_max, c.depth = f.depth - 1;
        c.parent = top, c.stage = 0, c.has = false, c.best_t = 0.0f, c.best_n = f3(0, 0, 0);
        ++top;
    }
    if (!result_has) return false;
    t_out = result_t;
    n_out = result_n;
    return true;
}

}  // namespace crt
