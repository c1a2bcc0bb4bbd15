// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17, scalar, one ray at a time) of the `crust-rt` kernel
// of doubleailes/crust-render. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it, and only as the
// checker or the timed CPU arm. The shipped library (crust-render_b200/csrc)
// has its own host builder and device kernels and never links this file.
//
// The Rust reference cannot be compiled in this environment (no cargo/rustc,
// crates not vendored), so this restatement is the oracle of record. Its
// fidelity is pinned by the reference's own known-answer tests, re-expressed in
// tests/test_oracle_*.py (see each test for the file:line it mirrors).
//
// Build flags that matter: -ffp-contract=off (the reference never contracts
// FMAs: docs/simd.md:241-245) and plain IEEE f32 SSE2 arithmetic.
//
// All `file:line` citations are relative to /root/reference/crates/crust-rt/src.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <future>
#include <limits>
#include <memory>
#include <optional>
#include <thread>
#include <unordered_set>
#include <vector>

namespace orc {

constexpr float INF = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------
// glam 0.33 Vec3A (SSE2 backend) restated lane-wise. dot = (x*x' + y*y') + z*z',
// normalize = v * (1/sqrt(dot)), min/max lane-wise (no NaNs reach them).
// ---------------------------------------------------------------------------
struct V3 {
    float x = 0, y = 0, z = 0;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 splat(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 vmin(V3 a, V3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float length_squared(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
inline float max_element(V3 a) { return std::max(std::max(a.x, a.y), a.z); }
// glam >= 0.27 lerp: self * (1 - s) + rhs * s  (unverifiable here)
inline V3 lerp(V3 a, V3 b, float s) { return a * (1.0f - s) + b * s; }

// glam Affine3A: column-major 3x3 + translation.
struct Affine {
    V3 x_axis{1, 0, 0}, y_axis{0, 1, 0}, z_axis{0, 0, 1}, translation{0, 0, 0};
};
// Mat3A * Vec3A = (x_axis*v.x + y_axis*v.y) + z_axis*v.z
inline V3 mul_mat3(const Affine& m, V3 v) { return (m.x_axis * v.x + m.y_axis * v.y) + m.z_axis * v.z; }
inline V3 transform_point(const Affine& m, V3 p) { return mul_mat3(m, p) + m.translation; }
inline V3 transform_vector(const Affine& m, V3 v) { return mul_mat3(m, v); }
// Mat3A::inverse (cross-product form) then Affine3A::inverse: t' = -(M^-1 * t)
inline Affine inverse(const Affine& a) {
    V3 t0 = cross(a.y_axis, a.z_axis);
    V3 t1 = cross(a.z_axis, a.x_axis);
    V3 t2 = cross(a.x_axis, a.y_axis);
    float det = dot(a.z_axis, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    Affine r;  // transpose of (c0 c1 c2)
    r.x_axis = {c0.x, c1.x, c2.x};
    r.y_axis = {c0.y, c1.y, c2.y};
    r.z_axis = {c0.z, c1.z, c2.z};
    r.translation = -mul_mat3(r, a.translation);
    return r;
}
// 3x3 only; columns of the transpose are the rows.
struct Mat3 {
    V3 x_axis, y_axis, z_axis;
};
inline Mat3 transpose3(const Affine& a) {
    return {{a.x_axis.x, a.y_axis.x, a.z_axis.x},
            {a.x_axis.y, a.y_axis.y, a.z_axis.y},
            {a.x_axis.z, a.y_axis.z, a.z_axis.z}};
}
inline V3 mul(const Mat3& m, V3 v) { return (m.x_axis * v.x + m.y_axis * v.y) + m.z_axis * v.z; }

// ray.rs:8-55
constexpr uint32_t MASK_CAMERA = 1u << 0, MASK_SHADOW = 1u << 1, MASK_INDIRECT = 1u << 2, MASK_ALL = 0xFFFFFFFFu;
constexpr uint32_t INVALID_ID = 0xFFFFFFFFu;  // lib.rs:53
struct Ray {
    V3 origin, dir;
    float time = 0.0f;
    uint32_t mask = MASK_ALL;
    V3 at(float t) const { return origin + t * dir; }
};

// aabb.rs:4-42
struct AABB {
    V3 minimum, maximum;
};
inline AABB surrounding_box(const AABB& a, const AABB& b) { return {vmin(a.minimum, b.minimum), vmax(a.maximum, b.maximum)}; }
inline bool aabb_hit(const AABB& b, const Ray& ray, float t_min, float t_max) {
    for (int a = 0; a < 3; ++a) {
        float inv_d = 1.0f / ray.dir[a];
        float t0 = (b.minimum[a] - ray.origin[a]) * inv_d;
        float t1 = (b.maximum[a] - ray.origin[a]) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        t_min = std::fmax(t_min, t0);  // f32::max ignores NaN like fmaxf
        t_max = std::fmin(t_max, t1);
        if (t_max <= t_min) return false;
    }
    return true;
}
// aabb.rs:47-58
inline AABB triangle_aabb(V3 v0, V3 v1, V3 v2) {
    V3 mn = vmin(vmin(v0, v1), v2), mx = vmax(vmax(v0, v1), v2);
    const float PAD = 1e-4f;
    for (int a = 0; a < 3; ++a)
        if (mx[a] - mn[a] < PAD) {
            mn[a] -= PAD;
            mx[a] += PAD;
        }
    return {mn, mx};
}

// ---------------------------------------------------------------------------
// triangle.rs
// ---------------------------------------------------------------------------
struct RayShear {  // triangle.rs:23-81
    int kx, ky, kz;
    float sx, sy, sz;
    V3 o;
    explicit RayShear(const Ray& ray) {
        V3 d = ray.dir, ad = vabs(d);
        kz = ad.x > ad.y ? (ad.x > ad.z ? 0 : 2) : (ad.y > ad.z ? 1 : 2);
        kx = (kz + 1) % 3;
        ky = (kz + 2) % 3;
        if (d[kz] < 0.0f) std::swap(kx, ky);
        sx = d[kx] / d[kz];
        sy = d[ky] / d[kz];
        sz = 1.0f / d[kz];
        o = ray.origin;
    }
};
struct TUV {
    float t, u, v;
};
// triangle.rs:110-172
inline std::optional<TUV> triangle_intersect_sheared(const RayShear& sh, V3 origin, V3 v0, V3 v1, V3 v2, float t_min,
                                                     float t_max) {
    int kx = sh.kx, ky = sh.ky, kz = sh.kz;
    float sx = sh.sx, sy = sh.sy, sz = sh.sz;
    V3 a = v0 - origin, b = v1 - origin, c = v2 - origin;
    float ax = a[kx] - sx * a[kz], ay = a[ky] - sy * a[kz];
    float bx = b[kx] - sx * b[kz], by = b[ky] - sy * b[kz];
    float cx = c[kx] - sx * c[kz], cy = c[ky] - sy * c[kz];
    float e0 = bx * cy - by * cx, e1 = cx * ay - cy * ax, e2 = ax * by - ay * bx;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        e0 = (float)((double)bx * (double)cy - (double)by * (double)cx);
        e1 = (float)((double)cx * (double)ay - (double)cy * (double)ax);
        e2 = (float)((double)ax * (double)by - (double)ay * (double)bx);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return std::nullopt;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return std::nullopt;
    float az = sz * a[kz], bz = sz * b[kz], cz = sz * c[kz];
    float t_scaled = e0 * az + e1 * bz + e2 * cz;
    if (det < 0.0f && (t_scaled > t_min * det || t_scaled < t_max * det)) return std::nullopt;
    if (det > 0.0f && (t_scaled < t_min * det || t_scaled > t_max * det)) return std::nullopt;
    float inv_det = 1.0f / det;
    return TUV{t_scaled * inv_det, e1 * inv_det, e2 * inv_det};
}
// triangle.rs:97-106
inline std::optional<TUV> triangle_intersect(const Ray& ray, V3 v0, V3 v1, V3 v2, float t_min, float t_max) {
    return triangle_intersect_sheared(RayShear(ray), ray.origin, v0, v1, v2, t_min, t_max);
}

struct TriIn {  // (v0, v1, v2, prim index, mask) tuple of Tri4::new
    V3 v0, v1, v2;
    uint32_t pi, mask;
};
struct Hit4 {  // triangle.rs:199-212
    uint32_t hits = 0, fallback = 0;
    float t[4] = {0, 0, 0, 0}, u[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
};
struct Tri4 {            // triangle.rs:182-253
    float v[3][3][4];    // v[vertex][axis][lane]
    uint32_t prim[4];
    uint32_t active, mask_and, mask_or, masks[4];
    Tri4(const TriIn* tris, size_t n) {
        active = 0;
        mask_and = 0xFFFFFFFFu;
        mask_or = 0;
        for (int lane = 0; lane < 4; ++lane) {
            prim[lane] = 0xFFFFFFFFu;
            masks[lane] = 0;
            const TriIn& t = tris[std::min<size_t>(lane, n - 1)];
            for (int axis = 0; axis < 3; ++axis) {
                v[0][axis][lane] = t.v0[axis];
                v[1][axis][lane] = t.v1[axis];
                v[2][axis][lane] = t.v2[axis];
            }
            if ((size_t)lane < n) {
                prim[lane] = t.pi;
                masks[lane] = t.mask;
                active |= 1u << lane;
                mask_and &= t.mask;
                mask_or |= t.mask;
            }
        }
    }
    uint32_t visible_lanes(uint32_t ray_mask) const {  // triangle.rs:257-271
        if (ray_mask & mask_and) return active;
        if ((ray_mask & mask_or) == 0) return 0;
        uint32_t m = 0;
        for (int lane = 0; lane < 4; ++lane)
            if ((active & (1u << lane)) && (masks[lane] & ray_mask)) m |= 1u << lane;
        return m;
    }
    // triangle.rs:276-348 — lane-wise transcription of the Vec4 code.
    Hit4 intersect(const RayShear& sh, uint32_t ray_mask, float t_min, float t_max) const {
        Hit4 out;
        uint32_t m = visible_lanes(ray_mask);
        if (m == 0) return out;
        int kx = sh.kx, ky = sh.ky, kz = sh.kz;
        float okx = sh.o[kx], oky = sh.o[ky], okz = sh.o[kz];
        float e0[4], e1[4], e2[4], akz[4], bkz[4], ckz[4], det[4];
        uint32_t zero_any = 0, neg = 0, pos = 0, det_zero = 0;
        for (int l = 0; l < 4; ++l) {
            akz[l] = v[0][kz][l] - okz;
            bkz[l] = v[1][kz][l] - okz;
            ckz[l] = v[2][kz][l] - okz;
            float ax = (v[0][kx][l] - okx) - sh.sx * akz[l];
            float ay = (v[0][ky][l] - oky) - sh.sy * akz[l];
            float bx = (v[1][kx][l] - okx) - sh.sx * bkz[l];
            float by = (v[1][ky][l] - oky) - sh.sy * bkz[l];
            float cx = (v[2][kx][l] - okx) - sh.sx * ckz[l];
            float cy = (v[2][ky][l] - oky) - sh.sy * ckz[l];
            e0[l] = bx * cy - by * cx;
            e1[l] = cx * ay - cy * ax;
            e2[l] = ax * by - ay * bx;
            if (e0[l] == 0.0f || e1[l] == 0.0f || e2[l] == 0.0f) zero_any |= 1u << l;
            if (e0[l] < 0.0f || e1[l] < 0.0f || e2[l] < 0.0f) neg |= 1u << l;
            if (e0[l] > 0.0f || e1[l] > 0.0f || e2[l] > 0.0f) pos |= 1u << l;
            det[l] = e0[l] + e1[l] + e2[l];
            if (det[l] == 0.0f) det_zero |= 1u << l;
        }
        uint32_t fallback = m & zero_any;
        m &= ~fallback;
        m &= ~(neg & pos);
        m &= ~det_zero;
        out.fallback = fallback;
        if (m == 0) return out;
        float t_scaled[4];
        uint32_t in_range = 0;
        for (int l = 0; l < 4; ++l) {
            t_scaled[l] = e0[l] * (sh.sz * akz[l]) + e1[l] * (sh.sz * bkz[l]) + e2[l] * (sh.sz * ckz[l]);
            float abs_det = std::fabs(det[l]);
            float ts = det[l] < 0.0f ? -t_scaled[l] : t_scaled[l];
            if (ts >= t_min * abs_det && ts <= t_max * abs_det) in_range |= 1u << l;
        }
        m &= in_range;
        if (m == 0) return out;
        for (int l = 0; l < 4; ++l) {
            float inv_det = 1.0f / det[l];
            out.t[l] = t_scaled[l] * inv_det;
            out.u[l] = e1[l] * inv_det;
            out.v[l] = e2[l] * inv_det;
        }
        out.hits = m;
        return out;
    }
};

// triangle.rs:367-429
inline std::optional<AABB> clip_triangle_aabb(V3 v0, V3 v1, V3 v2, int axis, float mn, float mx) {
    V3 poly[8];
    int n = 3;
    poly[0] = v0;
    poly[1] = v1;
    poly[2] = v2;
    for (int pass = 0; pass < 2; ++pass) {
        float bound = pass == 0 ? mn : mx;
        bool keep_ge = pass == 0;
        V3 out[8];
        int m = 0;
        for (int i = 0; i < n; ++i) {
            V3 a = poly[i], b = poly[(i + 1) % n];
            float da = keep_ge ? a[axis] - bound : bound - a[axis];
            float db = keep_ge ? b[axis] - bound : bound - b[axis];
            if (da >= 0.0f) out[m++] = a;
            if ((da > 0.0f) != (db > 0.0f) && da != db) {
                float t = da / (da - db);
                out[m++] = a + (b - a) * t;
            }
        }
        for (int i = 0; i < m; ++i) poly[i] = out[i];
        n = m;
        if (n == 0) return std::nullopt;
    }
    V3 lo = poly[0], hi = poly[0];
    for (int i = 1; i < n; ++i) {
        lo = vmin(lo, poly[i]);
        hi = vmax(hi, poly[i]);
    }
    lo[axis] = std::max(lo[axis], mn);
    hi[axis] = std::min(hi[axis], mx);
    const float PAD = 1e-4f;
    for (int a = 0; a < 3; ++a)
        if (hi[a] - lo[a] < PAD) {
            lo[a] -= PAD;
            hi[a] += PAD;
        }
    return AABB{lo, hi};
}

// ---------------------------------------------------------------------------
// curve.rs
// ---------------------------------------------------------------------------
struct TN {
    float t;
    V3 n;
};
// curve.rs:17-93
inline std::optional<TN> rounded_cone_intersect(const Ray& ray, V3 p0, V3 p1, float r0, float r1, float t_min,
                                                float t_max) {
    r0 = std::max(r0, 1e-6f);
    r1 = std::max(r1, 1e-6f);
    float len = length(ray.dir);
    if (len < 1e-20f) return std::nullopt;
    V3 rd = ray.dir / len, o = ray.origin;
    float n_min = t_min * len, n_max = t_max * len;
    V3 ba = p1 - p0, oa = o - p0, ob = o - p1;
    float rr = r0 - r1;
    float m0 = dot(ba, ba), m1 = dot(ba, oa), m2 = dot(ba, rd), m3 = dot(rd, oa), m5 = dot(oa, oa), m6 = dot(ob, rd),
          m7 = dot(ob, ob);
    bool have = false;
    TN best{0, {0, 0, 0}};
    auto consider = [&](float t, V3 n) {
        if (t >= n_min && t <= n_max && (!have || t < best.t)) {
            best = {t, n};
            have = true;
        }
    };
    float d2 = m0 - rr * rr;
    if (d2 > 0.0f) {
        float k2 = d2 - m2 * m2;
        float k1 = d2 * m3 - m1 * m2 + m2 * rr * r0;
        float k0 = d2 * m5 - m1 * m1 + m1 * rr * r0 * 2.0f - m0 * r0 * r0;
        float h = k1 * k1 - k0 * k2;
        if (h >= 0.0f && std::fabs(k2) > 1e-12f) {
            float sq = std::sqrt(h);
            float ts[2] = {(-k1 - sq) / k2, (-k1 + sq) / k2};
            for (float t : ts) {
                float y = m1 - r0 * rr + t * m2;
                if (y > 0.0f && y < d2) consider(t, normalize(d2 * (oa + t * rd) - ba * y));
            }
        }
    }
    float h0 = m3 * m3 - m5 + r0 * r0;
    if (h0 >= 0.0f) {
        float sq = std::sqrt(h0);
        float ts[2] = {-m3 - sq, -m3 + sq};
        for (float t : ts) consider(t, (oa + t * rd) / r0);
    }
    float h1 = m6 * m6 - m7 + r1 * r1;
    if (h1 >= 0.0f) {
        float sq = std::sqrt(h1);
        float ts[2] = {-m6 - sq, -m6 + sq};
        for (float t : ts) consider(t, (ob + t * rd) / r1);
    }
    if (!have) return std::nullopt;
    return TN{best.t / len, best.n};
}

constexpr uint32_t MAX_RECURSION_DEPTH = 10;  // curve.rs:98
// curve.rs:122-141
inline uint32_t flatness_depth(const V3 cp[4], float max_width) {
    V3 d0 = cp[0] - cp[1] * 2.0f + cp[2];
    V3 d1 = cp[1] - cp[2] * 2.0f + cp[3];
    float l0 = max_element(vmax(vabs(d0), vabs(d1)));
    if (l0 <= 0.0f || max_width <= 0.0f) return 0;
    float eps = max_width * 0.05f;
    float r0 = std::log2(1.41421356237309504880f * 6.0f * l0 / (8.0f * eps)) * 0.5f;
    if (!std::isfinite(r0)) return 0;
    float f = std::max(std::floor(r0), 0.0f);
    uint32_t d = f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f;
    return std::min(d, MAX_RECURSION_DEPTH);
}
// curve.rs:178-217 (subdivide_bezier :104-112, bezier_bounds :146-150 inlined)
inline std::optional<TN> subdivide_and_intersect(const Ray& ray, const V3 cp[4], float u0, float u1, float r0_full,
                                                 float r1_full, uint32_t depth, float t_min, float t_max) {
    auto radius_at = [&](float u) { return r0_full + (r1_full - r0_full) * u; };
    if (depth == 0) return rounded_cone_intersect(ray, cp[0], cp[3], radius_at(u0), radius_at(u1), t_min, t_max);
    V3 p01 = (cp[0] + cp[1]) * 0.5f, p12 = (cp[1] + cp[2]) * 0.5f, p23 = (cp[2] + cp[3]) * 0.5f;
    V3 p012 = (p01 + p12) * 0.5f, p123 = (p12 + p23) * 0.5f, p0123 = (p012 + p123) * 0.5f;
    V3 split[7] = {cp[0], p01, p012, p0123, p123, p23, cp[3]};
    float u_mid = (u0 + u1) * 0.5f;
    std::optional<TN> best;
    for (int h = 0; h < 2; ++h) {
        const V3* sub = split + 3 * h;
        float su0 = h == 0 ? u0 : u_mid, su1 = h == 0 ? u_mid : u1;
        float cur_t_max = best ? best->t : t_max;
        float radius = std::max(radius_at(su0), radius_at(su1));
        AABB bounds{vmin(vmin(vmin(sub[0], sub[1]), sub[2]), sub[3]) - splat(radius),
                    vmax(vmax(vmax(sub[0], sub[1]), sub[2]), sub[3]) + splat(radius)};
        if (!aabb_hit(bounds, ray, t_min, cur_t_max)) continue;
        auto hit = subdivide_and_intersect(ray, sub, su0, su1, r0_full, r1_full, depth - 1, t_min, cur_t_max);
        if (hit) best = hit;
    }
    return best;
}
// curve.rs:166-176
inline std::optional<TN> cubic_curve_intersect(const Ray& ray, const V3 cp[4], float r0, float r1, float t_min,
                                               float t_max) {
    uint32_t depth = flatness_depth(cp, 2.0f * std::max(r0, r1));
    return subdivide_and_intersect(ray, cp, 0.0f, 1.0f, r0, r1, depth, t_min, t_max);
}

// ---------------------------------------------------------------------------
// Traversal counters: semantics of bvh.rs:39-127 ([top-level, inside an
// instance]), kept per query context instead of global atomics, plus the
// per-kind split SURVEY §8(d) uses for the bytes-per-ray figure.
// ---------------------------------------------------------------------------
struct Stats {
    uint64_t queries[2] = {0, 0}, nodes[2] = {0, 0}, leaves[2] = {0, 0}, packets[2] = {0, 0}, prim_tests[2] = {0, 0};
    uint64_t tri_lanes = 0;       // active lanes of every packet tested
    uint64_t sphere_tests = 0, curve_tests = 0, cubic_tests = 0, instance_tests = 0;
    uint64_t stack_high_water = 0;
    void add(const Stats& o) {
        for (int i = 0; i < 2; ++i) {
            queries[i] += o.queries[i];
            nodes[i] += o.nodes[i];
            leaves[i] += o.leaves[i];
            packets[i] += o.packets[i];
            prim_tests[i] += o.prim_tests[i];
        }
        tri_lanes += o.tri_lanes;
        sphere_tests += o.sphere_tests;
        curve_tests += o.curve_tests;
        cubic_tests += o.cubic_tests;
        instance_tests += o.instance_tests;
        stack_high_water = std::max(stack_high_water, o.stack_high_water);
    }
};
struct Ctx {  // per-thread query context (replaces the thread_local DEPTH)
    Stats* stats = nullptr;
    uint32_t depth = 0;
    int slot() const { return depth > 0 ? 1 : 0; }
};

// ---------------------------------------------------------------------------
// prim.rs
// ---------------------------------------------------------------------------
struct PrimHit {  // prim.rs:18-25
    float t;
    V3 outward;
    float u, v;
    uint32_t geom_id, prim_id;
};
class Scene;
enum class Kind : uint8_t { Triangle, Sphere, Curve, CubicCurve, Instance };

struct InstanceData {  // prim.rs:261-280
    std::shared_ptr<const Scene> scene;
    Affine l2w, w2l;
    Mat3 normal_mat;
    bool has_end = false;
    Affine l2w_end;
    AABB bounds;
};
struct PrimNode {  // prim.rs:404-415 (tagged union, flattened)
    Kind kind;
    uint32_t geom_id = 0, prim_id = 0, mask = MASK_ALL;
    V3 p[4];               // triangle v0..v2 | sphere center | curve p0,p1 | cubic cp[4]
    float r0 = 0, r1 = 0;  // sphere radius in r0
    bool has_normals = false;
    V3 nrm[3];
    std::unique_ptr<InstanceData> inst;

    bool is_triangle() const { return kind == Kind::Triangle; }
    AABB bbox() const;
    std::optional<AABB> clipped_aabb(int axis, float mn, float mx) const;
    std::optional<PrimHit> hit(const Ray& ray, float t_min, float t_max, Ctx& cx) const;
    bool hit_any(const Ray& ray, float t_min, float t_max, Ctx& cx) const;
    std::optional<PrimHit> hit_from_barycentric(float t, float u, float v) const;  // prim.rs:76-95
};

// prim.rs:285-294
inline Affine lerp_affine(const Affine& a, const Affine& b, float t) {
    Affine r;
    r.x_axis = lerp(a.x_axis, b.x_axis, t);
    r.y_axis = lerp(a.y_axis, b.y_axis, t);
    r.z_axis = lerp(a.z_axis, b.z_axis, t);
    r.translation = lerp(a.translation, b.translation, t);
    return r;
}
// prim.rs:298-319
inline AABB transformed_aabb(const AABB& local, const Affine& m) {
    V3 mn = splat(INF), mx = splat(-INF);
    for (int i = 0; i < 8; ++i) {
        V3 corner{(i & 1) == 0 ? local.minimum.x : local.maximum.x, (i & 2) == 0 ? local.minimum.y : local.maximum.y,
                  (i & 4) == 0 ? local.minimum.z : local.maximum.z};
        V3 p = transform_point(m, corner);
        mn = vmin(mn, p);
        mx = vmax(mx, p);
    }
    const float PAD = 1e-4f;
    for (int a = 0; a < 3; ++a)
        if (mx[a] - mn[a] < PAD) {
            mn[a] -= PAD;
            mx[a] += PAD;
        }
    return {mn, mx};
}

// ---------------------------------------------------------------------------
// bvh.rs
// ---------------------------------------------------------------------------
constexpr size_t MAX_DEPTH = 60, MIN_LEAF = 2, MIN_LEAF_PACKED = 4, MAX_LEAF = 8, BINS = 12, PARALLEL_THRESHOLD = 4096,
                 SBVH_MAX_DEPTH = 32;  // bvh.rs:142-167
constexpr float SBVH_ALPHA = 1e-5f;
constexpr uint32_t EMPTY_LANE = 0xFFFFFFFFu, VALID_MASK = 0xF, LEAF_SHIFT = 4;

struct BNode {  // bvh.rs:171-178
    AABB bbox;
    uint32_t first_or_right, count;
};
struct WideNode {  // bvh.rs:196-215 — 128 bytes
    float bmin_x[4], bmin_y[4], bmin_z[4], bmax_x[4], bmax_y[4], bmax_z[4];
    uint32_t child[4];
    uint32_t flags;
    uint32_t _pad[3];
    WideNode() {
        for (int l = 0; l < 4; ++l) {
            bmin_x[l] = bmin_y[l] = bmin_z[l] = bmax_x[l] = bmax_y[l] = bmax_z[l] = INF;
            child[l] = EMPTY_LANE;
        }
        flags = 0;
        _pad[0] = _pad[1] = _pad[2] = 0;
    }
    void set_lane_bounds(int lane, const AABB& b) {
        bmin_x[lane] = b.minimum.x;
        bmin_y[lane] = b.minimum.y;
        bmin_z[lane] = b.minimum.z;
        bmax_x[lane] = b.maximum.x;
        bmax_y[lane] = b.maximum.y;
        bmax_z[lane] = b.maximum.z;
        flags |= 1u << lane;
    }
    bool is_leaf(int lane) const { return (flags & (1u << (LEAF_SHIFT + lane))) != 0; }
    void mark_leaf(int lane) { flags |= 1u << (LEAF_SHIFT + lane); }
};
static_assert(sizeof(WideNode) == 128, "bvh.rs:1613-1616");
struct Leaf {  // bvh.rs:220-227
    uint32_t pkt_first, pkt_count, idx_first, idx_count;
};
struct PrimRef {  // bvh.rs:281-290
    AABB bbox;
    uint32_t idx;
    V3 centroid() const { return 0.5f * (bbox.minimum + bbox.maximum); }
};
struct Subtree {
    std::vector<BNode> nodes;
    std::vector<uint32_t> indices;
};

inline float surface_area(const AABB& b) {  // bvh.rs:811-814
    V3 d = b.maximum - b.minimum;
    return 2.0f * (d.x * d.y + d.y * d.z + d.z * d.x);
}
inline AABB union_all(const std::vector<PrimRef>& refs) {  // bvh.rs:816-820
    AABB acc = refs[0].bbox;
    for (size_t i = 1; i < refs.size(); ++i) acc = surrounding_box(acc, refs[i].bbox);
    return acc;
}
inline std::optional<AABB> intersect_aabb(const AABB& a, const AABB& b) {  // bvh.rs:823-831
    V3 lo = vmax(a.minimum, b.minimum), hi = vmin(a.maximum, b.maximum);
    if (lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z) return AABB{lo, hi};
    return std::nullopt;
}
// Rust `f32 as usize`: saturating, NaN -> 0.
inline size_t f32_as_usize(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 18446744073709551616.0f) return std::numeric_limits<size_t>::max();
    return (size_t)f;
}
inline int longest_axis(V3 e) { return (e.x >= e.y && e.x >= e.z) ? 0 : (e.y >= e.z ? 1 : 2); }

struct ObjSplit {  // bvh.rs:876-885
    int axis;
    float cmin, scale;
    size_t split_bin;
    float cost;
    AABB left_bbox, right_bbox;
};
struct SpatSplit {  // bvh.rs:888-892
    int axis;
    float pos, cost;
};

struct OptBox {  // Option<AABB> accumulation helper
    bool some = false;
    AABB b;
    void add(const AABB& x) {
        b = some ? surrounding_box(b, x) : x;
        some = true;
    }
    void merge(const OptBox& o) {
        if (o.some) add_keep_order(o);
    }
    // (Some(x), Some(y)) => surrounding_box(x, y); (x, y) => x.or(y)
    void add_keep_order(const OptBox& o) {
        if (some)
            b = surrounding_box(b, o.b);
        else {
            b = o.b;
            some = true;
        }
    }
};

// bvh.rs:894-969
inline std::optional<ObjSplit> best_object_split(const std::vector<PrimRef>& refs) {
    V3 cmin = refs[0].centroid(), cmax = cmin;
    for (size_t i = 1; i < refs.size(); ++i) {
        V3 c = refs[i].centroid();
        cmin = vmin(cmin, c);
        cmax = vmax(cmax, c);
    }
    V3 extent = cmax - cmin;
    int axis = longest_axis(extent);
    if (extent[axis] <= 1e-6f) return std::nullopt;
    float scale = (float)BINS / extent[axis];
    auto bin_of = [&](const PrimRef& r) {
        return std::min(f32_as_usize((r.centroid()[axis] - cmin[axis]) * scale), BINS - 1);
    };
    size_t bin_counts[BINS] = {};
    OptBox bin_bounds[BINS];
    for (const PrimRef& r : refs) {
        size_t b = bin_of(r);
        bin_counts[b] += 1;
        bin_bounds[b].add(r.bbox);
    }
    std::optional<ObjSplit> best;
    for (size_t split = 0; split + 1 < BINS; ++split) {
        OptBox lb, rb;
        size_t lc = 0, rc = 0;
        for (size_t b = 0; b <= split; ++b) {
            lc += bin_counts[b];
            lb.merge(bin_bounds[b]);
        }
        for (size_t b = split + 1; b < BINS; ++b) {
            rc += bin_counts[b];
            rb.merge(bin_bounds[b]);
        }
        if (lc == 0 || rc == 0) continue;
        float cost = surface_area(lb.b) * (float)lc + surface_area(rb.b) * (float)rc;
        if (!best || cost < best->cost) best = ObjSplit{axis, cmin[axis], scale, split, cost, lb.b, rb.b};
    }
    return best;
}

// bvh.rs:974-1054
inline std::optional<SpatSplit> best_spatial_split(const std::vector<PrimNode>& prims, const std::vector<PrimRef>& refs,
                                                   const AABB& bbox) {
    V3 extent = bbox.maximum - bbox.minimum;
    int axis = longest_axis(extent);
    if (extent[axis] <= 1e-6f) return std::nullopt;
    float lo = bbox.minimum[axis];
    float width = extent[axis] / (float)BINS;
    auto bin_of = [&](float x) { return std::min(f32_as_usize((x - lo) / width), BINS - 1); };
    size_t entry[BINS] = {}, exit_[BINS] = {};
    OptBox bounds[BINS];
    for (const PrimRef& r : refs) {
        size_t b0 = bin_of(r.bbox.minimum[axis]), b1 = bin_of(r.bbox.maximum[axis]);
        entry[b0] += 1;
        exit_[b1] += 1;
        if (b0 == b1) {
            bounds[b0].add(r.bbox);
            continue;
        }
        for (size_t b = b0; b <= b1; ++b) {  // empty when b0 > b1, like Rust's b0..=b1
            float bin_lo = lo + (float)b * width, bin_hi = lo + (float)(b + 1) * width;
            auto c = prims[r.idx].clipped_aabb(axis, bin_lo, bin_hi);
            if (!c) continue;
            auto ci = intersect_aabb(*c, r.bbox);
            if (ci) bounds[b].add(*ci);
        }
    }
    std::optional<SpatSplit> best;
    for (size_t split = 0; split + 1 < BINS; ++split) {
        OptBox lb, rb;
        size_t lc = 0, rc = 0;
        for (size_t b = 0; b <= split; ++b) {
            lc += entry[b];
            lb.merge(bounds[b]);
        }
        for (size_t b = split + 1; b < BINS; ++b) {
            rc += exit_[b];
            rb.merge(bounds[b]);
        }
        if (lc == 0 || rc == 0) continue;
        // `.expect("lc > 0")`: a Rust panic if every contributing bin ended up
        // without bounds. Unreachable on finite input; treat as "no candidate".
        if (!lb.some || !rb.some) continue;
        float cost = surface_area(lb.b) * (float)lc + surface_area(rb.b) * (float)rc;
        if (!best || cost < best->cost) best = SpatSplit{axis, lo + (float)(split + 1) * width, cost};
    }
    return best;
}

inline Subtree make_leaf(const AABB& bbox, const std::vector<PrimRef>& refs) {  // bvh.rs:833-842
    Subtree s;
    s.nodes.push_back(BNode{bbox, 0, (uint32_t)refs.size()});
    s.indices.reserve(refs.size());
    for (const PrimRef& r : refs) s.indices.push_back(r.idx);
    return s;
}
inline Subtree merge(const AABB& bbox, Subtree left, Subtree right) {  // bvh.rs:846-872
    Subtree out;
    out.nodes.reserve(1 + left.nodes.size() + right.nodes.size());
    uint32_t right_node_offset = 1 + (uint32_t)left.nodes.size();
    out.nodes.push_back(BNode{bbox, right_node_offset, 0});
    for (BNode n : left.nodes) {
        if (n.count == 0) n.first_or_right += 1;
        out.nodes.push_back(n);
    }
    uint32_t leaf_offset = (uint32_t)left.indices.size();
    for (BNode n : right.nodes) {
        if (n.count == 0)
            n.first_or_right += right_node_offset;
        else
            n.first_or_right += leaf_offset;
        out.nodes.push_back(n);
    }
    out.indices = std::move(left.indices);
    out.indices.insert(out.indices.end(), right.indices.begin(), right.indices.end());
    return out;
}
inline size_t min_leaf_for(const std::vector<PrimNode>& prims, const std::vector<PrimRef>& refs) {  // bvh.rs:1201-1210
    for (const PrimRef& r : refs)
        if (!prims[r.idx].is_triangle()) return MIN_LEAF;
    return MIN_LEAF_PACKED;
}
// bvh.rs:1224-1237
inline void partition_by_bin(std::vector<PrimRef>&& refs, const ObjSplit& o, std::vector<PrimRef>& left,
                             std::vector<PrimRef>& right) {
    auto bin_of = [&](const PrimRef& r) {
        return std::min(f32_as_usize((r.centroid()[o.axis] - o.cmin) * o.scale), BINS - 1);
    };
    size_t n_left = 0;
    for (const PrimRef& r : refs) n_left += bin_of(r) <= o.split_bin;
    left.reserve(n_left);
    right.reserve(refs.size() - n_left);
    for (const PrimRef& r : refs) (bin_of(r) <= o.split_bin ? left : right).push_back(r);
    std::vector<PrimRef>().swap(refs);
}

struct BuildPool {  // stands in for rayon::join: bounded number of extra threads
    std::atomic<int> free_threads;
    explicit BuildPool(int n) : free_threads(n) {}
};

Subtree build_subtree(const std::vector<PrimNode>& prims, std::vector<PrimRef> refs, size_t depth, float root_area,
                      BuildPool& pool);

// bvh.rs:1173-1196
inline Subtree object_partition_or_leaf(const std::vector<PrimNode>& prims, std::vector<PrimRef> refs, const AABB& bbox,
                                        const std::optional<ObjSplit>& object, size_t depth, float root_area,
                                        BuildPool& pool) {
    size_t count = refs.size();
    if (object && count > min_leaf_for(prims, refs)) {
        std::vector<PrimRef> l, r;
        partition_by_bin(std::move(refs), *object, l, r);
        if (l.empty() || r.empty()) {
            l.insert(l.end(), r.begin(), r.end());
            return make_leaf(bbox, l);
        }
        Subtree left = build_subtree(prims, std::move(l), depth + 1, root_area, pool);
        Subtree right = build_subtree(prims, std::move(r), depth + 1, root_area, pool);
        return merge(bbox, std::move(left), std::move(right));
    }
    return make_leaf(bbox, refs);
}

// bvh.rs:1059-1169
inline Subtree build_subtree(const std::vector<PrimNode>& prims, std::vector<PrimRef> refs, size_t depth,
                             float root_area, BuildPool& pool) {
    AABB bbox = union_all(refs);
    size_t count = refs.size();
    size_t min_leaf = min_leaf_for(prims, refs);
    if (count <= min_leaf || depth >= MAX_DEPTH) return make_leaf(bbox, refs);

    std::optional<ObjSplit> object = best_object_split(refs);
    std::optional<SpatSplit> spatial;
    if (object && depth < SBVH_MAX_DEPTH) {
        auto ov = intersect_aabb(object->left_bbox, object->right_bbox);
        float overlap = ov ? surface_area(*ov) : 0.0f;
        if (overlap / root_area > SBVH_ALPHA) {
            spatial = best_spatial_split(prims, refs, bbox);
            if (spatial && !(spatial->cost < object->cost)) spatial.reset();
        }
    }

    std::vector<PrimRef> left_refs, right_refs;
    if (spatial) {
        const SpatSplit& s = *spatial;
        for (const PrimRef& r : refs) {
            if (r.bbox.maximum[s.axis] <= s.pos)
                left_refs.push_back(r);
            else if (r.bbox.minimum[s.axis] >= s.pos)
                right_refs.push_back(r);
            else {
                const PrimNode& prim = prims[r.idx];
                if (auto c = prim.clipped_aabb(s.axis, -INF, s.pos))
                    if (auto ci = intersect_aabb(*c, r.bbox)) left_refs.push_back(PrimRef{*ci, r.idx});
                if (auto c = prim.clipped_aabb(s.axis, s.pos, INF))
                    if (auto ci = intersect_aabb(*c, r.bbox)) right_refs.push_back(PrimRef{*ci, r.idx});
            }
        }
        std::vector<PrimRef>().swap(refs);
        if (left_refs.empty() || right_refs.empty()) {
            left_refs.insert(left_refs.end(), right_refs.begin(), right_refs.end());
            return object_partition_or_leaf(prims, std::move(left_refs), bbox, object, depth, root_area, pool);
        }
    } else if (object) {
        if (count <= MAX_LEAF && object->cost >= surface_area(bbox) * (float)count) return make_leaf(bbox, refs);
        partition_by_bin(std::move(refs), *object, left_refs, right_refs);
    } else {
        size_t mid = count / 2;
        right_refs.assign(refs.begin() + mid, refs.end());
        refs.resize(mid);
        left_refs = std::move(refs);
    }

    bool parallel = std::max(left_refs.size(), right_refs.size()) > PARALLEL_THRESHOLD;
    if (parallel && pool.free_threads.fetch_sub(1) > 0) {
        auto fut = std::async(std::launch::async, [&prims, &pool, depth, root_area, rr = std::move(right_refs)]() mutable {
            Subtree s = build_subtree(prims, std::move(rr), depth + 1, root_area, pool);
            pool.free_threads.fetch_add(1);
            return s;
        });
        Subtree l = build_subtree(prims, std::move(left_refs), depth + 1, root_area, pool);
        Subtree r = fut.get();
        return merge(bbox, std::move(l), std::move(r));
    }
    if (parallel) pool.free_threads.fetch_add(1);
    Subtree l = build_subtree(prims, std::move(left_refs), depth + 1, root_area, pool);
    Subtree r = build_subtree(prims, std::move(right_refs), depth + 1, root_area, pool);
    return merge(bbox, std::move(l), std::move(r));
}

constexpr size_t STACK_INLINE = 32;  // bvh.rs:693 (only the high-water mark is observable here)

class Bvh {  // bvh.rs:263-276
public:
    std::vector<WideNode> wide;
    std::vector<Leaf> leaves;
    std::vector<Tri4> packets;
    std::vector<uint32_t> indices;
    std::vector<PrimNode> prims;
    std::optional<AABB> root_bbox;
    size_t binary_nodes = 0;

    Bvh() = default;
    Bvh(std::vector<PrimNode> p, int threads);  // bvh.rs:300-327

    std::optional<PrimHit> hit(const Ray& ray, float t_min, float t_max, Ctx& cx) const;  // bvh.rs:441-509
    bool hit_any(const Ray& ray, float t_min, float t_max, Ctx& cx) const;                // bvh.rs:585-611
    size_t leaf_ref_count() const {                                                       // bvh.rs:431-434
        size_t packed = 0;
        for (const Tri4& p : packets) packed += __builtin_popcount(p.active);
        return packed + indices.size();
    }

private:
    // bvh.rs:1256-1290
    uint32_t push_leaf(const uint32_t* range, size_t n) {
        uint32_t pkt_first = (uint32_t)packets.size();
        TriIn batch[4];
        size_t nb = 0;
        uint32_t idx_first = (uint32_t)indices.size(), idx_count = 0;
        for (size_t i = 0; i < n; ++i) {
            uint32_t pi = range[i];
            const PrimNode& p = prims[pi];
            if (p.is_triangle()) {
                batch[nb++] = TriIn{p.p[0], p.p[1], p.p[2], pi, p.mask};
                if (nb == 4) {
                    packets.emplace_back(batch, 4);
                    nb = 0;
                }
            } else {
                indices.push_back(pi);
                idx_count += 1;
            }
        }
        if (nb) packets.emplace_back(batch, nb);
        leaves.push_back(Leaf{pkt_first, (uint32_t)packets.size() - pkt_first, idx_first, idx_count});
        return (uint32_t)leaves.size() - 1;
    }
    // bvh.rs:1328-1375
    uint32_t collapse_node(const std::vector<BNode>& binary, const std::vector<uint32_t>& bidx, uint32_t b_idx) {
        size_t slot = wide.size();
        wide.emplace_back();
        uint32_t kids[4] = {b_idx + 1, binary[b_idx].first_or_right, 0, 0};
        int n_kids = 2;
        while (n_kids < 4) {
            int best = -1;
            float best_a = 0;
            for (int i = 0; i < n_kids; ++i) {
                const BNode& k = binary[kids[i]];
                if (k.count == 0) {
                    float a = surface_area(k.bbox);
                    if (best < 0 || a > best_a) {
                        best = i;
                        best_a = a;
                    }
                }
            }
            if (best < 0) break;
            uint32_t k = kids[best];
            kids[best] = k + 1;
            kids[n_kids] = binary[k].first_or_right;
            n_kids += 1;
        }
        for (int lane = 0; lane < n_kids; ++lane) {
            const BNode& k = binary[kids[lane]];
            wide[slot].set_lane_bounds(lane, k.bbox);
            if (k.count > 0) {
                uint32_t li = push_leaf(bidx.data() + k.first_or_right, k.count);
                wide[slot].child[lane] = li;
                wide[slot].mark_leaf(lane);
            } else {
                uint32_t ci = collapse_node(binary, bidx, kids[lane]);
                wide[slot].child[lane] = ci;
            }
        }
        return (uint32_t)slot;
    }
    std::optional<PrimHit> intersect_leaf(uint32_t leaf_idx, const Ray& ray, const RayShear* shear, float t_min,
                                          float t_max, Ctx& cx) const;  // bvh.rs:514-572
    bool occlude_leaf(uint32_t leaf_idx, const Ray& ray, const RayShear* shear, float t_min, float t_max,
                      Ctx& cx) const;  // bvh.rs:617-653
};

// scene.rs:345-479 (committed scene) — geometry expansion lives in the builder below.
class Scene {
public:
    Bvh bvh;
    uint32_t n_geoms = 0;
    bool has_motion = false;
    std::optional<PrimHit> intersect_outward(const Ray& ray, float t_min, float t_max, Ctx& cx) const {
        return bvh.hit(ray, t_min, t_max, cx);
    }
    bool occluded(const Ray& ray, float t_min, float t_max, Ctx& cx) const { return bvh.hit_any(ray, t_min, t_max, cx); }
};

// ---- PrimNode methods -------------------------------------------------------
inline AABB PrimNode::bbox() const {
    switch (kind) {
        case Kind::Triangle: return triangle_aabb(p[0], p[1], p[2]);                       // prim.rs:112-114
        case Kind::Sphere: return {p[0] - splat(r0), p[0] + splat(r0)};                     // prim.rs:163-168
        case Kind::Curve:                                                                  // prim.rs:202-206
            return surrounding_box({p[0] - splat(r0), p[0] + splat(r0)}, {p[1] - splat(r1), p[1] + splat(r1)});
        case Kind::CubicCurve: {                                                           // prim.rs:246-253
            float radius = std::max(r0, r1);
            return {vmin(vmin(vmin(p[0], p[1]), p[2]), p[3]) - splat(radius),
                    vmax(vmax(vmax(p[0], p[1]), p[2]), p[3]) + splat(radius)};
        }
        case Kind::Instance: return inst->bounds;  // prim.rs:380-382
    }
    return {};
}
inline std::optional<AABB> PrimNode::clipped_aabb(int axis, float mn, float mx) const {
    if (kind == Kind::Triangle) return clip_triangle_aabb(p[0], p[1], p[2], axis, mn, mx);  // prim.rs:116-118
    AABB b = bbox();                                                                         // prim.rs:39-48
    if (b.minimum[axis] > mx || b.maximum[axis] < mn) return std::nullopt;
    b.minimum[axis] = std::max(b.minimum[axis], mn);
    b.maximum[axis] = std::min(b.maximum[axis], mx);
    return b;
}
inline std::optional<PrimHit> PrimNode::hit_from_barycentric(float t, float u, float v) const {
    V3 outward;
    if (has_normals) {
        outward = normalize(nrm[0] * (1.0f - u - v) + nrm[1] * u + nrm[2] * v);
    } else {
        V3 n = cross(p[1] - p[0], p[2] - p[0]);
        if (is_zero(n)) return std::nullopt;
        outward = normalize(n);
    }
    return PrimHit{t, outward, u, v, geom_id, prim_id};
}
inline std::optional<PrimHit> PrimNode::hit(const Ray& ray, float t_min, float t_max, Ctx& cx) const {
    if ((ray.mask & mask) == 0) return std::nullopt;  // prim.rs:51-54
    switch (kind) {
        case Kind::Triangle: {  // prim.rs:99-105
            auto r = triangle_intersect(ray, p[0], p[1], p[2], t_min, t_max);
            if (!r) return std::nullopt;
            return hit_from_barycentric(r->t, r->u, r->v);
        }
        case Kind::Sphere: {  // prim.rs:133-161
            if (cx.stats) cx.stats->sphere_tests++;
            V3 oc = ray.origin - p[0];
            float a = length_squared(ray.dir);
            float half_b = dot(oc, ray.dir);
            float c = length_squared(oc) - r0 * r0;
            float disc = half_b * half_b - a * c;
            if (disc < 0.0f) return std::nullopt;
            float sqrt_d = std::sqrt(disc);
            float root = (-half_b - sqrt_d) / a;
            if (root <= t_min || root >= t_max) {
                root = (-half_b + sqrt_d) / a;
                if (root <= t_min || root >= t_max) return std::nullopt;
            }
            return PrimHit{root, (ray.at(root) - p[0]) / r0, 0.0f, 0.0f, geom_id, 0};
        }
        case Kind::Curve: {  // prim.rs:186-200
            if (cx.stats) cx.stats->curve_tests++;
            auto r = rounded_cone_intersect(ray, p[0], p[1], r0, r1, t_min, t_max);
            if (!r) return std::nullopt;
            return PrimHit{r->t, r->n, 0.0f, 0.0f, geom_id, prim_id};
        }
        case Kind::CubicCurve: {  // prim.rs:230-244
            if (cx.stats) cx.stats->cubic_tests++;
            auto r = cubic_curve_intersect(ray, p, r0, r1, t_min, t_max);
            if (!r) return std::nullopt;
            return PrimHit{r->t, r->n, 0.0f, 0.0f, geom_id, prim_id};
        }
        case Kind::Instance: {  // prim.rs:345-365
            if (cx.stats) cx.stats->instance_tests++;
            const InstanceData& I = *inst;
            Affine w2l = I.w2l;
            Mat3 nm = I.normal_mat;
            if (I.has_end && ray.time > 0.0f) {  // prim.rs:323-331
                w2l = inverse(lerp_affine(I.l2w, I.l2w_end, ray.time));
                nm = transpose3(w2l);
            }
            Ray local{transform_point(w2l, ray.origin), transform_vector(w2l, ray.dir), ray.time, ray.mask};
            cx.depth++;
            auto inner = I.scene->intersect_outward(local, t_min, t_max, cx);
            cx.depth--;
            if (!inner) return std::nullopt;
            inner->outward = normalize(mul(nm, inner->outward));
            inner->geom_id = geom_id;
            return inner;
        }
    }
    return std::nullopt;
}
inline bool PrimNode::hit_any(const Ray& ray, float t_min, float t_max, Ctx& cx) const {
    switch (kind) {
        case Kind::Triangle:  // prim.rs:107-110
            return (ray.mask & mask) != 0 && triangle_intersect(ray, p[0], p[1], p[2], t_min, t_max).has_value();
        case Kind::Instance: {  // prim.rs:367-378
            if ((ray.mask & mask) == 0) return false;
            if (cx.stats) cx.stats->instance_tests++;
            const InstanceData& I = *inst;
            Affine w2l = I.w2l;
            if (I.has_end && ray.time > 0.0f) w2l = inverse(lerp_affine(I.l2w, I.l2w_end, ray.time));
            Ray local{transform_point(w2l, ray.origin), transform_vector(w2l, ray.dir), ray.time, ray.mask};
            cx.depth++;
            bool occ = I.scene->occluded(local, t_min, t_max, cx);
            cx.depth--;
            return occ;
        }
        default: return hit(ray, t_min, t_max, cx).has_value();  // prim.rs:31-33
    }
}

// ---- Bvh methods -----------------------------------------------------------
inline Bvh::Bvh(std::vector<PrimNode> p, int threads) : prims(std::move(p)) {
    if (prims.empty()) return;
    std::vector<PrimRef> refs(prims.size());
    for (size_t i = 0; i < prims.size(); ++i) refs[i] = PrimRef{prims[i].bbox(), (uint32_t)i};
    AABB rb = union_all(refs);
    BuildPool pool(std::max(0, threads - 1));
    Subtree st = build_subtree(prims, std::move(refs), 0, surface_area(rb), pool);
    binary_nodes = st.nodes.size();
    // bvh.rs:1299-1320
    wide.reserve(st.nodes.size() / 2 + 1);
    if (st.nodes[0].count > 0) {
        wide.emplace_back();
        wide[0].set_lane_bounds(0, st.nodes[0].bbox);
        wide[0].child[0] = push_leaf(st.indices.data() + st.nodes[0].first_or_right, st.nodes[0].count);
        wide[0].mark_leaf(0);
    } else {
        collapse_node(st.nodes, st.indices, 0);
    }
    root_bbox = rb;
}

// bvh.rs:662-668
inline float safe_inv(float d) {
    const float TINY = 1e-20f, HUGE_ = 1e20f;
    return std::fabs(d) < TINY ? std::copysign(HUGE_, d) : 1.0f / d;
}

struct RaySlab {  // bvh.rs:761-809
    V3 o, inv;
    float t_min;
    RaySlab(const Ray& r, float tm) : o(r.origin), inv{safe_inv(r.dir.x), safe_inv(r.dir.y), safe_inv(r.dir.z)}, t_min(tm) {}
    // SSE minps/maxps lane semantics are irrelevant here: safe_inv keeps NaNs out.
    uint32_t slab4(const WideNode& n, float t_max, float tnear[4]) const {
        uint32_t mask = 0;
        for (int l = 0; l < 4; ++l) {
            float t0x = (n.bmin_x[l] - o.x) * inv.x, t1x = (n.bmax_x[l] - o.x) * inv.x;
            float t0y = (n.bmin_y[l] - o.y) * inv.y, t1y = (n.bmax_y[l] - o.y) * inv.y;
            float t0z = (n.bmin_z[l] - o.z) * inv.z, t1z = (n.bmax_z[l] - o.z) * inv.z;
            float tn = std::max(std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::min(t0z, t1z)), t_min);
            float tf = std::min(std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::max(t0z, t1z)), t_max);
            tnear[l] = tn;
            if (tn <= tf) mask |= 1u << l;
        }
        return mask;
    }
};

inline std::optional<PrimHit> Bvh::hit(const Ray& ray, float t_min, float t_max, Ctx& cx) const {
    if (wide.empty()) return std::nullopt;
    Stats* st = cx.stats;
    int slot = cx.slot();
    if (st) st->queries[slot]++;
    float closest = t_max;
    std::optional<PrimHit> best;
    RaySlab rs(ray, t_min);
    std::optional<RayShear> shear;
    if (!packets.empty()) shear.emplace(ray);  // bvh.rs:422-425
    std::vector<uint32_t> spill;
    uint32_t inline_stack[STACK_INLINE];
    size_t sp = 1;
    inline_stack[0] = 0;
    auto push = [&](uint32_t n) {
        if (sp < STACK_INLINE)
            inline_stack[sp] = n;
        else
            spill.push_back(n);
        sp++;
        if (st) st->stack_high_water = std::max<uint64_t>(st->stack_high_water, sp);
    };
    while (sp > 0) {
        sp--;
        uint32_t node_idx;
        if (sp < STACK_INLINE)
            node_idx = inline_stack[sp];
        else {
            node_idx = spill.back();
            spill.pop_back();
        }
        if (st) st->nodes[slot]++;
        const WideNode& node = wide[node_idx];
        float tn[4];
        uint32_t mask = rs.slab4(node, closest, tn) & node.flags & VALID_MASK;
        if (mask == 0) continue;
        float ord_t[4];
        int ord_l[4], n_hit = 0;
        while (mask) {
            int l = __builtin_ctz(mask);
            mask &= mask - 1;
            float t = tn[l];
            int i = n_hit;
            while (i > 0 && ord_t[i - 1] > t) {
                ord_t[i] = ord_t[i - 1];
                ord_l[i] = ord_l[i - 1];
                i--;
            }
            ord_t[i] = t;
            ord_l[i] = l;
            n_hit++;
        }
        for (int i = 0; i < n_hit; ++i) {
            int l = ord_l[i];
            if (node.is_leaf(l)) {
                auto h = intersect_leaf(node.child[l], ray, shear ? &*shear : nullptr, t_min, closest, cx);
                if (h) {
                    closest = h->t;
                    best = h;
                }
            }
        }
        for (int i = n_hit - 1; i >= 0; --i) {
            int l = ord_l[i];
            if (!node.is_leaf(l)) push(node.child[l]);
        }
    }
    return best;
}

inline std::optional<PrimHit> Bvh::intersect_leaf(uint32_t leaf_idx, const Ray& ray, const RayShear* shear, float t_min,
                                                  float t_max, Ctx& cx) const {
    const Leaf& leaf = leaves[leaf_idx];
    Stats* st = cx.stats;
    int slot = cx.slot();
    if (st) {
        st->leaves[slot]++;
        st->packets[slot] += leaf.pkt_count;
        st->prim_tests[slot] += leaf.idx_count;
    }
    float closest = t_max;
    std::optional<PrimHit> best;
    for (uint32_t pk = leaf.pkt_first; pk < leaf.pkt_first + leaf.pkt_count; ++pk) {
        const Tri4& packet = packets[pk];
        if (st) st->tri_lanes += __builtin_popcount(packet.active);
        Hit4 out = packet.intersect(*shear, ray.mask, t_min, closest);
        uint32_t hits = out.hits;
        while (hits) {
            int lane = __builtin_ctz(hits);
            hits &= hits - 1;
            if (out.t[lane] > closest) continue;  // strict: exact tie -> later primitive wins
            auto h = prims[packet.prim[lane]].hit_from_barycentric(out.t[lane], out.u[lane], out.v[lane]);
            if (h) {
                closest = h->t;
                best = h;
            }
        }
        uint32_t fb = out.fallback;
        while (fb) {
            int lane = __builtin_ctz(fb);
            fb &= fb - 1;
            auto h = prims[packet.prim[lane]].hit(ray, t_min, closest, cx);
            if (h) {
                closest = h->t;
                best = h;
            }
        }
    }
    for (uint32_t i = leaf.idx_first; i < leaf.idx_first + leaf.idx_count; ++i) {
        auto h = prims[indices[i]].hit(ray, t_min, closest, cx);
        if (h) {
            closest = h->t;
            best = h;
        }
    }
    return best;
}

inline bool Bvh::hit_any(const Ray& ray, float t_min, float t_max, Ctx& cx) const {
    if (wide.empty()) return false;
    Stats* st = cx.stats;
    int slot = cx.slot();
    // The reference leaves hit_any uninstrumented; counting it here (same
    // meaning as in `hit`) is what gives occlusion rays a bytes-per-ray figure.
    if (st) st->queries[slot]++;
    RaySlab rs(ray, t_min);
    std::optional<RayShear> shear;
    if (!packets.empty()) shear.emplace(ray);
    std::vector<uint32_t> stack;
    stack.reserve(STACK_INLINE);
    stack.push_back(0);
    while (!stack.empty()) {
        uint32_t node_idx = stack.back();
        stack.pop_back();
        if (st) st->nodes[slot]++;
        const WideNode& node = wide[node_idx];
        float tn[4];
        uint32_t mask = rs.slab4(node, t_max, tn) & node.flags & VALID_MASK;
        while (mask) {
            int l = __builtin_ctz(mask);
            mask &= mask - 1;
            if (node.is_leaf(l)) {
                if (occlude_leaf(node.child[l], ray, shear ? &*shear : nullptr, t_min, t_max, cx)) return true;
            } else {
                stack.push_back(node.child[l]);
                if (st) st->stack_high_water = std::max<uint64_t>(st->stack_high_water, stack.size());
            }
        }
    }
    return false;
}

inline bool Bvh::occlude_leaf(uint32_t leaf_idx, const Ray& ray, const RayShear* shear, float t_min, float t_max,
                              Ctx& cx) const {
    const Leaf& leaf = leaves[leaf_idx];
    Stats* st = cx.stats;
    int slot = cx.slot();
    if (st) st->leaves[slot]++;
    for (uint32_t pk = leaf.pkt_first; pk < leaf.pkt_first + leaf.pkt_count; ++pk) {
        const Tri4& packet = packets[pk];
        if (st) {
            st->packets[slot]++;
            st->tri_lanes += __builtin_popcount(packet.active);
        }
        Hit4 out = packet.intersect(*shear, ray.mask, t_min, t_max);
        if (out.hits) return true;
        uint32_t fb = out.fallback;
        while (fb) {
            int lane = __builtin_ctz(fb);
            fb &= fb - 1;
            if (prims[packet.prim[lane]].hit_any(ray, t_min, t_max, cx)) return true;
        }
    }
    for (uint32_t i = leaf.idx_first; i < leaf.idx_first + leaf.idx_count; ++i) {
        if (st) st->prim_tests[slot]++;
        if (prims[indices[i]].hit_any(ray, t_min, t_max, cx)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------
// scene.rs: Geometry / SceneBuilder / Scene::intersect
// ---------------------------------------------------------------------------
struct CurveSegment {  // scene.rs:18-23
    V3 p0, p1;
    float r0, r1;
};
struct CubicCurveSegment {  // scene.rs:76-80
    V3 cp[4];
    float r0, r1;
};
struct Geometry {  // scene.rs:87-122
    enum class Type { TriangleMesh, Sphere, RoundCurves, CubicCurves, Instance } type = Type::TriangleMesh;
    std::vector<V3> vertices, normals;
    bool has_normals = false;
    std::vector<uint32_t> indices;  // 3 per triangle
    V3 center;
    float radius = 0;
    std::vector<CurveSegment> segments;
    std::vector<CubicCurveSegment> cubic;
    std::shared_ptr<const Scene> scene;
    Affine transform, transform_end;
    bool has_end = false;
};
struct RayHit {  // scene.rs:134-142
    float t;
    V3 normal;
    bool front_face;
    float u, v;
    uint32_t geom_id, prim_id;
};

class SceneBuilder {  // scene.rs:147-342
public:
    std::vector<std::pair<Geometry, uint32_t>> geoms;
    uint32_t attach_masked(Geometry g, uint32_t mask) {
        geoms.emplace_back(std::move(g), mask);
        return (uint32_t)geoms.size() - 1;
    }
    bool set_geometry(uint32_t id, Geometry g) {  // reference panics on a bad id (scene.rs:196-200)
        if (id >= geoms.size()) return false;
        geoms[id].first = std::move(g);
        return true;
    }
    std::shared_ptr<Scene> commit(int threads) {  // scene.rs:226-341
        auto out = std::make_shared<Scene>();
        out->n_geoms = (uint32_t)geoms.size();
        std::vector<PrimNode> prims;
        bool has_motion = false;
        for (size_t gi = 0; gi < geoms.size(); ++gi) {
            Geometry& g = geoms[gi].first;
            uint32_t mask = geoms[gi].second, geom_id = (uint32_t)gi;
            switch (g.type) {
                case Geometry::Type::TriangleMesh: {
                    size_t ntri = g.indices.size() / 3, nv = g.vertices.size();
                    for (size_t t = 0; t < ntri; ++t) {
                        size_t i0 = g.indices[3 * t], i1 = g.indices[3 * t + 1], i2 = g.indices[3 * t + 2];
                        if (i0 >= nv || i1 >= nv || i2 >= nv) continue;
                        PrimNode p;
                        p.kind = Kind::Triangle;
                        p.p[0] = g.vertices[i0];
                        p.p[1] = g.vertices[i1];
                        p.p[2] = g.vertices[i2];
                        if (g.has_normals && i0 < g.normals.size() && i1 < g.normals.size() && i2 < g.normals.size()) {
                            p.has_normals = true;
                            p.nrm[0] = g.normals[i0];
                            p.nrm[1] = g.normals[i1];
                            p.nrm[2] = g.normals[i2];
                        }
                        p.geom_id = geom_id;
                        p.prim_id = (uint32_t)t;
                        p.mask = mask;
                        prims.push_back(std::move(p));
                    }
                    break;
                }
                case Geometry::Type::Sphere: {
                    PrimNode p;
                    p.kind = Kind::Sphere;
                    p.p[0] = g.center;
                    p.r0 = g.radius;
                    p.geom_id = geom_id;
                    p.mask = mask;
                    prims.push_back(std::move(p));
                    break;
                }
                case Geometry::Type::RoundCurves:
                    for (size_t s = 0; s < g.segments.size(); ++s) {
                        PrimNode p;
                        p.kind = Kind::Curve;
                        p.p[0] = g.segments[s].p0;
                        p.p[1] = g.segments[s].p1;
                        p.r0 = g.segments[s].r0;
                        p.r1 = g.segments[s].r1;
                        p.geom_id = geom_id;
                        p.prim_id = (uint32_t)s;
                        p.mask = mask;
                        prims.push_back(std::move(p));
                    }
                    break;
                case Geometry::Type::CubicCurves:
                    for (size_t s = 0; s < g.cubic.size(); ++s) {
                        PrimNode p;
                        p.kind = Kind::CubicCurve;
                        for (int k = 0; k < 4; ++k) p.p[k] = g.cubic[s].cp[k];
                        p.r0 = g.cubic[s].r0;
                        p.r1 = g.cubic[s].r1;
                        p.geom_id = geom_id;
                        p.prim_id = (uint32_t)s;
                        p.mask = mask;
                        prims.push_back(std::move(p));
                    }
                    break;
                case Geometry::Type::Instance: {
                    if (!g.scene || !g.scene->bvh.root_bbox) break;  // empty instanced scene
                    AABB inner = *g.scene->bvh.root_bbox;
                    PrimNode p;
                    p.kind = Kind::Instance;
                    p.inst = std::make_unique<InstanceData>();
                    InstanceData& I = *p.inst;
                    I.scene = g.scene;
                    I.l2w = g.transform;
                    I.w2l = inverse(g.transform);
                    I.normal_mat = transpose3(I.w2l);
                    I.has_end = g.has_end;
                    I.l2w_end = g.transform_end;
                    I.bounds = g.has_end ? surrounding_box(transformed_aabb(inner, g.transform),
                                                           transformed_aabb(inner, g.transform_end))
                                         : transformed_aabb(inner, g.transform);
                    has_motion |= g.has_end || g.scene->has_motion;
                    p.geom_id = geom_id;
                    p.mask = mask;
                    prims.push_back(std::move(p));
                    break;
                }
            }
        }
        geoms.clear();
        out->has_motion = has_motion;
        out->bvh = Bvh(std::move(prims), threads);
        return out;
    }
};

// scene.rs:354-366
inline std::optional<RayHit> scene_intersect(const Scene& s, const Ray& ray, float t_min, float t_max, Ctx& cx) {
    auto h = s.bvh.hit(ray, t_min, t_max, cx);
    if (!h) return std::nullopt;
    bool front_face = dot(ray.dir, h->outward) < 0.0f;
    return RayHit{h->t, front_face ? h->outward : -h->outward, front_face, h->u, h->v, h->geom_id, h->prim_id};
}

}  // namespace orc
