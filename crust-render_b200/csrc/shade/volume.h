// Free-standing participating volumes: oriented boxes of medium outside the surface BVH, sampled by weighted
// delta tracking (interactions) and ratio tracking (shadow transmittance, analytic when every crossed region is
// homogeneous). Restates crates/crust-core/src/volume.rs (:22-166 DensityField, :168-283 VolumeRegion, :285-329
// PhaseMix, :331-540 Volumes) and medium.rs:145-149 (hg_phase). Host + device: a grid's voxels live in host memory
// for the oracle and in HBM for the kernels.
//
// The reference walks with `openqmc::pcg::Rng` streams taken off a sampler domain (`v.new_domain(K_VOLUME).rng()`,
// tracer.rs:1033,1139,1180); openqmc-rs is not vendored, so — like every sample value (base.h) — the stream here is
// this repo's own (PARITY UNPINNED for values): a 32-bit PCG keyed by the domain.
#pragma once
#include "material.h"

namespace shade {

constexpr int MAX_VOLUME_REGIONS = 16;  // device-side span list; crt_world_set_volumes rejects more (the reference has no cap)
enum : uint32_t { FIELD_HOMOGENEOUS = 0, FIELD_NOISE = 1, FIELD_GRID = 2 };

struct Rng {
    uint32_t state;
};
CRT_HD Rng rng_new(uint32_t seed) { return Rng{pcg_hash(seed ^ 0x5851F42Du)}; }
CRT_HD Rng sampler_rng(PathSampler s) { return Rng{pcg_hash(hash_combine(s.pattern, s.index * 0x9E3779B1u + 0x2C9277B5u))}; }
CRT_HD float rng_next_f32(Rng& r) {  // [0, 1)
    r.state = r.state * 747796405u + 2891336453u;
    uint32_t w = ((r.state >> ((r.state >> 28u) + 4u)) ^ r.state) * 277803737u;
    return u32_to_unit_f32((w >> 22u) ^ w);
}

CRT_HD float hg_phase(float cos_theta, float g) {  // medium.rs:145-149
    float denom = fmaxf(1.0f + g * g - 2.0f * g * cos_theta, 1e-6f);
    return (1.0f - g * g) / (4.0f * PI_F * denom * sqrtf(denom));
}

struct VolumeRegion {
    // world_to_local = local_to_world.inverse() (volume.rs:222); glam Mat4 columns, rows 0..2 — transform_point3 /
    // transform_vector3 never read the bottom row
    V3 wx, wy, wz, ww;
    V3 half_extent;
    V3 aabb_min, aabb_max;  // world box of the 8 transformed corners (:207-219)
    V3 sigma_s, sigma_a;    // at density 1, densityScale folded in (:205-206)
    V3 emission;
    float g;                // clamped to +-0.99 (:227)
    float majorant;         // max_channel(sigma_a + sigma_s) * field.max_value() (:220)
    uint32_t field;
    float scale;            // FIELD_NOISE (:27-40)
    uint32_t octaves;
    float gain, lacunarity, threshold;
    uint32_t seed;
    uint32_t nx, ny, nz;    // FIELD_GRID (:43-48), x-fastest
    const float* data;
};

// ---- DensityField (volume.rs:50-166) ------------------------------------------------------------------
CRT_HD float vol_hash3(int32_t ix, int32_t iy, int32_t iz, uint32_t seed) {  // :88-99
    uint32_t h = ((uint32_t)ix * 0x8da6b343u) ^ ((uint32_t)iy * 0xd8163841u) ^ ((uint32_t)iz * 0xcb1ab31fu) ^ (seed * 0x9e3779b9u);
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    h *= 0x297a2d39u;
    h ^= h >> 15;
    return (float)(h >> 8) / 16777216.0f;
}
CRT_HD float vol_smoothstep(float t) { return t * t * (3.0f - 2.0f * t); }
CRT_HD float vol_value_noise(V3 p, float freq, uint32_t seed) {  // :106-125
    V3 q = p * freq;
    V3 base = v3(floorf(q.x), floorf(q.y), floorf(q.z));
    int32_t ix = (int32_t)base.x, iy = (int32_t)base.y, iz = (int32_t)base.z;
    V3 f = q - base;
    float fx = vol_smoothstep(f.x), fy = vol_smoothstep(f.y), fz = vol_smoothstep(f.z);
    float c[8];
    for (int n = 0; n < 8; ++n) c[n] = vol_hash3(ix + (n & 1), iy + ((n >> 1) & 1), iz + ((n >> 2) & 1), seed);
    float x00 = c[0] + (c[1] - c[0]) * fx;
    float x10 = c[2] + (c[3] - c[2]) * fx;
    float x01 = c[4] + (c[5] - c[4]) * fx;
    float x11 = c[6] + (c[7] - c[6]) * fx;
    float y0 = x00 + (x10 - x00) * fy;
    float y1 = x01 + (x11 - x01) * fy;
    return y0 + (y1 - y0) * fz;
}
CRT_HD float vol_fbm(V3 p, float scale, uint32_t octaves, float gain, float lacunarity, uint32_t seed) {  // :128-141
    octaves = octaves > 1u ? octaves : 1u;
    float sum = 0.0f, norm = 0.0f, amp = 1.0f, freq = scale;
    for (uint32_t o = 0; o < octaves; ++o) {
        sum += amp * vol_value_noise(p, freq, seed + o);
        norm += amp;
        amp *= gain;
        freq *= lacunarity;
    }
    return sum / fmaxf(norm, 1e-6f);
}
CRT_HD void vol_grid_coord(float v, uint32_t n, uint32_t& i0, uint32_t& i1, float& f) {  // :146-156
    float x = v * (float)n - 0.5f;
    float i = floorf(x);
    f = x - i;
    float ic = fmaxf(i, 0.0f);
    i0 = ic >= (float)(n - 1) ? n - 1 : (uint32_t)ic;
    i1 = i0 + 1 < n - 1 ? i0 + 1 : n - 1;
    f = i < 0.0f ? 0.0f : fminf(f, 1.0f);
}
CRT_HD float vol_grid_trilinear(V3 u, uint32_t nx, uint32_t ny, uint32_t nz, const float* data) {  // :144-171
    uint32_t x0, x1, y0, y1, z0, z1;
    float fx, fy, fz;
    vol_grid_coord(u.x, nx, x0, x1, fx);
    vol_grid_coord(u.y, ny, y0, y1, fy);
    vol_grid_coord(u.z, nz, z0, z1, fz);
    const float wzs[2] = {1.0f - fz, fz}, wys[2] = {1.0f - fy, fy}, wxs[2] = {1.0f - fx, fx};
    const uint32_t zs[2] = {z0, z1}, ys[2] = {y0, y1}, xs[2] = {x0, x1};
    float out = 0.0f;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) out += wzs[a] * wys[b] * wxs[c] * data[xs[c] + (size_t)nx * (ys[b] + (size_t)ny * zs[a])];
    return out;
}
CRT_HD float field_density(const VolumeRegion& r, V3 u) {  // :52-72
    if (r.field == FIELD_HOMOGENEOUS) return 1.0f;
    if (r.field == FIELD_NOISE) {
        float fbm = vol_fbm(u, r.scale, r.octaves, r.gain, r.lacunarity, r.seed);
        float t = clampf(r.threshold, 0.0f, 0.999f);
        return fmaxf((fbm - t) / (1.0f - t), 0.0f);
    }
    return vol_grid_trilinear(u, r.nx, r.ny, r.nz, r.data);
}

// ---- VolumeRegion (volume.rs:232-283) ----------------------------------------------------------------
CRT_HD V3 region_point(const VolumeRegion& r, V3 p) { return ((r.wx * p.x + r.wy * p.y) + r.wz * p.z) + r.ww; }  // Mat4::transform_point3
CRT_HD V3 region_vector(const VolumeRegion& r, V3 v) { return (r.wx * v.x + r.wy * v.y) + r.wz * v.z; }          // Mat4::transform_vector3
CRT_HD float region_density(const VolumeRegion& r, V3 p_world) {                                                // :233-241
    V3 p = region_point(r, p_world);
    V3 h = r.half_extent;
    if (fabsf(p.x) > h.x || fabsf(p.y) > h.y || fabsf(p.z) > h.z) return 0.0f;
    V3 u = (p + h) / (h * 2.0f);
    return field_density(r, u);
}
CRT_HD bool region_intersect(const VolumeRegion& r, V3 ro, V3 rd, float& t0_out, float& t1_out) {  // :247-272
    V3 o = region_point(r, ro), d = region_vector(r, rd);
    const float oa[3] = {o.x, o.y, o.z}, da[3] = {d.x, d.y, d.z}, ha[3] = {r.half_extent.x, r.half_extent.y, r.half_extent.z};
    float t0 = 0.0f, t1 = INFINITY;
    for (int a = 0; a < 3; ++a) {
        float h = ha[a];
        if (fabsf(da[a]) < 1e-9f) {
            if (fabsf(oa[a]) > h) return false;
            continue;
        }
        float inv = 1.0f / da[a];
        float ta = (-h - oa[a]) * inv, tb = (h - oa[a]) * inv;
        if (ta > tb) {
            float s = ta;
            ta = tb;
            tb = s;
        }
        t0 = fmaxf(t0, ta);
        t1 = fminf(t1, tb);
        if (t1 <= t0) return false;
    }
    t0_out = t0;
    t1_out = t1;
    return true;
}
CRT_HD bool aabb_hit(V3 mn, V3 mx, V3 ro, V3 rd, float t_min, float t_max) {  // crust-rt aabb.rs:24-42
    const float lo[3] = {mn.x, mn.y, mn.z}, hi[3] = {mx.x, mx.y, mx.z}, o[3] = {ro.x, ro.y, ro.z}, d[3] = {rd.x, rd.y, rd.z};
    for (int a = 0; a < 3; ++a) {
        float inv_d = 1.0f / d[a];
        float t0 = (lo[a] - o[a]) * inv_d, t1 = (hi[a] - o[a]) * inv_d;
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

// ---- PhaseMix (volume.rs:285-329) ----------------------------------------------------------------------
struct PhaseMix {
    int n;
    float w[MAX_VOLUME_REGIONS], g[MAX_VOLUME_REGIONS];  // weights sum to 1
};
CRT_HD V3 phase_sample(const PhaseMix& m, V3 wi, float lobe_u, float u1, float u2) {  // :306-317
    float pick = lobe_u;
    float g = m.g[m.n - 1];
    for (int i = 0; i < m.n; ++i) {
        if (pick < m.w[i]) {
            g = m.g[i];
            break;
        }
        pick -= m.w[i];
    }
    return sample_henyey_greenstein(wi, g, u1, u2);
}
CRT_HD float phase_pdf(const PhaseMix& m, float cos_theta) {  // :320-325
    float s = 0.0f;
    for (int i = 0; i < m.n; ++i) s += m.w[i] * hg_phase(cos_theta, m.g[i]);
    return s;
}

// ---- Volumes (volume.rs:331-540) --------------------------------------------------------------------------
struct VolumeSpans {
    int n;
    bool overflow;  // the segment crosses more regions than the list holds: the excess is ignored (and reported)
    int idx[MAX_VOLUME_REGIONS];
    float a[MAX_VOLUME_REGIONS], b[MAX_VOLUME_REGIONS];
    float majorant, start, end;
};
CRT_HD void volumes_active_intervals(const VolumeRegion* regions, int n_regions, V3 ro, V3 rd, float t_eps, float t_max, VolumeSpans& s) {  // :386-410
    s.n = 0;
    s.overflow = false;
    s.majorant = 0.0f;
    s.start = INFINITY;  // :424-425
    s.end = 0.0f;
    for (int i = 0; i < n_regions; ++i) {
        const VolumeRegion& r = regions[i];
        if (r.majorant <= 0.0f) continue;
        if (!aabb_hit(r.aabb_min, r.aabb_max, ro, rd, t_eps, t_max)) continue;
        float t0, t1;
        if (!region_intersect(r, ro, rd, t0, t1)) continue;
        float a = fmaxf(t0, t_eps), b = fminf(t1, t_max);
        if (b > a && s.n >= MAX_VOLUME_REGIONS) s.overflow = true;
        if (b > a && s.n < MAX_VOLUME_REGIONS) {
            s.idx[s.n] = i;
            s.a[s.n] = a;
            s.b[s.n] = b;
            ++s.n;
            s.majorant += r.majorant;
            s.start = fminf(s.start, a);
            s.end = fmaxf(s.end, b);
        }
    }
}

struct VolumeEvent {  // volume.rs:331-362
    bool scatter;
    float t;
    V3 p;
    V3 weight;   // Scatter: the event's path weight; Passthrough: the transmittance estimate
    V3 emitted;  // pre-weighted emission along the walked part of the segment
    PhaseMix phase;
};

// Volumes::sample_interaction (:415-487)
// `overflow` (optional) is set when the segment crosses more than MAX_VOLUME_REGIONS regions at once — the one limit left:
// a world may hold any number of regions (the reference's Vec, volume.rs:356-366), a single ray segment is walked through
// at most that many.
CRT_HD_NOINLINE void volumes_sample_interaction(const VolumeRegion* regions, int n_regions, V3 ro, V3 rd, float t_eps, float t_max, Rng& rng, VolumeEvent& ev,
                                                bool* overflow = nullptr) {
    VolumeSpans sp;
    volumes_active_intervals(regions, n_regions, ro, rd, t_eps, t_max, sp);
    if (overflow && sp.overflow) *overflow = true;
    ev.scatter = false;
    ev.weight = splat(1.0f);
    ev.emitted = splat(0.0f);
    ev.phase.n = 0;
    if (sp.n == 0 || sp.majorant <= 0.0f) return;
    const float majorant = sp.majorant;
    float t = sp.start;
    V3 w = splat(1.0f), emitted = splat(0.0f);
    for (;;) {
        t += -logf(1.0f - rng_next_f32(rng)) / majorant;
        if (t >= sp.end) {
            ev.weight = w;
            ev.emitted = emitted;
            return;
        }
        V3 p = ro + rd * t;  // Ray::at
        V3 sigma_s_x = splat(0.0f), sigma_t_x = splat(0.0f);
        int nl = 0;
        for (int k = 0; k < sp.n; ++k) {
            if (t < sp.a[k] || t > sp.b[k]) continue;
            const VolumeRegion& r = regions[sp.idx[k]];
            float d = region_density(r, p);
            if (d <= 0.0f) continue;
            V3 ss = r.sigma_s * d;
            sigma_s_x += ss;
            sigma_t_x += (r.sigma_a + r.sigma_s) * d;
            emitted += w * (r.sigma_a * d) * r.emission / majorant;
            float m = max_element(ss);
            if (m > 0.0f) {
                ev.phase.w[nl] = m;
                ev.phase.g[nl] = r.g;
                ++nl;
            }
        }
        float p_scatter = clampf(max_element(sigma_s_x) / majorant, 0.0f, 1.0f);
        if (rng_next_f32(rng) < p_scatter) {
            float total = 0.0f;
            for (int k = 0; k < nl; ++k) total += ev.phase.w[k];
            for (int k = 0; k < nl; ++k) ev.phase.w[k] /= total;
            ev.phase.n = nl;
            ev.scatter = true;
            ev.t = t;
            ev.p = p;
            ev.weight = w * sigma_s_x / (majorant * p_scatter);
            ev.emitted = emitted;
            return;
        }
        w *= (splat(majorant) - sigma_t_x) / (majorant * (1.0f - p_scatter));
        if (max_element(w) < 1e-5f) {
            ev.weight = splat(0.0f);
            ev.emitted = emitted;
            return;
        }
    }
}

// Volumes::transmittance (:493-539)
CRT_HD_NOINLINE V3 volumes_transmittance(const VolumeRegion* regions, int n_regions, V3 ro, V3 rd, float t_eps, float t_max, Rng& rng, bool* overflow = nullptr) {
    VolumeSpans sp;
    volumes_active_intervals(regions, n_regions, ro, rd, t_eps, t_max, sp);
    if (overflow && sp.overflow) *overflow = true;
    if (sp.n == 0 || sp.majorant <= 0.0f) return splat(1.0f);
    bool all_homogeneous = true;
    for (int k = 0; k < sp.n; ++k) all_homogeneous = all_homogeneous && regions[sp.idx[k]].field == FIELD_HOMOGENEOUS;
    if (all_homogeneous) {
        V3 tr = splat(1.0f);
        for (int k = 0; k < sp.n; ++k) {
            const VolumeRegion& r = regions[sp.idx[k]];
            V3 e = ((r.sigma_a + r.sigma_s) * 1.0f) * (sp.b[k] - sp.a[k]);
            tr *= v3(expf(-e.x), expf(-e.y), expf(-e.z));
        }
        return tr;
    }
    const float majorant = sp.majorant;
    float t = sp.start;
    V3 w = splat(1.0f);
    for (;;) {
        t += -logf(1.0f - rng_next_f32(rng)) / majorant;
        if (t >= sp.end) return w;
        V3 p = ro + rd * t;
        V3 sigma_t_x = splat(0.0f);
        for (int k = 0; k < sp.n; ++k) {
            if (t < sp.a[k] || t > sp.b[k]) continue;
            const VolumeRegion& r = regions[sp.idx[k]];
            sigma_t_x += (r.sigma_a + r.sigma_s) * region_density(r, p);
        }
        w *= (splat(majorant) - sigma_t_x) / majorant;
        if (max_element(w) < 1e-5f) return splat(0.0f);
    }
}

// ---- VolumeRegion::new (volume.rs:194-231), host only --------------------------------------------------------
// glam Mat4::inverse (the cofactor expansion glam shares with GLM), on column-major m[16].
inline void mat4_inverse(const float m[16], float out[16]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3], m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7], m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11],
                m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    const float coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    const float coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    const float coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    const float coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    const float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    const float fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07}, fac2[4] = {coef08, coef08, coef10, coef11};
    const float fac3[4] = {coef12, coef12, coef14, coef15}, fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    const float vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01}, vec2[4] = {m12, m02, m02, m02}, vec3[4] = {m13, m03, m03, m03};
    const float sign_a[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sign_b[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
    float inv[16];
    for (int k = 0; k < 4; ++k) {
        inv[0 + k] = ((vec1[k] * fac0[k] - vec2[k] * fac1[k]) + vec3[k] * fac2[k]) * sign_a[k];
        inv[4 + k] = ((vec0[k] * fac0[k] - vec2[k] * fac3[k]) + vec3[k] * fac4[k]) * sign_b[k];
        inv[8 + k] = ((vec0[k] * fac1[k] - vec1[k] * fac3[k]) + vec3[k] * fac5[k]) * sign_a[k];
        inv[12 + k] = ((vec0[k] * fac2[k] - vec1[k] * fac4[k]) + vec2[k] * fac5[k]) * sign_b[k];
    }
    const float d0 = m00 * inv[0], d1 = m01 * inv[4], d2 = m02 * inv[8], d3 = m03 * inv[12];
    const float rcp_det = 1.0f / (((d0 + d1) + d2) + d3);
    for (int k = 0; k < 16; ++k) out[k] = inv[k] * rcp_det;
}
struct VolumeFieldDesc {
    uint32_t field = FIELD_HOMOGENEOUS;
    float scale = 4.0f;
    uint32_t octaves = 4;
    float gain = 0.5f, lacunarity = 2.0f, threshold = 0.3f;
    uint32_t seed = 0;
    uint32_t nx = 0, ny = 0, nz = 0;
    const float* data = nullptr;
};
inline VolumeRegion volume_region_new(const float local_to_world[16], V3 half_extent, V3 sigma_s, V3 sigma_a, float g, V3 emission, float density_scale,
                                      const VolumeFieldDesc& f) {
    VolumeRegion r{};
    sigma_s = sigma_s * density_scale;
    sigma_a = sigma_a * density_scale;
    const float* m = local_to_world;
    const V3 cx = v3(m[0], m[1], m[2]), cy = v3(m[4], m[5], m[6]), cz = v3(m[8], m[9], m[10]), cw = v3(m[12], m[13], m[14]);
    V3 mn = splat(INFINITY), mx = splat(-INFINITY);
    for (int n = 0; n < 8; ++n) {
        V3 c = v3((n & 1) == 0 ? -half_extent.x : half_extent.x, (n & 2) == 0 ? -half_extent.y : half_extent.y, (n & 4) == 0 ? -half_extent.z : half_extent.z);
        V3 w = ((cx * c.x + cy * c.y) + cz * c.z) + cw;
        mn = vmin(mn, w);
        mx = vmax(mx, w);
    }
    float max_value = 1.0f;  // DensityField::max_value (:75-83)
    if (f.field == FIELD_GRID) {
        max_value = 0.0f;
        for (size_t i = 0; i < (size_t)f.nx * f.ny * f.nz; ++i) max_value = fmaxf(max_value, f.data[i]);
    }
    float inv[16];
    mat4_inverse(m, inv);
    r.wx = v3(inv[0], inv[1], inv[2]);
    r.wy = v3(inv[4], inv[5], inv[6]);
    r.wz = v3(inv[8], inv[9], inv[10]);
    r.ww = v3(inv[12], inv[13], inv[14]);
    r.half_extent = half_extent;
    r.aabb_min = mn;
    r.aabb_max = mx;
    r.sigma_s = sigma_s;
    r.sigma_a = sigma_a;
    r.emission = emission;
    r.g = clampf(g, -0.99f, 0.99f);
    r.majorant = max_element(sigma_a + sigma_s) * max_value;
    r.field = f.field;
    r.scale = f.scale;
    r.octaves = f.octaves;
    r.gain = f.gain;
    r.lacunarity = f.lacunarity;
    r.threshold = f.threshold;
    r.seed = f.seed;
    r.nx = f.nx;
    r.ny = f.ny;
    r.nz = f.nz;
    r.data = f.data;
    return r;
}

}  // namespace shade
