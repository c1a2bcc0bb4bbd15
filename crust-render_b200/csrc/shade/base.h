// Shading-side math shared by host and device code (header-only, CRT_HD).
//
// Restates the pure functions the integrator of crust-core calls:
//   crates/utils/src/common.rs:38-49,128-188  (MIS heuristics, warps, align_to_normal)
//   crates/crust-core/src/lib.rs:21-24        (PathSampler — see below)
// plus a small f32 vector type with glam's operation order (dot = (x + y) + z,
// normalize = v * (1 / sqrt(dot))).
//
// PathSampler. The reference uses `openqmc::SobolSampler` (crate `openqmc-rs
// = "0.1"`, not vendored under /root/reference, no lock file), so its bit
// pattern cannot be reproduced here: PARITY UNPINNED for sample values. This
// header defines the sampler both our CPU oracle and the CUDA kernels use: a
// 4-D Sobol sequence with hash-based Owen scrambling (Burley 2020, "Practical
// Hash-based Owen Scrambling") and OpenQMC's interface — a Copy state made of a
// 16-bit-ish sample index and a 32-bit pattern id, `new(x, y, frame, index)`,
// `new_domain(key)`, `draw_sample<N<=4>()`, `draw_rnd<N>()`. Integer arithmetic
// only, so host and device produce identical values.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CRT_HD __host__ __device__ __forceinline__
#define CRT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CRT_HD inline
#define CRT_HD_NOINLINE inline
#endif

namespace shade {

constexpr float PI_F = 3.14159265358979323846f;

struct V3 {
    float x, y, z;
};
CRT_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
CRT_HD V3 splat(float s) { return V3{s, s, s}; }
CRT_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
CRT_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
CRT_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
CRT_HD V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
CRT_HD V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
CRT_HD V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
CRT_HD V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
CRT_HD V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
CRT_HD V3& operator+=(V3& a, V3 b) {
    a = a + b;
    return a;
}
CRT_HD V3& operator*=(V3& a, V3 b) {
    a = a * b;
    return a;
}
CRT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CRT_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
CRT_HD float length_squared(V3 a) { return dot(a, a); }
CRT_HD float length(V3 a) { return sqrtf(dot(a, a)); }
CRT_HD V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
CRT_HD V3 vmin(V3 a, V3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
CRT_HD V3 vmax(V3 a, V3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
CRT_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
CRT_HD V3 clamp01(V3 a) { return vmin(vmax(a, splat(0.0f)), splat(1.0f)); }
CRT_HD float max_element(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
CRT_HD float min_element(V3 a) { return fminf(fminf(a.x, a.y), a.z); }
CRT_HD V3 lerp(V3 a, V3 b, float s) { return a * (1.0f - s) + b * s; }  // glam >= 0.27
CRT_HD V3 vpowf(V3 a, float e) { return {powf(a.x, e), powf(a.y, e), powf(a.z, e)}; }
CRT_HD V3 vexp(V3 a) { return {expf(a.x), expf(a.y), expf(a.z)}; }
CRT_HD bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
CRT_HD float powi5(float x) {  // Rust powi(5): x*x -> ^4 -> *x (LLVM expands by squaring)
    float x2 = x * x;
    return x2 * x2 * x;
}
CRT_HD float powi6(float x) {
    float x2 = x * x;
    return x2 * x2 * x2;
}
// tracer.rs (guiding::luminance) — Rec.709
CRT_HD float luminance(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }

// ---- utils/src/common.rs ---------------------------------------------------------
CRT_HD float balance_heuristic(float a, float b) { return a / (a + b + 1e-6f); }  // :38-40
CRT_HD float power_heuristic(float a, float b) {                                  // :45-49
    float a2 = a * a, b2 = b * b;
    return a2 / (a2 + b2 + 1e-6f);
}
CRT_HD V3 cosine_hemisphere(float u, float v) {  // :128-135
    float z = sqrtf(1.0f - v);
    float phi = 2.0f * PI_F * u;
    return {cosf(phi) * sqrtf(v), sinf(phi) * sqrtf(v), z};
}
CRT_HD V3 concentric_disk(float u, float v) {  // :158-175
    float sx = 2.0f * u - 1.0f, sy = 2.0f * v - 1.0f;
    if (sx == 0.0f && sy == 0.0f) return v3(0, 0, 0);
    float r, theta;
    if (fabsf(sx) > fabsf(sy)) {
        r = sx;
        theta = 0.78539816339744830962f * (sy / sx);
    } else {
        r = sy;
        theta = 1.57079632679489661923f - 0.78539816339744830962f * (sx / sy);
    }
    return {r * cosf(theta), r * sinf(theta), 0.0f};
}
CRT_HD V3 align_to_normal(V3 local, V3 n) {  // :177-188
    V3 up = fabsf(n.z) < 0.999f ? v3(0, 0, 1) : v3(1, 0, 0);
    V3 t = normalize(cross(n, up));
    V3 b = cross(n, t);
    return local.x * t + local.y * b + local.z * n;
}

// ---- PathSampler -------------------------------------------------------------------
CRT_HD uint32_t reverse_bits32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
CRT_HD uint32_t pcg_hash(uint32_t v) {
    uint32_t s = v * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
CRT_HD uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
// Laine-Karras style permutation with the constants of Burley 2020 (nested uniform scramble in base 2
// when applied to the bit-reversed integer).
CRT_HD uint32_t lk_permute(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
CRT_HD uint32_t owen_scramble(uint32_t x, uint32_t seed) { return reverse_bits32(lk_permute(reverse_bits32(x), seed)); }

// Sobol generator matrices, dimensions 1-4 of Joe & Kuo (new-joe-kuo-6; s/a/m = van der Corput,
// 1/0/{1}, 2/1/{1,3}, 3/1/{1,3,1}), 16 columns each: 2^16 points per pixel. Generated by the
// recurrence in tools/gen_sobol.py; tests/test_sampler.py checks the net properties.
CRT_HD uint32_t sobol_u32(uint32_t index, int dim) {
    if (dim == 0) return reverse_bits32(index);
    const uint32_t D[3][16] = {
        {0x80000000u, 0xc0000000u, 0xa0000000u, 0xf0000000u, 0x88000000u, 0xcc000000u, 0xaa000000u, 0xff000000u,
         0x80800000u, 0xc0c00000u, 0xa0a00000u, 0xf0f00000u, 0x88880000u, 0xcccc0000u, 0xaaaa0000u, 0xffff0000u},
        {0x80000000u, 0xc0000000u, 0x60000000u, 0x90000000u, 0xe8000000u, 0x5c000000u, 0x8e000000u, 0xc5000000u,
         0x68800000u, 0x9cc00000u, 0xee600000u, 0x55900000u, 0x80680000u, 0xc09c0000u, 0x60ee0000u, 0x90550000u},
        {0x80000000u, 0xc0000000u, 0x20000000u, 0x50000000u, 0xf8000000u, 0x74000000u, 0xa2000000u, 0x93000000u,
         0xd8800000u, 0x25400000u, 0x59e00000u, 0xe6d00000u, 0x78080000u, 0xb40c0000u, 0x82020000u, 0xc3050000u}};
    uint32_t x = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k)
        if ((index >> k) & 1u) x ^= D[dim - 1][k];
    return x;
}

struct PathSampler {
    uint32_t pattern;  // domain id
    uint32_t index;    // sample index within the pixel
};
CRT_HD float u32_to_unit_f32(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // [0, 1)
CRT_HD PathSampler sampler_new(int32_t x, int32_t y, int32_t frame, int32_t index) {  // tracer.rs:559-560
    uint32_t h = pcg_hash((uint32_t)x + 0x9E3779B9u);
    h = pcg_hash(h ^ ((uint32_t)y * 0x85EBCA6Bu + 0xC2B2AE35u));
    h = pcg_hash(h ^ ((uint32_t)frame * 0x27D4EB2Fu + 0x165667B1u));
    return PathSampler{h, (uint32_t)index};
}
CRT_HD PathSampler new_domain(PathSampler s, int32_t key) {
    return PathSampler{pcg_hash(hash_combine(s.pattern, (uint32_t)key * 0x9E3779B1u + 0x7F4A7C15u)), s.index};
}
// One 4-D Owen-scrambled Sobol point (all four dimensions are always computed, like the
// reference's block, tracer.rs:549-551).
CRT_HD_NOINLINE void draw_sample4(PathSampler s, float out[4]) {  // one copy per kernel: ~300 instructions, called at every event of a vertex
    uint32_t idx = owen_scramble(s.index, hash_combine(s.pattern, 0x51633E2Du));  // shuffle the sequence order
    idx &= 0xFFFFu;                                                                // 2^16 samples per pixel
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int d = 0; d < 4; ++d) {
        uint32_t x = sobol_u32(idx, d);
        x = owen_scramble(x, hash_combine(s.pattern, 0xA511E9B3u + (uint32_t)d * 0x9E3779B9u));
        out[d] = u32_to_unit_f32(x);
    }
}
// PRNG side (`draw_rnd_f32::<1>`, tracer.rs:1221,1482): unstratified, keyed by domain and index.
CRT_HD float draw_rnd1(PathSampler s) { return u32_to_unit_f32(pcg_hash(hash_combine(s.pattern, s.index * 0x9E3779B1u + 0x632BE5ABu))); }

}  // namespace shade
