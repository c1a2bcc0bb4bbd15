// Lights, camera, pixel filter and render settings as plain records, with the
// sampling functions the integrator needs. Restates
//   crates/crust-core/src/light.rs:21-446   (area / distant / dome lights, LightList)
//   crates/crust-core/src/camera.rs:25-84   (thin-lens camera)
//   crates/crust-core/src/filter.rs:28-260  (filter importance sampling)
//   crates/crust-core/src/tracer.rs:62-105,639-661 (SamplingStrategy, RenderSettings)
// Host + device. Environment maps of dome lights: envmap.h.
#pragma once
#include "envmap.h"
#include "material.h"

namespace shade {

// ---- lights -----------------------------------------------------------------------------
enum : uint32_t { LIGHT_SPHERE = 0, LIGHT_RECT = 1, LIGHT_DISTANT = 2, LIGHT_DOME = 3 };
constexpr uint32_t NO_GEOM = 0xFFFFFFFFu;

struct Light {
    uint32_t kind;
    uint32_t geom_id;  // world geom_id of the emissive geometry (area lights), else NO_GEOM
    V3 radiance;       // area: Emissive colour; distant: irradiance; dome: tint
    // sphere: a = center, r = radius. rect: a = origin, b = edge_u, c = edge_v, n = unit normal.
    // distant: n = unit travel direction, r = cos(half angle), s = solid angle.
    V3 a, b, c, n;
    float r, s;
    // dome with an environment map (light.rs:320-389): index into the world's EnvMap table (-1 = uniform dome),
    // light_to_world in a, b, c (columns) and world_to_light in wa, wb, wc (columns)
    int32_t env;
    V3 wa, wb, wc;
};
CRT_HD V3 mat3_mul(V3 c0, V3 c1, V3 c2, V3 v) { return (c0 * v.x + c1 * v.y) + c2 * v.z; }  // glam Mat3A * Vec3A
struct LightSample {  // light.rs:92-111
    V3 direction, radiance;
    float distance, pdf;
};
CRT_HD V3 light_normal_at(const Light& l, V3 p) { return l.kind == LIGHT_SPHERE ? normalize(p - l.a) : l.n; }  // :38-40, 72-74
CRT_HD float light_area(const Light& l) {                                                                     // :42-44, 76-78
    return l.kind == LIGHT_SPHERE ? 4.0f * PI_F * l.r * l.r : length(cross(l.b, l.c));
}
CRT_HD V3 light_sample_point(const Light& l, float u, float v) {  // :27-36, 68-70
    if (l.kind == LIGHT_SPHERE) {
        float theta = 2.0f * PI_F * u;
        float phi = acosf(1.0f - 2.0f * v);
        V3 n = v3(sinf(phi) * cosf(theta), sinf(phi) * sinf(theta), cosf(phi));
        return l.a + l.r * n;
    }
    return l.a + u * l.b + v * l.c;
}
CRT_HD float light_solid_angle_pdf(const Light& l, V3 from, V3 light_point) {  // :191-198
    V3 direction = light_point - from;
    float d2 = length_squared(direction);
    V3 dir_to_light = normalize(direction);
    float cosine = fmaxf(dot(light_normal_at(l, light_point), -dir_to_light), 0.0f);
    return d2 / (cosine * light_area(l) + 1e-4f);
}
// Lights at infinity, out of line: a shading kernel keeps only the area-light path inline (code that is compiled into
// the hot instruction stream but never taken still costs instruction-cache lines).
CRT_HD_NOINLINE bool light_sample_li_infinite(const Light& l, float u, float v, LightSample& out, const EnvMap* envs) {
    if (l.kind == LIGHT_DISTANT) {  // :281-295
        float cos_theta = 1.0f - u * (1.0f - l.r);
        float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
        float phi = 2.0f * PI_F * v;
        V3 local = v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
        out.direction = normalize(align_to_normal(local, -l.n));
        out.distance = INFINITY;
        out.radiance = l.radiance / fmaxf(l.s, 1e-12f);
        out.pdf = 1.0f / fmaxf(l.s, 1e-12f);
        return true;
    }
    if (l.env >= 0 && envs) {  // environment-mapped dome :352-363
        V3 local, radiance;
        float pdf;
        if (!env_sample(envs[l.env], u, v, local, radiance, pdf)) return false;
        out.direction = normalize(mat3_mul(l.a, l.b, l.c, local));
        out.distance = INFINITY;
        out.radiance = l.radiance * radiance;
        out.pdf = pdf;
        return pdf > 0.0f;
    }
    // uniform dome :365-372
    float z = 1.0f - 2.0f * u;
    float rr = sqrtf(fmaxf(1.0f - z * z, 0.0f));
    float phi = 6.28318530717958647692f * v;
    out.direction = v3(rr * cosf(phi), z, rr * sinf(phi));
    out.distance = INFINITY;
    out.radiance = l.radiance;
    out.pdf = 1.0f / (4.0f * PI_F);
    return true;
}
CRT_HD bool light_sample_li(const Light& l, V3 from, float u, float v, LightSample& out, const EnvMap* envs = nullptr) {
    if (l.kind <= LIGHT_RECT) {  // AreaLight::sample_li :202-215
        V3 lp = light_sample_point(l, u, v);
        V3 to_light = lp - from;
        float distance = length(to_light);
        if (distance < 1e-6f) return false;
        out.direction = to_light / distance;
        out.distance = distance;
        out.radiance = l.radiance;
        out.pdf = light_solid_angle_pdf(l, from, lp);
        return true;
    }
    return light_sample_li_infinite(l, u, v, out, envs);
}
CRT_HD float light_pdf_at_point(const Light& l, V3 from, V3 light_point) {  // :217-219, default 0
    return l.kind <= LIGHT_RECT ? light_solid_angle_pdf(l, from, light_point) : 0.0f;
}
CRT_HD_NOINLINE bool light_escaped(const Light& l, V3 direction, V3& radiance, float& pdf, const EnvMap* envs = nullptr) {  // :297-300, 384-387
    if (l.kind == LIGHT_DISTANT) {
        if (!(dot(direction, -l.n) >= l.r)) return false;
        radiance = l.radiance / fmaxf(l.s, 1e-12f);
        pdf = 1.0f / fmaxf(l.s, 1e-12f);
        return true;
    }
    if (l.kind == LIGHT_DOME) {
        if (l.env >= 0 && envs) {  // radiance_toward / pdf_toward :334-347
            V3 local = mat3_mul(l.wa, l.wb, l.wc, direction);
            radiance = l.radiance * env_radiance(envs[l.env], local);
            pdf = env_pdf(envs[l.env], local);
            return true;
        }
        radiance = l.radiance;
        pdf = 1.0f / (4.0f * PI_F);
        return true;
    }
    return false;
}
CRT_HD int light_pick(int n_lights, float u) {  // LightList::pick :421-430 (-1 = None)
    if (n_lights == 0) return -1;
    float f = u * (float)n_lights;
    int i = f > 0.0f ? (int)f : 0;
    return i < n_lights - 1 ? i : n_lights - 1;
}
CRT_HD int light_find_by_geom(const Light* lights, int n, uint32_t geom_id) {  // :436-440
    for (int i = 0; i < n; ++i)
        if (lights[i].geom_id != NO_GEOM && lights[i].geom_id == geom_id) return i;
    return -1;
}

// ---- camera (camera.rs) --------------------------------------------------------------------
struct Camera {
    V3 origin, lower_left_corner, horizontal, vertical, u, v;
    float lens_radius;
};
CRT_HD Camera camera_new(V3 lookfrom, V3 lookat, V3 vup, float vfov_deg, float aspect, float aperture, float focus_dist) {  // :27-59
    float theta = vfov_deg * PI_F / 180.0f;  // utils::degrees_to_radians
    float h = tanf(theta / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width = aspect * viewport_height;
    V3 w = normalize(lookfrom - lookat);
    V3 u = normalize(cross(vup, w));
    V3 v = cross(w, u);
    Camera c;
    c.origin = lookfrom;
    c.horizontal = focus_dist * viewport_width * u;
    c.vertical = focus_dist * viewport_height * v;
    c.lower_left_corner = c.origin - c.horizontal / 2.0f - c.vertical / 2.0f - focus_dist * w;
    c.u = u;
    c.v = v;
    c.lens_radius = aperture / 2.0f;
    return c;
}
CRT_HD void camera_get_ray(const Camera& c, float s, float t, float lens_u, float lens_v, V3& origin, V3& dir) {  // :71-84
    V3 offset = v3(0, 0, 0);
    if (c.lens_radius > 0.0f) {
        V3 rd = c.lens_radius * concentric_disk(lens_u, lens_v);
        offset = c.u * rd.x + c.v * rd.y;
    }
    origin = c.origin + offset;
    dir = c.lower_left_corner + s * c.horizontal + t * c.vertical - c.origin - offset;
}

// ---- pixel filter (filter.rs) -----------------------------------------------------------------
enum : uint32_t { FILTER_BOX = 0, FILTER_TRIANGLE = 1, FILTER_GAUSSIAN = 2, FILTER_BLACKMAN = 3, FILTER_MITCHELL = 4 };
constexpr int FILTER_BINS = 64;
struct Filter {
    uint32_t kind;
    float radius;
    // tabulated kinds only (filter.rs:162-172)
    float cdf[FILTER_BINS + 1], abs_mid[FILTER_BINS], abs_sum, signed_sum;
};
CRT_HD float filter_eval(const Filter& f, float x) {  // filter.rs:108-151
    float r = f.radius;
    if (fabsf(x) > r) return 0.0f;
    switch (f.kind) {
        case FILTER_BOX: return 1.0f;
        case FILTER_TRIANGLE: return r - fabsf(x);
        case FILTER_GAUSSIAN: {
            float sigma = r / 3.0f;
            float gx = expf(-0.5f * x * x / (sigma * sigma)), gr = expf(-0.5f * r * r / (sigma * sigma));
            return fmaxf(gx - gr, 0.0f);
        }
        case FILTER_BLACKMAN: {
            float t = (x + r) / (2.0f * r);
            float w = 6.28318530717958647692f * t;
            return 0.35875f - 0.48829f * cosf(w) + 0.14128f * cosf(2.0f * w) - 0.01168f * cosf(3.0f * w);
        }
        default: {
            const float B = 1.0f / 3.0f, C = 1.0f / 3.0f;
            float xx = fabsf(2.0f * x / r);
            float x2 = xx * xx, x3 = x2 * xx;
            float v = xx < 1.0f ? (12.0f - 9.0f * B - 6.0f * C) * x3 + (-18.0f + 12.0f * B + 6.0f * C) * x2 + (6.0f - 2.0f * B)
                                : (-B - 6.0f * C) * x3 + (6.0f * B + 30.0f * C) * x2 + (-12.0f * B - 48.0f * C) * xx + (8.0f * B + 24.0f * C);
            return v / 6.0f;
        }
    }
}
inline Filter filter_new(uint32_t kind, float radius) {  // host: FilterSampler::new + FilterTable::new, filter.rs:175-231
    Filter f{};
    f.kind = kind;
    f.radius = fmaxf(radius, 0.01f);
    if (kind == FILTER_BOX || kind == FILTER_TRIANGLE) return f;
    float r = f.radius;
    f.signed_sum = 0.0f;
    for (int b = 0; b < FILTER_BINS; ++b) {
        float x = -r + ((float)b + 0.5f) * (2.0f * r / (float)FILTER_BINS);
        float v = filter_eval(f, x);
        f.abs_mid[b] = fabsf(v);
        f.signed_sum += v;
    }
    f.abs_sum = 0.0f;
    for (int b = 0; b < FILTER_BINS; ++b) f.abs_sum += f.abs_mid[b];
    f.cdf[0] = 0.0f;
    for (int b = 0; b < FILTER_BINS; ++b) f.cdf[b + 1] = f.cdf[b] + f.abs_mid[b] / f.abs_sum;
    f.cdf[FILTER_BINS] = 1.0f;
    return f;
}
CRT_HD void filter_sample(const Filter& f, float u, float& offset, float& weight) {  // filter.rs:197-222, 236-246
    float r = f.radius;
    if (f.kind == FILTER_BOX) {
        offset = (0.5f - r) + (2.0f * r) * u;
        weight = 1.0f;
        return;
    }
    if (f.kind == FILTER_TRIANGLE) {
        float x = u < 0.5f ? r * (sqrtf(2.0f * u) - 1.0f) : r * (1.0f - sqrtf(2.0f * (1.0f - u)));
        offset = 0.5f + x;
        weight = 1.0f;
        return;
    }
    int pp = 0;  // partition_point over cdf[..BINS]: first index with cdf > u
    while (pp < FILTER_BINS && f.cdf[pp] <= u) ++pp;
    int b = (pp > 1 ? pp : 1) - 1;
    float lo = f.cdf[b], hi = f.cdf[b + 1];
    float t = hi > lo ? (u - lo) / (hi - lo) : 0.5f;
    float x = -r + ((float)b + t) * (2.0f * r / (float)FILTER_BINS);
    float am = fmaxf(f.abs_mid[b], 1.17549435e-38f);
    weight = filter_eval(f, x) * f.abs_sum / (am * f.signed_sum);
    offset = 0.5f + x;
}

// ---- settings (tracer.rs:62-105, 639-661) --------------------------------------------------------
enum : uint32_t { STRATEGY_POWER_MIS = 0, STRATEGY_BALANCE_MIS = 1, STRATEGY_LIGHT_ONLY = 2, STRATEGY_BSDF_ONLY = 3 };
CRT_HD bool strategy_samples_lights(uint32_t s) { return s != STRATEGY_BSDF_ONLY; }
CRT_HD float strategy_light_weight(uint32_t s, float light_pdf, float bounce_pdf) {
    switch (s) {
        case STRATEGY_POWER_MIS: return power_heuristic(light_pdf, bounce_pdf);
        case STRATEGY_BALANCE_MIS: return balance_heuristic(light_pdf, bounce_pdf);
        case STRATEGY_LIGHT_ONLY: return 1.0f;
        default: return 0.0f;
    }
}
CRT_HD float strategy_bounce_weight(uint32_t s, float bounce_pdf, float light_pdf) {
    switch (s) {
        case STRATEGY_POWER_MIS: return power_heuristic(bounce_pdf, light_pdf);
        case STRATEGY_BALANCE_MIS: return balance_heuristic(bounce_pdf, light_pdf);
        case STRATEGY_LIGHT_ONLY: return 0.0f;
        default: return 1.0f;
    }
}

struct RenderSettings {
    uint32_t width, height;
    uint32_t samples_per_pixel, max_depth, min_samples_per_pixel;
    float variance_threshold;
    int32_t frame;
    uint32_t sampling_strategy;
    uint32_t adaptive;  // final passes run with adaptive = true (tracer.rs:216-223)
};

// tracer.rs:20-30 domain keys
enum : int32_t { K_CAMERA = 0, K_PATH = 1, K_TIME = 2, K_NEE = 0, K_NEE_SHADOW = 1, K_BSDF = 2, K_GUIDE = 3, K_PHASE = 4, K_RR = 5, K_MEDIUM = 6, K_VOLUME = 7 };
constexpr int RR_START_BOUNCE = 3;      // tracer.rs:46
constexpr float RR_MIN_PROB = 0.05f;    // tracer.rs:47

}  // namespace shade
