// Lat-long environment maps for dome lights: direction <-> pixel mapping and luminance importance
// sampling. Restates crates/crust-core/src/environment.rs (:24-92 Distribution1D, :106-260 EnvironmentMap).
// Host + device: the arrays an EnvMap points at live in host memory for the oracle and in HBM for the kernels.
#pragma once
#include "base.h"

#include <vector>

namespace shade {

struct EnvMap {
    int32_t width, height;
    const float* pixels;         // 3 floats per texel, row 0 = +Y pole (environment.rs:9-16)
    const float* marg_func;      // [height]   row integrals (luminance * sin theta)
    const float* marg_cdf;       // [height + 1]
    const float* cond_func;      // [height * width]
    const float* cond_cdf;       // [height * (width + 1)]
    const float* cond_integral;  // [height]
    float marg_integral;
    int32_t pad;
};

// Distribution1D::sample (:58-79): the first index whose cdf exceeds `u`, minus one.
CRT_HD void dist_sample(const float* func, const float* cdf, int n, float integral, float u, float& x, float& pdf, int& bin) {
    int lo = 0, hi = n + 1;  // lower bound of u in cdf[0 .. n]
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    int b;
    if (lo <= n && cdf[lo] == u)
        b = lo < n - 1 ? lo : n - 1;  // Ok(i) => i.min(len - 1)
    else
        b = (lo > 0 ? lo - 1 : 0) < n - 1 ? (lo > 0 ? lo - 1 : 0) : n - 1;  // Err(i) => i.saturating_sub(1).min(len - 1)
    float span = cdf[b + 1] - cdf[b];
    float within = span > 0.0f ? (u - cdf[b]) / span : 0.5f;
    x = ((float)b + within) / (float)n;
    pdf = integral > 0.0f ? func[b] / integral : 1.0f;
    bin = b;
}
CRT_HD void env_direction_to_uv(V3 d, float& u, float& v) {  // :151-155
    v = acosf(clampf(d.y, -1.0f, 1.0f)) / PI_F;
    u = 0.5f + atan2f(d.x, -d.z) / 6.28318530717958647692f;
    u = u - floorf(u);  // rem_euclid(1.0)
    if (u >= 1.0f) u = 0.0f;
    v = clampf(v, 0.0f, 1.0f);
}
CRT_HD V3 env_uv_to_direction(float u, float v) {  // :157-162
    float theta = v * PI_F, phi = (u - 0.5f) * 6.28318530717958647692f;
    float sin_theta = sinf(theta);
    return v3(sin_theta * sinf(phi), cosf(theta), -sin_theta * cosf(phi));
}
CRT_HD void env_texel(const EnvMap& m, V3 direction, int& x, int& y, float& v_out) {
    float u, v;
    env_direction_to_uv(direction, u, v);
    x = (int)(u * (float)m.width);
    y = (int)(v * (float)m.height);
    x = x < m.width - 1 ? x : m.width - 1;
    y = y < m.height - 1 ? y : m.height - 1;
    v_out = v;
}
CRT_HD_NOINLINE V3 env_radiance(const EnvMap& m, V3 direction) {  // :165-170
    int x, y;
    float v;
    env_texel(m, direction, x, y, v);
    const float* p = m.pixels + 3 * ((size_t)y * m.width + x);
    return v3(p[0], p[1], p[2]);
}
CRT_HD float env_solid_angle_pdf(float pdf_uv, float v) {  // :197-203
    float sin_theta = sinf(v * PI_F);
    if (sin_theta <= 0.0f) return 0.0f;
    return pdf_uv / (2.0f * PI_F * PI_F * sin_theta);
}
CRT_HD_NOINLINE bool env_sample(const EnvMap& m, float u1, float u2, V3& direction, V3& radiance, float& pdf) {  // :174-184
    if (m.marg_integral <= 0.0f) return false;
    float v, pdf_v, u, pdf_u;
    int row, col;
    dist_sample(m.marg_func, m.marg_cdf, m.height, m.marg_integral, u2, v, pdf_v, row);
    dist_sample(m.cond_func + (size_t)row * m.width, m.cond_cdf + (size_t)row * (m.width + 1), m.width, m.cond_integral[row], u1, u, pdf_u, col);
    direction = env_uv_to_direction(u, v);
    pdf = env_solid_angle_pdf(pdf_u * pdf_v, v);
    if (!(pdf > 0.0f)) return false;
    radiance = env_radiance(m, direction);
    return true;
}
CRT_HD_NOINLINE float env_pdf(const EnvMap& m, V3 direction) {  // :188-195
    if (m.marg_integral <= 0.0f) return 0.0f;
    int x, y;
    float v;
    env_texel(m, direction, x, y, v);
    float ci = m.cond_integral[y];
    float pc = ci > 0.0f ? m.cond_func[(size_t)y * m.width + x] / ci : 1.0f;
    float pm = m.marg_func[y] / m.marg_integral;
    return env_solid_angle_pdf(pc * pm, v);
}

// EnvironmentMap::new (:117-141) + Distribution1D::new (:31-56): host-side tables.
struct EnvMapTables {
    int32_t width = 0, height = 0;
    std::vector<float> pixels, marg_func, marg_cdf, cond_func, cond_cdf, cond_integral;
    float marg_integral = 0.0f;
    static void build_dist(const float* func, int n, float* cdf, float& integral) {
        cdf[0] = 0.0f;
        double running = 0.0;
        for (int i = 0; i < n; ++i) {
            running += (double)func[i] / (double)n;
            cdf[i + 1] = (float)running;
        }
        integral = (float)running;
        if (integral > 0.0f) {
            for (int i = 0; i <= n; ++i) cdf[i] /= integral;
        } else {
            for (int i = 0; i <= n; ++i) cdf[i] = (float)i / (float)n;
        }
    }
    bool build(int32_t w, int32_t h, const float* rgb) {
        if (w <= 0 || h <= 0 || !rgb) return false;
        width = w, height = h;
        pixels.assign(rgb, rgb + 3 * (size_t)w * h);
        cond_func.resize((size_t)w * h);
        cond_cdf.resize((size_t)(w + 1) * h);
        cond_integral.resize(h);
        marg_func.resize(h);
        marg_cdf.resize(h + 1);
        for (int y = 0; y < h; ++y) {
            float theta = ((float)y + 0.5f) / (float)h * PI_F;
            float sin_theta = sinf(theta);
            for (int x = 0; x < w; ++x) {
                const float* p = rgb + 3 * ((size_t)y * w + x);
                cond_func[(size_t)y * w + x] = fmaxf(luminance(v3(p[0], p[1], p[2])), 0.0f) * sin_theta;
            }
            build_dist(cond_func.data() + (size_t)y * w, w, cond_cdf.data() + (size_t)y * (w + 1), cond_integral[y]);
            marg_func[y] = cond_integral[y];
        }
        build_dist(marg_func.data(), h, marg_cdf.data(), marg_integral);
        return true;
    }
    EnvMap view() const {  // pointing at these host vectors
        return EnvMap{width, height, pixels.data(), marg_func.data(), marg_cdf.data(), cond_func.data(), cond_cdf.data(), cond_integral.data(), marg_integral, 0};
    }
};

}  // namespace shade
