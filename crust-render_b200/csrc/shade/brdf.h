// BRDF building blocks — restates crates/crust-core/src/material/brdf.rs
// (line numbers in comments refer to that file). Host + device.
#pragma once
#include "base.h"

namespace shade {

CRT_HD V3 fresnel_schlick(float cos_theta, V3 f0) { return f0 + (splat(1.0f) - f0) * powi5(1.0f - cos_theta); }  // :10-12
CRT_HD float fresnel_schlick_scalar(float cos_theta, float f0) { return f0 + (1.0f - f0) * powi5(1.0f - cos_theta); }  // :15-17
CRT_HD float f0_from_ior(float ior) {  // :28-31
    float r = (ior - 1.0f) / (ior + 1.0f);
    return r * r;
}
CRT_HD void roughness_to_alpha_aniso(float roughness, float anisotropy, float& ax, float& ay) {  // :40-46
    float a = roughness * roughness;
    float inv = 1.0f - clampf(anisotropy, 0.0f, 1.0f);
    float x = a * sqrtf(2.0f / (1.0f + inv * inv));
    float y = inv * x;
    ax = fmaxf(x, 1e-4f);
    ay = fmaxf(y, 1e-4f);
}
CRT_HD float ggx_d_aniso(float n_dot_h, float h_dot_t, float h_dot_b, float ax, float ay) {  // :50-55
    float tx = h_dot_t / ax, ty = h_dot_b / ay;
    float term = tx * tx + ty * ty + n_dot_h * n_dot_h;
    return 1.0f / (PI_F * ax * ay * term * term);
}
CRT_HD float ggx_lambda_aniso(float v_dot_n, float v_dot_t, float v_dot_b, float ax, float ay) {  // :58-64
    float vt = v_dot_t * ax, vb = v_dot_b * ay;
    float a2 = vt * vt + vb * vb;
    float n2 = fmaxf(v_dot_n * v_dot_n, 1e-8f);
    return (-1.0f + sqrtf(1.0f + a2 / n2)) * 0.5f;
}
CRT_HD float ggx_g2_smith_aniso(float vn, float vt, float vb, float ln, float lt, float lb, float ax, float ay) {  // :68-81
    return 1.0f / (1.0f + ggx_lambda_aniso(vn, vt, vb, ax, ay) + ggx_lambda_aniso(ln, lt, lb, ax, ay));
}
CRT_HD V3 sample_vndf_ggx_aniso_local(V3 v_local, float ax, float ay, float u1, float u2) {  // :87-110
    V3 vh = normalize(v3(ax * v_local.x, ay * v_local.y, v_local.z));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    V3 t1 = lensq > 0.0f ? v3(-vh.y, vh.x, 0.0f) / sqrtf(lensq) : v3(1, 0, 0);
    V3 t2 = cross(vh, t1);
    float r = sqrtf(u1);
    float phi = 2.0f * PI_F * u2;
    float t1c = r * cosf(phi);
    float t2c_pre = r * sinf(phi);
    float s = 0.5f * (1.0f + vh.z);
    float t2c = (1.0f - s) * sqrtf(fmaxf(1.0f - t1c * t1c, 0.0f)) + s * t2c_pre;
    V3 nh = t1 * t1c + t2 * t2c + vh * sqrtf(fmaxf(1.0f - t1c * t1c - t2c * t2c, 0.0f));
    return normalize(v3(ax * nh.x, ay * nh.y, fmaxf(nh.z, 0.0f)));
}
CRT_HD float pdf_vndf_ggx_aniso_local(V3 v_local, V3 h_local, float ax, float ay) {  // :114-124
    float n_dot_v = fmaxf(v_local.z, 1e-6f), n_dot_h = fmaxf(h_local.z, 1e-6f);
    float d = ggx_d_aniso(n_dot_h, h_local.x, h_local.y, ax, ay);
    float g1 = 1.0f / (1.0f + ggx_lambda_aniso(n_dot_v, v_local.x, v_local.y, ax, ay));
    return d * g1 / (4.0f * n_dot_v);
}
CRT_HD float pdf_vndf_h_aniso_local(V3 v_local, V3 h_local, float ax, float ay) {  // :130-137
    float n_dot_v = fmaxf(v_local.z, 1e-6f);
    float v_dot_h = fmaxf(dot(v_local, h_local), 0.0f);
    float d = ggx_d_aniso(fmaxf(h_local.z, 1e-6f), h_local.x, h_local.y, ax, ay);
    float g1 = 1.0f / (1.0f + ggx_lambda_aniso(n_dot_v, v_local.x, v_local.y, ax, ay));
    return d * g1 * v_dot_h / n_dot_v;
}

// ---- EON diffuse (:139-213) ---------------------------------------------------------
constexpr float EON_A = 0.5f - 2.0f / (3.0f * PI_F);
constexpr float EON_B = 2.0f / 3.0f - 28.0f / (15.0f * PI_F);
CRT_HD float eon_albedo_exact(float mu, float roughness) {  // :161-168 (tests only)
    mu = clampf(mu, 1e-4f, 1.0f);
    float af = 1.0f / (1.0f + EON_A * roughness);
    float bf = roughness * af;
    float si = sqrtf(fmaxf(1.0f - mu * mu, 0.0f));
    float g = si * (acosf(mu) - si * mu) + (2.0f / 3.0f) * ((si / mu) * (1.0f - si * si * si) - si);
    return af + (bf / PI_F) * g;
}
CRT_HD float eon_albedo_approx(float mu, float roughness) {  // :172-180
    float mucomp = 1.0f - clampf(mu, 0.0f, 1.0f);
    const float G1 = 0.057108529f, G2 = 0.49188187f, G3 = -0.33218144f, G4 = 0.071442995f;
    float g_over_pi = mucomp * (G1 + mucomp * (G2 + mucomp * (G3 + mucomp * G4)));
    return (1.0f + roughness * g_over_pi) / (1.0f + EON_A * roughness);
}
CRT_HD V3 eon_diffuse(V3 rho, float roughness, V3 v_local, V3 l_local) {  // :186-213
    rho = clamp01(rho);
    float mu_i = v_local.z, mu_o = l_local.z;
    float s = dot(v_local, l_local) - mu_i * mu_o;
    float s_over_t = s > 0.0f ? s / fmaxf(fmaxf(mu_i, mu_o), 1e-6f) : s;
    float af = 1.0f / (1.0f + EON_A * roughness);
    V3 f_ss = rho * (af / PI_F) * (1.0f + roughness * s_over_t);
    float e_o = eon_albedo_approx(mu_o, roughness), e_i = eon_albedo_approx(mu_i, roughness);
    float avg_e = af * (1.0f + EON_B * roughness);
    V3 rho_ms = (rho * rho) * avg_e / (splat(1.0f) - rho * (1.0f - avg_e));
    const float EPS = 1.0e-7f;
    V3 f_ms = rho_ms * (1.0f / PI_F) * (fmaxf(1.0f - e_o, EPS) * fmaxf(1.0f - e_i, EPS) / fmaxf(1.0f - avg_e, EPS));
    return f_ss + f_ms;
}

CRT_HD V3 fresnel_f82_tint(float cos_theta, V3 f0, V3 tint) {  // :221-228
    const float MU_BAR = 1.0f / 7.0f;
    float mu = clampf(cos_theta, 0.0f, 1.0f);
    float denom = MU_BAR * powi6(1.0f - MU_BAR);
    V3 fs_bar = f0 + (splat(1.0f) - f0) * powi5(1.0f - MU_BAR);
    V3 a = fs_bar * (splat(1.0f) - tint) / denom;
    V3 fs_mu = f0 + (splat(1.0f) - f0) * powi5(1.0f - mu);
    return clamp01(fs_mu - a * mu * powi6(1.0f - mu));
}
CRT_HD float fresnel_dielectric(float cos_i, float eta_i, float eta_t) {  // :234-244
    cos_i = clampf(cos_i, 0.0f, 1.0f);
    float sin2_t = (eta_i / eta_t) * (eta_i / eta_t) * (1.0f - cos_i * cos_i);
    if (sin2_t >= 1.0f) return 1.0f;
    float cos_t = sqrtf(1.0f - sin2_t);
    float r_par = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t);
    float r_perp = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t);
    return 0.5f * (r_par * r_par + r_perp * r_perp);
}
CRT_HD void tangent_frame(V3 n, V3& t, V3& bt) {  // :249-256 (Duff et al. 2017)
    float sign = n.z >= 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    t = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    bt = v3(b, sign + n.y * n.y * a, -n.y);
}
CRT_HD V3 to_tangent(V3 v, V3 t, V3 b, V3 n) { return {dot(v, t), dot(v, b), dot(v, n)}; }  // :259-261
CRT_HD V3 from_tangent(V3 vl, V3 t, V3 b, V3 n) { return t * vl.x + b * vl.y + n * vl.z; }  // :264-266

CRT_HD float sheen_charlie_d(float n_dot_h, float roughness) {  // :270-275
    float alpha = fmaxf(roughness, 0.05f);
    float inv_alpha = 1.0f / alpha;
    float sin2 = fmaxf(1.0f - n_dot_h * n_dot_h, 0.0f);
    return (2.0f + inv_alpha) * powf(sin2, inv_alpha * 0.5f) / (2.0f * PI_F);
}
CRT_HD float sheen_charlie_v(float n_dot_v, float n_dot_l) {  // :278-280
    return 1.0f / (4.0f * fmaxf(n_dot_l + n_dot_v - n_dot_l * n_dot_v, 1e-4f));
}
CRT_HD float sheen_charlie(float nv, float nl, float nh, float roughness) { return sheen_charlie_d(nh, roughness) * sheen_charlie_v(nv, nl); }  // :283-285

CRT_HD V3 coat_darkening_factor(V3 base_color, float coat_ior, float darkening) {  // :292-297
    float f_avg = f0_from_ior(coat_ior) + (1.0f - f0_from_ior(coat_ior)) * 0.05f;
    V3 dark = base_color / vmax(splat(1.0f) - f_avg * (splat(1.0f) - base_color), splat(1e-4f));
    return splat(1.0f) * (1.0f - darkening) + dark * darkening;
}

// ---- thin film + dispersion (:299-421) -------------------------------------------------
CRT_HD float lambda_rgb(int c) { return c == 0 ? 615.0f : (c == 1 ? 545.0f : 465.0f); }  // :312
CRT_HD float cauchy_ior(float n_d, float v_d, float lambda_nm) {  // :333-341
    const float F = 486.1f, C = 656.3f, D = 587.6f;
    float b = (n_d - 1.0f) / (v_d * (1.0f / (F * F) - 1.0f / (C * C)));
    float a = n_d - b / (D * D);
    return a + b / (lambda_nm * lambda_nm);
}
CRT_HD float fresnel_amplitude(float eta_i, float eta_t, float cos_i, float cos_t) {  // :416-420
    float rs = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t);
    float rp = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t);
    return 0.5f * (rs + rp);
}
CRT_HD float thin_film_reflectance_lambda(float cos_theta_1, float eta_1, float eta_film, float eta_2, float thickness_nm,
                                          float lambda_nm) {  // :345-379
    float cos1 = clampf(cos_theta_1, 0.0f, 1.0f);
    float sin2_1 = 1.0f - cos1 * cos1;
    float e1f = eta_1 / eta_film;
    float sin2_film = (e1f * e1f) * sin2_1;
    if (sin2_film >= 1.0f) return 1.0f;
    float cos_film = sqrtf(1.0f - sin2_film);
    float ef2 = eta_film / eta_2;
    float sin2_base = (ef2 * ef2) * sin2_film;
    if (sin2_base >= 1.0f) return 1.0f;
    float cos_base = sqrtf(1.0f - sin2_base);
    float r_a = fresnel_amplitude(eta_1, eta_film, cos1, cos_film);
    float r_b = fresnel_amplitude(eta_film, eta_2, cos_film, cos_base);
    float opd = 2.0f * eta_film * thickness_nm * cos_film;
    float phi = 2.0f * PI_F * opd / lambda_nm;
    float cos_phi = cosf(phi);
    float num = r_a * r_a + 2.0f * r_a * r_b * cos_phi + r_b * r_b;
    float rab = r_a * r_b;
    float den = 1.0f + 2.0f * r_a * r_b * cos_phi + rab * rab;
    return clampf(num / fmaxf(den, 1e-8f), 0.0f, 1.0f);
}
CRT_HD_NOINLINE V3 thin_film_fresnel(float cos_theta_1, float eta_1, float eta_film, float eta_2, float thickness_nm) {  // :381-394
    return {thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, 615.0f),
            thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, 545.0f),
            thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, 465.0f)};
}
CRT_HD_NOINLINE V3 thin_film_fresnel_metal(float cos_theta_1, float eta_1, float eta_film, V3 f0, float thickness_nm) {  // :396-412
    float out[3];
    const float f0a[3] = {f0.x, f0.y, f0.z};
    for (int i = 0; i < 3; ++i) {
        float f0_c = clampf(f0a[i], 0.0f, 0.9999f);
        float sq = sqrtf(f0_c);
        float eta_2 = (1.0f + sq) / (1.0f - sq);
        out[i] = thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, lambda_rgb(i));
    }
    return {out[0], out[1], out[2]};
}

}  // namespace shade
