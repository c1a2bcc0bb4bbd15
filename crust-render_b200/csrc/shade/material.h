// OpenPBR surface + emissive materials and the carried interior medium.
// Restates crates/crust-core/src/material/{openpbr.rs, emissive.rs, material.rs}
// and crates/crust-core/src/medium.rs (line numbers in comments name the file).
// Host + device; no textures (Ptex is out of scope: SURVEY §2 row 26).
#pragma once
#include "brdf.h"

namespace shade {

// ---- medium.rs -------------------------------------------------------------------------
struct Medium {
    V3 sigma_a, sigma_s;
    float g;
};
CRT_HD float medium_sigma_t_max(const Medium& m) { return max_element(m.sigma_a + m.sigma_s); }  // medium.rs:131-133
CRT_HD bool medium_is_scattering(const Medium& m) { return max_element(m.sigma_s) > 1e-6f; }     // :125-127
CRT_HD V3 medium_transmittance(const Medium& m, float t) {                                      // :117-120
    V3 e = (m.sigma_a + m.sigma_s) * t;
    return {expf(-e.x), expf(-e.y), expf(-e.z)};
}
CRT_HD Medium medium_from_transmission(V3 tint, float depth, V3 scatter, float anisotropy) {  // :39-65
    if (depth <= 1e-6f) return Medium{splat(0), splat(0), 0.0f};
    V3 t = vmin(vmax(tint, splat(1e-4f)), splat(1.0f));
    V3 extinction = v3(-logf(t.x), -logf(t.y), -logf(t.z)) / depth;
    V3 sigma_s = vmax(scatter, splat(0.0f)) / depth;
    V3 sigma_a = extinction - sigma_s;
    float mn = min_element(sigma_a);
    if (mn < 0.0f) sigma_a = sigma_a - splat(mn);
    return Medium{sigma_a, sigma_s, clampf(anisotropy, -0.999f, 0.999f)};
}
CRT_HD Medium medium_from_subsurface(V3 albedo, float radius, V3 radius_scale, float g) {  // :76-96
    V3 mfp = vmax(splat(radius) * radius_scale, splat(1e-3f));
    V3 sigma_t = splat(1.0f) / mfp;
    g = clampf(g, -0.999f, 0.999f);
    V3 a = clamp01(albedo);
    V3 inner = splat(9.59217f) + 41.6808f * a + 17.7126f * a * a;
    V3 s = splat(4.09712f) + 4.20863f * a - v3(sqrtf(inner.x), sqrtf(inner.y), sqrtf(inner.z));
    V3 s2 = s * s;
    V3 alpha_ss = clamp01((splat(1.0f) - s2) / (splat(1.0f) - g * s2));
    V3 sigma_s = sigma_t * alpha_ss;
    return Medium{sigma_t - sigma_s, sigma_s, g};
}
CRT_HD Medium medium_blend(const Medium& a, float wa, const Medium& b, float wb) {  // :102-114
    Medium m;
    m.sigma_a = a.sigma_a * wa + b.sigma_a * wb;
    m.sigma_s = a.sigma_s * wa + b.sigma_s * wb;
    float sa = ((a.sigma_s.x + a.sigma_s.y + a.sigma_s.z) / 3.0f) * wa, sb = ((b.sigma_s.x + b.sigma_s.y + b.sigma_s.z) / 3.0f) * wb;
    m.g = sa + sb > 1e-8f ? (a.g * sa + b.g * sb) / (sa + sb) : 0.0f;
    return m;
}
CRT_HD_NOINLINE V3 sample_henyey_greenstein(V3 wi, float g, float u1, float u2) {  // :155-179
    float cos_theta;
    if (fabsf(g) < 1e-3f) {
        cos_theta = 1.0f - 2.0f * u1;
    } else {
        float sq = (1.0f - g * g) / (1.0f - g + 2.0f * g * u1);
        cos_theta = (1.0f + g * g - sq * sq) / (2.0f * g);
    }
    cos_theta = clampf(cos_theta, -1.0f, 1.0f);
    float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
    float phi = 2.0f * PI_F * u2;
    V3 up = fabsf(wi.z) < 0.999f ? v3(0, 0, 1) : v3(1, 0, 0);
    V3 t = normalize(cross(wi, up));
    V3 b = cross(wi, t);
    return normalize(t * (sin_theta * cosf(phi)) + b * (sin_theta * sinf(phi)) + wi * cos_theta);
}

// ---- openpbr.rs ----------------------------------------------------------------------------
enum : uint32_t { MAT_OPENPBR = 0, MAT_EMISSIVE = 1 };

struct LobePmf {  // openpbr.rs:305-311
    float p_diffuse, p_specular, p_coat, p_fuzz, p_transmission;
};

// One record per material (openpbr.rs:66-128 fields, spec defaults :130-173), plus the values
// `prepare_material` derives once on the host: the lobe PMF (:317-378 depends only on parameters)
// and the interior medium (:225-259).
struct Material {
    uint32_t kind;
    float base_weight;
    V3 base_color;
    float base_diffuse_roughness, base_metalness;
    float specular_weight;
    V3 specular_color;
    float specular_roughness, specular_ior, specular_roughness_anisotropy;
    float transmission_weight;
    V3 transmission_color;
    float transmission_depth;
    V3 transmission_scatter;
    float transmission_scatter_anisotropy, transmission_dispersion_scale, transmission_dispersion_abbe_number;
    float subsurface_weight;
    V3 subsurface_color;
    float subsurface_radius;
    V3 subsurface_radius_scale;
    float subsurface_scatter_anisotropy;
    float fuzz_weight;
    V3 fuzz_color;
    float fuzz_roughness;
    float coat_weight;
    V3 coat_color;
    float coat_roughness, coat_roughness_anisotropy, coat_ior, coat_darkening;
    float thin_film_weight, thin_film_thickness, thin_film_ior;
    float emission_luminance;
    V3 emission_color;
    float geometry_opacity;
    uint32_t geometry_thin_walled;
    V3 emissive_color;  // MAT_EMISSIVE (emissive.rs:12-14)
    // derived
    LobePmf pmf;
    uint32_t has_medium;
    Medium medium;
};

CRT_HD Material material_default() {  // openpbr.rs:130-173
    Material m{};
    m.kind = MAT_OPENPBR;
    m.base_weight = 1.0f;
    m.base_color = v3(0.8f, 0.8f, 0.8f);
    m.base_diffuse_roughness = 0.0f;
    m.base_metalness = 0.0f;
    m.specular_weight = 1.0f;
    m.specular_color = splat(1.0f);
    m.specular_roughness = 0.3f;
    m.specular_ior = 1.5f;
    m.specular_roughness_anisotropy = 0.0f;
    m.transmission_weight = 0.0f;
    m.transmission_color = splat(1.0f);
    m.transmission_depth = 0.0f;
    m.transmission_scatter = splat(0.0f);
    m.transmission_scatter_anisotropy = 0.0f;
    m.transmission_dispersion_scale = 0.0f;
    m.transmission_dispersion_abbe_number = 20.0f;
    m.subsurface_weight = 0.0f;
    m.subsurface_color = v3(0.8f, 0.8f, 0.8f);
    m.subsurface_radius = 1.0f;
    m.subsurface_radius_scale = v3(1.0f, 0.5f, 0.25f);
    m.subsurface_scatter_anisotropy = 0.0f;
    m.fuzz_weight = 0.0f;
    m.fuzz_color = splat(1.0f);
    m.fuzz_roughness = 0.5f;
    m.coat_weight = 0.0f;
    m.coat_color = splat(1.0f);
    m.coat_roughness = 0.0f;
    m.coat_roughness_anisotropy = 0.0f;
    m.coat_ior = 1.6f;
    m.coat_darkening = 1.0f;
    m.thin_film_weight = 0.0f;
    m.thin_film_thickness = 0.5f;
    m.thin_film_ior = 1.4f;
    m.emission_luminance = 0.0f;
    m.emission_color = splat(1.0f);
    m.geometry_opacity = 1.0f;
    m.geometry_thin_walled = 0;
    m.emissive_color = splat(0.0f);
    return m;
}

CRT_HD float luma(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }  // openpbr.rs:402-404

CRT_HD LobePmf lobe_pmf_from_params(const Material& m) {  // openpbr.rs:317-378
    float f0_diel = f0_from_ior(m.specular_ior), f0_coat = f0_from_ior(m.coat_ior);
    float base_luma = fmaxf(luma(m.base_color), 0.02f), spec_luma = fmaxf(luma(m.specular_color), 0.02f),
          fuzz_luma = fmaxf(luma(m.fuzz_color), 0.02f);
    float w_metal = m.base_metalness * m.specular_weight * fmaxf(luma(m.base_color * m.base_weight), 0.02f);
    float w_diel_spec = (1.0f - m.base_metalness) * m.specular_weight * spec_luma * f0_diel;
    float w_specular = fmaxf(w_metal + w_diel_spec, 1e-4f);
    float w_diffuse = fmaxf((1.0f - m.base_metalness) * (1.0f - m.transmission_weight) * m.base_weight * base_luma * (1.0f - f0_diel), 1e-4f);
    float w_coat = fmaxf(m.coat_weight * f0_coat, 1e-6f);
    float w_fuzz = fmaxf(m.fuzz_weight * fuzz_luma, 1e-6f);
    float trans_luma = fmaxf(luma(m.transmission_color), 0.02f);
    float w_transmission = m.transmission_weight > 0.0f ? fmaxf((1.0f - m.base_metalness) * m.transmission_weight * trans_luma, 1e-4f) : 0.0f;
    float total = w_diffuse + w_specular + w_coat + w_fuzz + w_transmission;
    return LobePmf{w_diffuse / total, w_specular / total, w_coat / total, w_fuzz / total, w_transmission / total};
}
enum Lobe : int { LOBE_DIFFUSE, LOBE_SPECULAR, LOBE_COAT, LOBE_FUZZ, LOBE_TRANSMISSION };
CRT_HD int lobe_pick(const LobePmf& p, float u) {  // openpbr.rs:380-398
    float acc = p.p_diffuse;
    if (u < acc) return LOBE_DIFFUSE;
    acc += p.p_specular;
    if (u < acc) return LOBE_SPECULAR;
    acc += p.p_coat;
    if (u < acc) return LOBE_COAT;
    acc += p.p_fuzz;
    if (u < acc) return LOBE_FUZZ;
    return LOBE_TRANSMISSION;
}

// Fills the derived fields (host, at material-table build time).
CRT_HD void prepare_material(Material& m) {
    m.pmf = lobe_pmf_from_params(m);
    m.has_medium = 0;
    m.medium = Medium{splat(0), splat(0), 0.0f};
    if (m.kind != MAT_OPENPBR) return;
    // interior_medium, openpbr.rs:225-259
    float trans_frac = m.transmission_weight;
    float sss_frac = (1.0f - m.transmission_weight) * m.subsurface_weight;
    float total = trans_frac + sss_frac;
    if (total <= 0.0f) return;
    Medium tv = medium_from_transmission(m.transmission_color, m.transmission_depth, m.transmission_scatter, m.transmission_scatter_anisotropy);
    Medium med = tv;
    if (sss_frac > 0.0f) {
        Medium sv = medium_from_subsurface(m.subsurface_color, m.subsurface_radius, m.subsurface_radius_scale, m.subsurface_scatter_anisotropy);
        med = medium_blend(tv, trans_frac / total, sv, sss_frac / total);
    }
    if (medium_sigma_t_max(med) <= 1e-6f) return;
    m.has_medium = 1;
    m.medium = med;
}

CRT_HD V3 eval_diffuse(const Material& m, V3 v_local, V3 l_local, float f_avg_diel) {  // openpbr.rs:410-446
    if (l_local.z <= 0.0f || v_local.z <= 0.0f) return splat(0.0f);
    float presence = m.base_weight * (1.0f - m.base_metalness) * (1.0f - m.transmission_weight);
    if (presence <= 0.0f) return splat(0.0f);
    V3 diffuse_color = lerp(m.base_color, m.subsurface_color, m.subsurface_weight);
    V3 rho = diffuse_color * presence;
    return eon_diffuse(rho, m.base_diffuse_roughness, v_local, l_local) * (1.0f - f_avg_diel);
}

CRT_HD_NOINLINE V3 eval_specular(const Material& m, V3 v_local, V3 l_local, V3 h_local, float ax, float ay) {  // openpbr.rs:448-551
    float n_dot_v = fmaxf(v_local.z, 1e-4f), n_dot_l = fmaxf(l_local.z, 1e-4f), n_dot_h = fmaxf(h_local.z, 1e-4f);
    float v_dot_h = fmaxf(dot(v_local, h_local), 1e-4f);
    float d = ggx_d_aniso(n_dot_h, h_local.x, h_local.y, ax, ay);
    float g = ggx_g2_smith_aniso(n_dot_v, v_local.x, v_local.y, n_dot_l, l_local.x, l_local.y, ax, ay);
    float f0_diel_scalar = f0_from_ior(m.specular_ior);
    float outer_ior = m.coat_weight > 0.0f ? m.coat_ior : 1.0f;
    float tf_thickness_nm = m.thin_film_thickness * 1000.0f;
    V3 diel_term = splat(0.0f);
    if (m.base_metalness < 1.0f) {
        V3 f0_diel_base = m.specular_color * f0_diel_scalar * m.specular_weight;
        V3 f_diel;
        if (m.thin_film_weight > 0.0f) {
            V3 f_normal = fresnel_schlick(v_dot_h, f0_diel_base);
            V3 f_iri = thin_film_fresnel(v_dot_h, outer_ior, m.thin_film_ior, m.specular_ior, tf_thickness_nm);
            f_diel = f_normal * (1.0f - m.thin_film_weight) + f_iri * m.thin_film_weight;
        } else {
            f_diel = fresnel_schlick(v_dot_h, f0_diel_base);
        }
        if (m.geometry_thin_walled && m.transmission_weight > 0.0f) {
            float f_phys = fresnel_schlick_scalar(v_dot_h, f0_diel_scalar);
            float boost = 2.0f / (1.0f + f_phys);
            f_diel = f_diel * (1.0f + (boost - 1.0f) * m.transmission_weight);
        }
        diel_term = f_diel * (1.0f - m.base_metalness);
    }
    V3 metal_term = splat(0.0f);
    if (m.base_metalness > 0.0f) {
        V3 metal_f0 = m.base_color * m.base_weight;
        V3 f_metal_base = fresnel_f82_tint(v_dot_h, metal_f0, m.specular_color);
        V3 f_metal = f_metal_base;
        if (m.thin_film_weight > 0.0f) {
            V3 f_iri = thin_film_fresnel_metal(v_dot_h, outer_ior, m.thin_film_ior, metal_f0, tf_thickness_nm);
            f_metal = f_metal_base * (1.0f - m.thin_film_weight) + f_iri * m.thin_film_weight;
        }
        f_metal = f_metal * m.specular_weight;
        metal_term = f_metal * m.base_metalness;
    }
    float brdf = d * g / (4.0f * n_dot_v * n_dot_l);
    return (metal_term + diel_term) * brdf;
}

CRT_HD_NOINLINE V3 eval_coat(const Material& m, V3 v_local, V3 l_local, V3 h_local, float ax, float ay) {  // openpbr.rs:553-576
    float n_dot_v = fmaxf(v_local.z, 1e-4f), n_dot_l = fmaxf(l_local.z, 1e-4f), n_dot_h = fmaxf(h_local.z, 1e-4f);
    float v_dot_h = fmaxf(dot(v_local, h_local), 1e-4f);
    float d = ggx_d_aniso(n_dot_h, h_local.x, h_local.y, ax, ay);
    float g = ggx_g2_smith_aniso(n_dot_v, v_local.x, v_local.y, n_dot_l, l_local.x, l_local.y, ax, ay);
    float f = fresnel_schlick_scalar(v_dot_h, f0_from_ior(m.coat_ior));
    float brdf = d * g / (4.0f * n_dot_v * n_dot_l);
    return splat(m.coat_weight * f * brdf);
}
CRT_HD_NOINLINE V3 coat_passage(const Material& m, float cos_theta) {  // openpbr.rs:587-603
    float cos_i = clampf(cos_theta, 1e-4f, 1.0f);
    float eta = fmaxf(m.coat_ior, 1e-4f);
    float sin2_t = (1.0f - cos_i * cos_i) / (eta * eta);
    float cos_t = sqrtf(fmaxf(1.0f - sin2_t, 0.0f));
    float path_length = 1.0f / fmaxf(cos_t, 1e-3f);
    V3 one_passage = vpowf(clamp01(m.coat_color), 0.5f * path_length);
    V3 absorb = lerp(splat(1.0f), one_passage, m.coat_weight);
    float f_coat = fresnel_schlick_scalar(cos_i, f0_from_ior(m.coat_ior));
    return absorb * (1.0f - m.coat_weight * f_coat);
}
CRT_HD V3 coat_attenuation(const Material& m, float cos_v, float cos_l) {  // openpbr.rs:614-620
    if (m.coat_weight <= 0.0f) return splat(1.0f);
    V3 dark = coat_darkening_factor(m.base_color, m.coat_ior, m.coat_darkening);
    return coat_passage(m, cos_v) * coat_passage(m, cos_l) * dark;
}
CRT_HD_NOINLINE V3 eval_fuzz(const Material& m, V3 v_local, V3 l_local, V3 h_local) {  // openpbr.rs:622-627
    float n_dot_v = fmaxf(v_local.z, 1e-4f), n_dot_l = fmaxf(l_local.z, 1e-4f), n_dot_h = fmaxf(h_local.z, 0.0f);
    return m.fuzz_color * m.fuzz_weight * sheen_charlie(n_dot_v, n_dot_l, n_dot_h, m.fuzz_roughness);
}

// ---- transmission (openpbr.rs:724-993) ---------------------------------------------------------
CRT_HD V3 dispersive_ior(float n_d, float abbe, float dispersion_scale) {  // :736-751
    if (dispersion_scale <= 0.0f || n_d == 1.0f) return splat(n_d);
    bool inverted = n_d < 1.0f;
    float n_above_one = inverted ? 1.0f / n_d : n_d;
    float v_d = fmaxf(fmaxf(abbe, 1.0f) / dispersion_scale, 1.0f);
    float out[3];
    for (int c = 0; c < 3; ++c) {
        float n = cauchy_ior(n_above_one, v_d, lambda_rgb(c));
        out[c] = inverted ? 1.0f / n : n;
    }
    return {out[0], out[1], out[2]};
}
CRT_HD bool transmission_is_continuous(const Material& m) { return m.transmission_weight > 0.0f && !m.geometry_thin_walled; }  // :827-829
CRT_HD V3 transmission_iors(const Material& m) {  // :833-839
    return dispersive_ior(m.specular_ior, m.transmission_dispersion_abbe_number, m.transmission_dispersion_scale);
}
CRT_HD void transmission_alphas(const Material& m, float& ax, float& ay) {  // :850-855
    roughness_to_alpha_aniso(fmaxf(m.specular_roughness, 0.01f), m.specular_roughness_anisotropy, ax, ay);
}
CRT_HD void eval_transmission_channel(const Material& m, V3 v_local, V3 l_local, bool entering, float ior, float& btdf_out,
                                      float& pdf_out) {  // :862-914
    btdf_out = 0.0f;
    pdf_out = 0.0f;
    float eta_i = entering ? 1.0f : ior, eta_t = entering ? ior : 1.0f;  // :844-846
    V3 h = -(v_local * eta_i + l_local * eta_t);
    if (length_squared(h) < 1e-12f) return;
    h = normalize(h);
    if (h.z < 0.0f) h = -h;
    float v_dot_h = dot(v_local, h), l_dot_h = dot(l_local, h);
    if (v_dot_h <= 1e-6f || l_dot_h >= -1e-6f) return;
    float ax, ay;
    transmission_alphas(m, ax, ay);
    float n_dot_v = fmaxf(v_local.z, 1e-6f), n_dot_l = fmaxf(-l_local.z, 1e-6f);
    float d = ggx_d_aniso(fmaxf(h.z, 1e-6f), h.x, h.y, ax, ay);
    float g = ggx_g2_smith_aniso(n_dot_v, v_local.x, v_local.y, n_dot_l, l_local.x, l_local.y, ax, ay);
    float f = fresnel_dielectric(v_dot_h, eta_i, eta_t);
    float denom = eta_i * v_dot_h + eta_t * l_dot_h;
    float denom2 = denom * denom;
    if (denom2 < 1e-10f) return;
    float btdf = (v_dot_h * -l_dot_h) / (n_dot_v * n_dot_l) * (eta_t * eta_t * (1.0f - f) * d * g / denom2);
    float p_h = pdf_vndf_h_aniso_local(v_local, h, ax, ay);
    float jacobian = eta_t * eta_t * -l_dot_h / denom2;
    btdf_out = fmaxf(btdf, 0.0f);
    pdf_out = p_h * jacobian;
}
CRT_HD_NOINLINE void eval_transmission(const Material& m, V3 v_local, V3 l_local, bool entering, V3& value, float& pdf) {  // :924-949
    V3 color = m.transmission_depth > 0.0f ? splat(1.0f) : m.transmission_color;
    V3 tint = color * (m.transmission_weight * (1.0f - m.base_metalness));
    V3 iors = transmission_iors(m);
    if (m.transmission_dispersion_scale <= 0.0f) {
        float b, p;
        eval_transmission_channel(m, v_local, l_local, entering, iors.y, b, p);
        value = tint * b;
        pdf = p;
        return;
    }
    float val[3];
    const float io[3] = {iors.x, iors.y, iors.z};
    pdf = 0.0f;
    for (int c = 0; c < 3; ++c) {
        float b, p;
        eval_transmission_channel(m, v_local, l_local, entering, io[c], b, p);
        val[c] = b;
        pdf += p / 3.0f;
    }
    value = tint * v3(val[0], val[1], val[2]);
}
CRT_HD bool sample_transmission_rough(const Material& m, V3 v_local, bool entering, float dispersion_u, float u1, float u2, V3& l_out) {  // :958-993
    V3 iors = transmission_iors(m);
    float ior = iors.y;
    if (m.transmission_dispersion_scale > 0.0f) ior = dispersion_u < 1.0f / 3.0f ? iors.x : (dispersion_u < 2.0f / 3.0f ? iors.y : iors.z);
    float eta_i = entering ? 1.0f : ior, eta_t = entering ? ior : 1.0f;
    float eta_rel = eta_i / eta_t;
    float ax, ay;
    transmission_alphas(m, ax, ay);
    V3 h = sample_vndf_ggx_aniso_local(v_local, ax, ay, u1, u2);
    float cos_i = dot(v_local, h);
    if (cos_i <= 1e-6f) return false;
    float sin2_t = eta_rel * eta_rel * (1.0f - cos_i * cos_i);
    if (sin2_t >= 1.0f) return false;
    float cos_t = sqrtf(1.0f - sin2_t);
    V3 l = normalize(-v_local * eta_rel + h * (eta_rel * cos_i - cos_t));
    if (l.z >= -1e-6f) return false;
    l_out = l;
    return true;
}

CRT_HD_NOINLINE V3 eval_all(const Material& m, V3 v_local, V3 l_local, bool entering) {  // openpbr.rs:629-683
    if (v_local.z <= 0.0f) return splat(0.0f);
    if (l_local.z <= 0.0f) {
        if (!transmission_is_continuous(m)) return splat(0.0f);
        V3 val;
        float pdf;
        eval_transmission(m, v_local, l_local, entering, val, pdf);
        return val;
    }
    V3 h_local = normalize(v_local + l_local);
    float ax, ay;
    roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy, ax, ay);
    float f_avg_diel = f0_from_ior(m.specular_ior);
    V3 diffuse = eval_diffuse(m, v_local, l_local, f_avg_diel);
    V3 specular = eval_specular(m, v_local, l_local, h_local, ax, ay);
    V3 coat = splat(0.0f), fuzz = splat(0.0f);
    if (m.coat_weight > 0.0f) {
        float axc, ayc;
        roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy, axc, ayc);
        coat = eval_coat(m, v_local, l_local, h_local, axc, ayc);
    }
    if (m.fuzz_weight > 0.0f) fuzz = eval_fuzz(m, v_local, l_local, h_local);
    V3 coat_atten = coat_attenuation(m, v_local.z, l_local.z);
    float base_atten = clampf(1.0f - m.fuzz_weight, 0.0f, 1.0f);
    return fuzz + base_atten * (coat + coat_atten * (diffuse + specular));
}
CRT_HD_NOINLINE float pdf_all(const Material& m, V3 v_local, V3 l_local, bool entering) {  // openpbr.rs:689-722
    const LobePmf& pmf = m.pmf;
    if (v_local.z <= 0.0f) return 0.0f;
    if (l_local.z <= 0.0f) {
        if (!transmission_is_continuous(m)) return 0.0f;
        V3 val;
        float pdf;
        eval_transmission(m, v_local, l_local, entering, val, pdf);
        return pmf.p_transmission * pdf;
    }
    V3 h_local = normalize(v_local + l_local);
    float ax, ay, axc, ayc;
    roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy, ax, ay);
    roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy, axc, ayc);
    float pdf_cosine = fmaxf(l_local.z, 0.0f) / PI_F;
    float pdf_specular = pdf_vndf_ggx_aniso_local(v_local, h_local, ax, ay);
    float pdf_coat = pdf_vndf_ggx_aniso_local(v_local, h_local, axc, ayc);
    return pmf.p_diffuse * pdf_cosine + pmf.p_specular * pdf_specular + pmf.p_coat * pdf_coat + pmf.p_fuzz * pdf_cosine;
}

// ---- Material trait surface (material.rs, openpbr.rs:1026-1219, emissive.rs) --------------------
struct SurfaceHit {  // hittable.rs HitRecord without the Ptex fields
    V3 p, normal;
    float t;
    bool front_face;
};
struct ScatterSample {  // material.rs:8-22 (ray = origin + direction + whether it enters the interior medium)
    V3 origin, dir, value;
    float pdf;
    bool delta, enters_medium;
};

// scatter_importance: `s` is the 4-D block drawn from the BSDF domain (openpbr.rs:1042).
CRT_HD_NOINLINE bool material_scatter(const Material& m, V3 ray_dir, const SurfaceHit& rec, const float s[4], ScatterSample& out) {
    if (m.kind != MAT_OPENPBR) return false;  // emissive surfaces do not scatter (emissive.rs:31-38)
    V3 t, b;
    tangent_frame(rec.normal, t, b);
    V3 v_world = -normalize(ray_dir);
    V3 v_local = to_tangent(v_world, t, b, rec.normal);
    if (v_local.z <= 0.0f) return false;
    int lobe = lobe_pick(m.pmf, s[0]);
    out.enters_medium = false;
    if (lobe == LOBE_TRANSMISSION) {
        if (transmission_is_continuous(m)) {  // :1055-1073
            V3 l_local;
            if (!sample_transmission_rough(m, v_local, rec.front_face, s[3], s[1], s[2], l_local)) return false;
            V3 l_world = from_tangent(l_local, t, b, rec.normal);
            out.pdf = fmaxf(pdf_all(m, v_local, l_local, rec.front_face), 1e-4f);
            V3 brdf = eval_all(m, v_local, l_local, rec.front_face);
            out.origin = rec.p + l_world * 1e-4f;
            out.dir = l_world;
            out.enters_medium = rec.front_face && m.has_medium;
            out.value = brdf * fabsf(l_local.z);
            out.delta = false;
            return true;
        }
        // thin-walled delta transmission (:779-812, 1077-1086)
        V3 dir = normalize(ray_dir);
        float cos_i = clampf(dot(-dir, rec.normal), 0.0f, 1.0f);
        float eta = fmaxf(m.specular_ior, 1e-4f);
        float sin2_t = (1.0f - cos_i * cos_i) / (eta * eta);
        V3 throughput = splat(0.0f);
        if (sin2_t < 1.0f) {
            float cos_t = sqrtf(1.0f - sin2_t);
            float f = fresnel_dielectric(cos_i, 1.0f, eta);
            float window_transmittance = (1.0f - f) / (1.0f + f);
            float path_length = 1.0f / fmaxf(cos_t, 1e-4f);
            V3 tint = vpowf(clamp01(m.transmission_color), path_length);
            throughput = tint * (window_transmittance * m.transmission_weight);
        }
        float p_select = fmaxf(m.pmf.p_transmission, 1e-4f);
        out.origin = rec.p;
        out.dir = dir;
        out.value = throughput / p_select;
        out.pdf = 1.0f;
        out.delta = true;
        return true;
    }
    V3 l_local;
    if (lobe == LOBE_DIFFUSE || lobe == LOBE_FUZZ) {
        l_local = cosine_hemisphere(s[1], s[2]);
    } else {
        float ax, ay;
        if (lobe == LOBE_SPECULAR)
            roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy, ax, ay);
        else
            roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy, ax, ay);
        V3 h_local = sample_vndf_ggx_aniso_local(v_local, ax, ay, s[1], s[2]);
        l_local = 2.0f * dot(v_local, h_local) * h_local - v_local;
        if (l_local.z <= 0.0f) return false;
    }
    out.pdf = fmaxf(pdf_all(m, v_local, l_local, rec.front_face), 1e-4f);
    V3 brdf = eval_all(m, v_local, l_local, rec.front_face);
    float n_dot_l = fmaxf(l_local.z, 0.0f);
    out.origin = rec.p;
    out.dir = from_tangent(l_local, t, b, rec.normal);
    out.value = brdf * n_dot_l;
    out.delta = false;
    return true;
}

// eval: Some((brdf*|cos|, pdf)) or None (openpbr.rs:1138-1158; emissive: None)
CRT_HD_NOINLINE bool material_eval(const Material& m, V3 ray_dir, const SurfaceHit& rec, V3 wi, V3& value, float& pdf) {
    if (m.kind != MAT_OPENPBR) return false;
    V3 t, b;
    tangent_frame(rec.normal, t, b);
    V3 v_local = to_tangent(-normalize(ray_dir), t, b, rec.normal);
    if (v_local.z <= 0.0f) return false;
    V3 l_local = to_tangent(normalize(wi), t, b, rec.normal);
    pdf = fmaxf(pdf_all(m, v_local, l_local, rec.front_face), 1e-4f);
    value = eval_all(m, v_local, l_local, rec.front_face) * fabsf(l_local.z);
    return true;
}
// "is eval available" without a direction (the Material::eval contract, material.rs:64-69):
CRT_HD bool material_can_eval(const Material& m, V3 ray_dir, const SurfaceHit& rec) {
    if (m.kind != MAT_OPENPBR) return false;
    V3 t, b;
    tangent_frame(rec.normal, t, b);
    return to_tangent(-normalize(ray_dir), t, b, rec.normal).z > 0.0f;
}
CRT_HD V3 material_emitted(const Material& m) {  // openpbr.rs:1202-1204, emissive.rs:27-29
    return m.kind == MAT_EMISSIVE ? m.emissive_color : m.emission_color * m.emission_luminance;
}
CRT_HD V3 material_emitted_directional(const Material& m, float cos_theta_o) {  // openpbr.rs:1211-1218
    if (m.kind == MAT_EMISSIVE) return m.emissive_color;
    V3 uncoated = m.emission_color * m.emission_luminance;
    if (m.coat_weight <= 0.0f) return uncoated;
    V3 dark = coat_darkening_factor(m.base_color, m.coat_ior, m.coat_darkening);
    return uncoated * coat_passage(m, cos_theta_o) * dark;
}

}  // namespace shade
