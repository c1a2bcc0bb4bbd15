"""Independent restatement of the reference's OpenPBR BSDF — TEST INFRASTRUCTURE, not product code.

Written from `crates/crust-core/src/material/openpbr.rs` and `material/brdf.rs` of the reference, in plain Python
float64, WITHOUT reading or including `crust-render_b200/csrc/shade/*.h` (the C++ transliteration that the product
and the C++ oracle both compile). It exists so that the shading math has a second, separately written witness: the
GPU evaluates `Material::eval` / `scatter_importance` through `crt_debug_material_eval` / `crt_debug_material_scatter`
and tests/test_gpu_shade_independent.py compares the numbers with this file. A transliteration slip in `brdf.h` or
`material.h` (a flipped sign, a swapped argument) shows up here even though the C++ oracle would share it.

A material is a dict with the field names of `OpenPBR` (openpbr.rs:66-128); vectors are 3-tuples / lists.
Only `tests/` may import this module.
"""
import math

PI = math.pi


# ---- tiny vector helpers (glam Vec3A semantics: component-wise) --------------------------------------------------
def v3(x, y, z):
    return (float(x), float(y), float(z))


def splat(s):
    return (s, s, s)


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def mul(a, b):
    if isinstance(b, tuple):
        return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
    return (a[0] * b, a[1] * b, a[2] * b)


def div(a, b):
    if isinstance(b, tuple):
        return (a[0] / b[0], a[1] / b[1], a[2] / b[2])
    return (a[0] / b, a[1] / b, a[2] / b)


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def neg(a):
    return (-a[0], -a[1], -a[2])


def length_squared(a):
    return dot(a, a)


def normalize(a):
    return mul(a, 1.0 / math.sqrt(dot(a, a)))


def clamp(x, lo, hi):
    return min(max(x, lo), hi)


def vclamp01(a):
    return tuple(clamp(c, 0.0, 1.0) for c in a)


def lerp(a, b, t):  # glam Vec3A::lerp: a + (b - a) * t
    return add(a, mul(sub(b, a), t))


def vpow(a, e):
    return tuple(c ** e for c in a)


# ---- material/brdf.rs ------------------------------------------------------------------------------------------------
def fresnel_schlick(cos_theta, f0):  # brdf.rs:10-12
    return add(f0, mul(sub(splat(1.0), f0), (1.0 - cos_theta) ** 5))


def fresnel_schlick_scalar(cos_theta, f0):  # :15-17
    return f0 + (1.0 - f0) * (1.0 - cos_theta) ** 5


def f0_from_ior(ior):  # :28-31
    r = (ior - 1.0) / (ior + 1.0)
    return r * r


def roughness_to_alpha_aniso(roughness, anisotropy):  # :39-45
    a = roughness * roughness
    inv = 1.0 - clamp(anisotropy, 0.0, 1.0)
    ax = a * math.sqrt(2.0 / (1.0 + inv * inv))
    ay = inv * ax
    return max(ax, 1e-4), max(ay, 1e-4)


def ggx_d_aniso(n_dot_h, h_dot_t, h_dot_b, ax, ay):  # :49-54
    tx = h_dot_t / ax
    ty = h_dot_b / ay
    term = tx * tx + ty * ty + n_dot_h * n_dot_h
    return 1.0 / (PI * ax * ay * term * term)


def ggx_lambda_aniso(v_dot_n, v_dot_t, v_dot_b, ax, ay):  # :57-63
    vt = v_dot_t * ax
    vb = v_dot_b * ay
    a2 = vt * vt + vb * vb
    n2 = max(v_dot_n * v_dot_n, 1e-8)
    return (-1.0 + math.sqrt(1.0 + a2 / n2)) * 0.5


def ggx_g2_smith_aniso(v_dot_n, v_dot_t, v_dot_b, l_dot_n, l_dot_t, l_dot_b, ax, ay):  # :67-80
    lv = ggx_lambda_aniso(v_dot_n, v_dot_t, v_dot_b, ax, ay)
    ll = ggx_lambda_aniso(l_dot_n, l_dot_t, l_dot_b, ax, ay)
    return 1.0 / (1.0 + lv + ll)


def sample_vndf_ggx_aniso_local(v_local, ax, ay, uv):  # :86-110 (Heitz 2018)
    vh = normalize((ax * v_local[0], ay * v_local[1], v_local[2]))
    lensq = vh[0] * vh[0] + vh[1] * vh[1]
    t1 = div((-vh[1], vh[0], 0.0), math.sqrt(lensq)) if lensq > 0.0 else (1.0, 0.0, 0.0)
    t2 = cross(vh, t1)
    u1, u2 = uv
    r = math.sqrt(u1)
    phi = 2.0 * PI * u2
    t1c = r * math.cos(phi)
    t2c_pre = r * math.sin(phi)
    s = 0.5 * (1.0 + vh[2])
    t2c = (1.0 - s) * math.sqrt(max(1.0 - t1c * t1c, 0.0)) + s * t2c_pre
    nh = add(add(mul(t1, t1c), mul(t2, t2c)), mul(vh, math.sqrt(max(1.0 - t1c * t1c - t2c * t2c, 0.0))))
    return normalize((ax * nh[0], ay * nh[1], max(nh[2], 0.0)))


def pdf_vndf_ggx_aniso_local(v_local, h_local, ax, ay):  # :113-124 (density of the reflected direction)
    n_dot_v = max(v_local[2], 1e-6)
    n_dot_h = max(h_local[2], 1e-6)
    d = ggx_d_aniso(n_dot_h, h_local[0], h_local[1], ax, ay)
    g1 = 1.0 / (1.0 + ggx_lambda_aniso(n_dot_v, v_local[0], v_local[1], ax, ay))
    return d * g1 / (4.0 * n_dot_v)


def pdf_vndf_h_aniso_local(v_local, h_local, ax, ay):  # :129-136 (density of the half vector)
    n_dot_v = max(v_local[2], 1e-6)
    v_dot_h = max(dot(v_local, h_local), 0.0)
    d = ggx_d_aniso(max(h_local[2], 1e-6), h_local[0], h_local[1], ax, ay)
    g1 = 1.0 / (1.0 + ggx_lambda_aniso(n_dot_v, v_local[0], v_local[1], ax, ay))
    return d * g1 * v_dot_h / n_dot_v


EON_A = 0.5 - 2.0 / (3.0 * PI)            # :148
EON_B = 2.0 / 3.0 - 28.0 / (15.0 * PI)    # :149


def eon_albedo_approx(mu, roughness):  # :168-177
    mucomp = 1.0 - clamp(mu, 0.0, 1.0)
    g1, g2, g3, g4 = 0.057108529, 0.49188187, -0.33218144, 0.071442995
    g_over_pi = mucomp * (g1 + mucomp * (g2 + mucomp * (g3 + mucomp * g4)))
    return (1.0 + roughness * g_over_pi) / (1.0 + EON_A * roughness)


def eon_diffuse(rho, roughness, v_local, l_local):  # :182-208
    rho = vclamp01(rho)
    mu_i, mu_o = v_local[2], l_local[2]
    s = dot(v_local, l_local) - mu_i * mu_o
    s_over_t = s / max(max(mu_i, mu_o), 1e-6) if s > 0.0 else s
    af = 1.0 / (1.0 + EON_A * roughness)
    f_ss = mul(mul(rho, af / PI), 1.0 + roughness * s_over_t)
    e_o = eon_albedo_approx(mu_o, roughness)
    e_i = eon_albedo_approx(mu_i, roughness)
    avg_e = af * (1.0 + EON_B * roughness)
    rho_ms = div(mul(mul(rho, rho), avg_e), sub(splat(1.0), mul(rho, 1.0 - avg_e)))
    eps = 1.0e-7
    f_ms = mul(mul(rho_ms, 1.0 / PI), max(1.0 - e_o, eps) * max(1.0 - e_i, eps) / max(1.0 - avg_e, eps))
    return add(f_ss, f_ms)


def fresnel_f82_tint(cos_theta, f0, tint):  # :217-225
    mu_bar = 1.0 / 7.0
    mu = clamp(cos_theta, 0.0, 1.0)

    def f_schlick(m):
        return add(f0, mul(sub(splat(1.0), f0), (1.0 - m) ** 5))

    denom = mu_bar * (1.0 - mu_bar) ** 6
    a = div(mul(f_schlick(mu_bar), sub(splat(1.0), tint)), denom)
    return vclamp01(sub(f_schlick(mu), mul(a, mu * (1.0 - mu) ** 6)))


def fresnel_dielectric(cos_i, eta_i, eta_t):  # :230-241
    cos_i = clamp(cos_i, 0.0, 1.0)
    sin2_t = (eta_i / eta_t) * (eta_i / eta_t) * (1.0 - cos_i * cos_i)
    if sin2_t >= 1.0:
        return 1.0
    cos_t = math.sqrt(1.0 - sin2_t)
    r_par = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t)
    r_perp = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t)
    return 0.5 * (r_par * r_par + r_perp * r_perp)


def tangent_frame(n):  # :245-252 (Duff et al. branchless ONB)
    sign = 1.0 if n[2] >= 0.0 else -1.0
    a = -1.0 / (sign + n[2])
    b = n[0] * n[1] * a
    t = (1.0 + sign * n[0] * n[0] * a, sign * b, -sign * n[0])
    bt = (b, sign + n[1] * n[1] * a, -n[1])
    return t, bt


def to_tangent(v, t, b, n):  # :255-257
    return (dot(v, t), dot(v, b), dot(v, n))


def from_tangent(vl, t, b, n):  # :260-262
    return add(add(mul(t, vl[0]), mul(b, vl[1])), mul(n, vl[2]))


def sheen_charlie(n_dot_v, n_dot_l, n_dot_h, roughness):  # :266-281
    alpha = max(roughness, 0.05)
    inv_alpha = 1.0 / alpha
    sin_theta_h_sq = max(1.0 - n_dot_h * n_dot_h, 0.0)
    d = (2.0 + inv_alpha) * sin_theta_h_sq ** (inv_alpha * 0.5) / (2.0 * PI)
    v = 1.0 / (4.0 * max(n_dot_l + n_dot_v - n_dot_l * n_dot_v, 1e-4))
    return d * v


def coat_darkening_factor(base_color, coat_ior, darkening):  # :288-294
    f_avg = f0_from_ior(coat_ior) + (1.0 - f0_from_ior(coat_ior)) * 0.05
    den = tuple(max(1.0 - f_avg * (1.0 - c), 1e-4) for c in base_color)
    dark = div(base_color, den)
    return add(mul(splat(1.0), 1.0 - darkening), mul(dark, darkening))


LAMBDA_RGB = (615.0, 545.0, 465.0)  # :300
FRAUNHOFER_C_NM, FRAUNHOFER_D_NM, FRAUNHOFER_F_NM = 656.3, 587.6, 486.1


def cauchy_ior(n_d, v_d, lambda_nm):  # :326-333
    b = (n_d - 1.0) / (v_d * (1.0 / (FRAUNHOFER_F_NM * FRAUNHOFER_F_NM) - 1.0 / (FRAUNHOFER_C_NM * FRAUNHOFER_C_NM)))
    a = n_d - b / (FRAUNHOFER_D_NM * FRAUNHOFER_D_NM)
    return a + b / (lambda_nm * lambda_nm)


def fresnel_amplitude(eta_i, eta_t, cos_i, cos_t):  # :417-421
    rs = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t)
    rp = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t)
    return 0.5 * (rs + rp)


def thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, lambda_nm):  # :337-372
    cos1 = clamp(cos_theta_1, 0.0, 1.0)
    sin2_1 = 1.0 - cos1 * cos1
    sin2_film = (eta_1 / eta_film) ** 2 * sin2_1
    if sin2_film >= 1.0:
        return 1.0
    cos_film = math.sqrt(1.0 - sin2_film)
    sin2_base = (eta_film / eta_2) ** 2 * sin2_film
    if sin2_base >= 1.0:
        return 1.0
    cos_base = math.sqrt(1.0 - sin2_base)
    r_a = fresnel_amplitude(eta_1, eta_film, cos1, cos_film)
    r_b = fresnel_amplitude(eta_film, eta_2, cos_film, cos_base)
    opd = 2.0 * eta_film * thickness_nm * cos_film
    phi = 2.0 * PI * opd / lambda_nm
    cos_phi = math.cos(phi)
    num = r_a * r_a + 2.0 * r_a * r_b * cos_phi + r_b * r_b
    den = 1.0 + 2.0 * r_a * r_b * cos_phi + (r_a * r_b) ** 2
    return clamp(num / max(den, 1e-8), 0.0, 1.0)


def thin_film_fresnel(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm):  # :375-388
    return tuple(thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, lam) for lam in LAMBDA_RGB)


def thin_film_fresnel_metal(cos_theta_1, eta_1, eta_film, f0, thickness_nm):  # :397-414
    out = []
    for i, lam in enumerate(LAMBDA_RGB):
        f0_c = clamp(f0[i], 0.0, 0.9999)
        sqrt_f0 = math.sqrt(f0_c)
        eta_2 = (1.0 + sqrt_f0) / (1.0 - sqrt_f0)
        out.append(thin_film_reflectance_lambda(cos_theta_1, eta_1, eta_film, eta_2, thickness_nm, lam))
    return tuple(out)


def cosine_hemisphere(uv):  # crates/utils/src/common.rs:128-135
    u, v = uv
    z = math.sqrt(1.0 - v)
    phi = 2.0 * PI * u
    return (math.cos(phi) * math.sqrt(v), math.sin(phi) * math.sqrt(v), z)


# ---- material/openpbr.rs ---------------------------------------------------------------------------------------------
DEFAULTS = dict(  # openpbr.rs:130-173
    base_weight=1.0, base_color=(0.8, 0.8, 0.8), base_diffuse_roughness=0.0, base_metalness=0.0, specular_weight=1.0,
    specular_color=(1.0, 1.0, 1.0), specular_roughness=0.3, specular_ior=1.5, specular_roughness_anisotropy=0.0,
    transmission_weight=0.0, transmission_color=(1.0, 1.0, 1.0), transmission_depth=0.0, transmission_scatter=(0.0, 0.0, 0.0),
    transmission_scatter_anisotropy=0.0, transmission_dispersion_scale=0.0, transmission_dispersion_abbe_number=20.0,
    subsurface_weight=0.0, subsurface_color=(0.8, 0.8, 0.8), subsurface_radius=1.0, subsurface_radius_scale=(1.0, 0.5, 0.25),
    subsurface_scatter_anisotropy=0.0, fuzz_weight=0.0, fuzz_color=(1.0, 1.0, 1.0), fuzz_roughness=0.5, coat_weight=0.0,
    coat_color=(1.0, 1.0, 1.0), coat_roughness=0.0, coat_roughness_anisotropy=0.0, coat_ior=1.6, coat_darkening=1.0,
    thin_film_weight=0.0, thin_film_thickness=0.5, thin_film_ior=1.4, emission_luminance=0.0, emission_color=(1.0, 1.0, 1.0),
    geometry_opacity=1.0, geometry_thin_walled=False,
)


class M:
    """An OpenPBR parameter set with attribute access (`OpenPBR { ..Default::default() }`)."""

    def __init__(self, **kw):
        for k, v in {**DEFAULTS, **kw}.items():
            setattr(self, k, tuple(float(c) for c in v) if isinstance(v, (tuple, list)) else v)


def luma(c):  # :402-404
    return 0.2126 * c[0] + 0.7152 * c[1] + 0.0722 * c[2]


def lobe_pmf(m):  # :317-378 -> (p_diffuse, p_specular, p_coat, p_fuzz, p_transmission)
    f0_diel = f0_from_ior(m.specular_ior)
    f0_coat = f0_from_ior(m.coat_ior)
    base_luma = max(luma(m.base_color), 0.02)
    spec_luma = max(luma(m.specular_color), 0.02)
    fuzz_luma = max(luma(m.fuzz_color), 0.02)
    w_metal = m.base_metalness * m.specular_weight * max(luma(mul(m.base_color, m.base_weight)), 0.02)
    w_diel_spec = (1.0 - m.base_metalness) * m.specular_weight * spec_luma * f0_diel
    w_specular = max(w_metal + w_diel_spec, 1e-4)
    w_diffuse = max((1.0 - m.base_metalness) * (1.0 - m.transmission_weight) * m.base_weight * base_luma * (1.0 - f0_diel), 1e-4)
    w_coat = max(m.coat_weight * f0_coat, 1e-6)
    w_fuzz = max(m.fuzz_weight * fuzz_luma, 1e-6)
    trans_luma = max(luma(m.transmission_color), 0.02)
    w_transmission = max((1.0 - m.base_metalness) * m.transmission_weight * trans_luma, 1e-4) if m.transmission_weight > 0.0 else 0.0
    total = w_diffuse + w_specular + w_coat + w_fuzz + w_transmission
    return (w_diffuse / total, w_specular / total, w_coat / total, w_fuzz / total, w_transmission / total)


DIFFUSE, SPECULAR, COAT, FUZZ, TRANSMISSION = range(5)


def lobe_pick(pmf, u):  # :380-398
    acc = pmf[0]
    if u < acc:
        return DIFFUSE
    acc += pmf[1]
    if u < acc:
        return SPECULAR
    acc += pmf[2]
    if u < acc:
        return COAT
    acc += pmf[3]
    if u < acc:
        return FUZZ
    return TRANSMISSION


def eval_diffuse(m, v_local, l_local, f_avg_diel):  # :410-446
    if l_local[2] <= 0.0 or v_local[2] <= 0.0:
        return splat(0.0)
    presence = m.base_weight * (1.0 - m.base_metalness) * (1.0 - m.transmission_weight)
    if presence <= 0.0:
        return splat(0.0)
    diffuse_color = lerp(m.base_color, m.subsurface_color, m.subsurface_weight)
    rho = mul(diffuse_color, presence)
    return mul(eon_diffuse(rho, m.base_diffuse_roughness, v_local, l_local), 1.0 - f_avg_diel)


def eval_specular(m, v_local, l_local, h_local, ax, ay):  # :448-551
    n_dot_v = max(v_local[2], 1e-4)
    n_dot_l = max(l_local[2], 1e-4)
    n_dot_h = max(h_local[2], 1e-4)
    v_dot_h = max(dot(v_local, h_local), 1e-4)
    d = ggx_d_aniso(n_dot_h, h_local[0], h_local[1], ax, ay)
    g = ggx_g2_smith_aniso(n_dot_v, v_local[0], v_local[1], n_dot_l, l_local[0], l_local[1], ax, ay)
    f0_diel_scalar = f0_from_ior(m.specular_ior)
    outer_ior = m.coat_ior if m.coat_weight > 0.0 else 1.0
    tf_thickness_nm = m.thin_film_thickness * 1000.0
    if m.base_metalness < 1.0:
        f0_diel_base = mul(mul(m.specular_color, f0_diel_scalar), m.specular_weight)
        if m.thin_film_weight > 0.0:
            f_normal = fresnel_schlick(v_dot_h, f0_diel_base)
            f_iri = thin_film_fresnel(v_dot_h, outer_ior, m.thin_film_ior, m.specular_ior, tf_thickness_nm)
            f_diel = add(mul(f_normal, 1.0 - m.thin_film_weight), mul(f_iri, m.thin_film_weight))
        else:
            f_diel = fresnel_schlick(v_dot_h, f0_diel_base)
        if m.geometry_thin_walled and m.transmission_weight > 0.0:
            f_phys = fresnel_schlick_scalar(v_dot_h, f0_diel_scalar)
            boost = 2.0 / (1.0 + f_phys)
            f_diel = mul(f_diel, 1.0 + (boost - 1.0) * m.transmission_weight)
        diel_term = mul(f_diel, 1.0 - m.base_metalness)
    else:
        diel_term = splat(0.0)
    if m.base_metalness > 0.0:
        metal_f0 = mul(m.base_color, m.base_weight)
        f_metal_base = fresnel_f82_tint(v_dot_h, metal_f0, m.specular_color)
        if m.thin_film_weight > 0.0:
            f_iri = thin_film_fresnel_metal(v_dot_h, outer_ior, m.thin_film_ior, metal_f0, tf_thickness_nm)
            f_metal = add(mul(f_metal_base, 1.0 - m.thin_film_weight), mul(f_iri, m.thin_film_weight))
        else:
            f_metal = f_metal_base
        metal_term = mul(mul(f_metal, m.specular_weight), m.base_metalness)
    else:
        metal_term = splat(0.0)
    brdf = d * g / (4.0 * n_dot_v * n_dot_l)
    return mul(add(metal_term, diel_term), brdf)


def eval_coat(m, v_local, l_local, h_local, ax_coat, ay_coat):  # :553-576
    n_dot_v = max(v_local[2], 1e-4)
    n_dot_l = max(l_local[2], 1e-4)
    n_dot_h = max(h_local[2], 1e-4)
    v_dot_h = max(dot(v_local, h_local), 1e-4)
    d = ggx_d_aniso(n_dot_h, h_local[0], h_local[1], ax_coat, ay_coat)
    g = ggx_g2_smith_aniso(n_dot_v, v_local[0], v_local[1], n_dot_l, l_local[0], l_local[1], ax_coat, ay_coat)
    f = fresnel_schlick_scalar(v_dot_h, f0_from_ior(m.coat_ior))
    brdf = d * g / (4.0 * n_dot_v * n_dot_l)
    return splat(m.coat_weight * f * brdf)


def coat_passage(m, cos_theta):  # :587-603
    cos_i = clamp(cos_theta, 1e-4, 1.0)
    eta = max(m.coat_ior, 1e-4)
    sin2_t = (1.0 - cos_i * cos_i) / (eta * eta)
    cos_t = math.sqrt(max(1.0 - sin2_t, 0.0))
    path_length = 1.0 / max(cos_t, 1e-3)
    one_passage = vpow(vclamp01(m.coat_color), 0.5 * path_length)
    absorb = lerp(splat(1.0), one_passage, m.coat_weight)
    f_coat = fresnel_schlick_scalar(cos_i, f0_from_ior(m.coat_ior))
    return mul(absorb, 1.0 - m.coat_weight * f_coat)


def coat_attenuation(m, cos_v, cos_l):  # :614-620
    if m.coat_weight <= 0.0:
        return splat(1.0)
    dark = coat_darkening_factor(m.base_color, m.coat_ior, m.coat_darkening)
    return mul(mul(coat_passage(m, cos_v), coat_passage(m, cos_l)), dark)


def eval_fuzz(m, v_local, l_local, h_local):  # :622-627
    n_dot_v = max(v_local[2], 1e-4)
    n_dot_l = max(l_local[2], 1e-4)
    n_dot_h = max(h_local[2], 0.0)
    return mul(mul(m.fuzz_color, m.fuzz_weight), sheen_charlie(n_dot_v, n_dot_l, n_dot_h, m.fuzz_roughness))


def dispersive_ior(n_d, abbe, dispersion_scale):  # :736-751
    if dispersion_scale <= 0.0 or n_d == 1.0:
        return splat(n_d)
    inverted = n_d < 1.0
    n_above_one = 1.0 / n_d if inverted else n_d
    v_d = max(max(abbe, 1.0) / dispersion_scale, 1.0)
    out = []
    for lam in LAMBDA_RGB:
        n = cauchy_ior(n_above_one, v_d, lam)
        out.append(1.0 / n if inverted else n)
    return tuple(out)


def transmission_is_continuous(m):  # :827-829
    return m.transmission_weight > 0.0 and not m.geometry_thin_walled


def transmission_iors(m):  # :833-839
    return dispersive_ior(m.specular_ior, m.transmission_dispersion_abbe_number, m.transmission_dispersion_scale)


def interface_iors(ior, entering):  # :844-846
    return (1.0, ior) if entering else (ior, 1.0)


def transmission_alphas(m):  # :850-855
    return roughness_to_alpha_aniso(max(m.specular_roughness, 0.01), m.specular_roughness_anisotropy)


def eval_transmission_channel(m, v_local, l_local, entering, ior):  # :862-917 (Walter et al. 2007)
    eta_i, eta_t = interface_iors(ior, entering)
    h = neg(add(mul(v_local, eta_i), mul(l_local, eta_t)))
    if length_squared(h) < 1e-12:
        return 0.0, 0.0
    h = normalize(h)
    if h[2] < 0.0:
        h = neg(h)
    v_dot_h = dot(v_local, h)
    l_dot_h = dot(l_local, h)
    if v_dot_h <= 1e-6 or l_dot_h >= -1e-6:
        return 0.0, 0.0
    ax, ay = transmission_alphas(m)
    n_dot_v = max(v_local[2], 1e-6)
    n_dot_l = max(-l_local[2], 1e-6)
    d = ggx_d_aniso(max(h[2], 1e-6), h[0], h[1], ax, ay)
    g = ggx_g2_smith_aniso(n_dot_v, v_local[0], v_local[1], n_dot_l, l_local[0], l_local[1], ax, ay)
    f = fresnel_dielectric(v_dot_h, eta_i, eta_t)
    denom = eta_i * v_dot_h + eta_t * l_dot_h
    denom2 = denom * denom
    if denom2 < 1e-10:
        return 0.0, 0.0
    btdf = (v_dot_h * -l_dot_h) / (n_dot_v * n_dot_l) * (eta_t * eta_t * (1.0 - f) * d * g / denom2)
    p_h = pdf_vndf_h_aniso_local(v_local, h, ax, ay)
    jacobian = eta_t * eta_t * -l_dot_h / denom2
    return max(btdf, 0.0), p_h * jacobian


def eval_transmission(m, v_local, l_local, entering):  # :924-949
    color = splat(1.0) if m.transmission_depth > 0.0 else m.transmission_color
    tint = mul(color, m.transmission_weight * (1.0 - m.base_metalness))
    iors = transmission_iors(m)
    if m.transmission_dispersion_scale <= 0.0:
        btdf, pdf = eval_transmission_channel(m, v_local, l_local, entering, iors[1])
        return mul(tint, btdf), pdf
    value, pdf = [], 0.0
    for ior in iors:
        btdf, p = eval_transmission_channel(m, v_local, l_local, entering, ior)
        value.append(btdf)
        pdf += p / 3.0
    return mul(tint, tuple(value)), pdf


def sample_transmission_rough(m, v_local, entering, dispersion_u, vndf_uv):  # :958-993
    iors = transmission_iors(m)
    if m.transmission_dispersion_scale > 0.0:
        ior = iors[0] if dispersion_u < 1.0 / 3.0 else (iors[1] if dispersion_u < 2.0 / 3.0 else iors[2])
    else:
        ior = iors[1]
    eta_i, eta_t = interface_iors(ior, entering)
    eta_rel = eta_i / eta_t
    ax, ay = transmission_alphas(m)
    h = sample_vndf_ggx_aniso_local(v_local, ax, ay, vndf_uv)
    cos_i = dot(v_local, h)
    if cos_i <= 1e-6:
        return None
    sin2_t = eta_rel * eta_rel * (1.0 - cos_i * cos_i)
    if sin2_t >= 1.0:
        return None
    cos_t = math.sqrt(1.0 - sin2_t)
    l = normalize(add(mul(neg(v_local), eta_rel), mul(h, eta_rel * cos_i - cos_t)))
    return None if l[2] >= -1e-6 else l


def eval_all(m, v_local, l_local, entering):  # :629-683
    if v_local[2] <= 0.0:
        return splat(0.0)
    if l_local[2] <= 0.0:
        if not transmission_is_continuous(m):
            return splat(0.0)
        return eval_transmission(m, v_local, l_local, entering)[0]
    h_local = normalize(add(v_local, l_local))
    ax, ay = roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy)
    f_avg_diel = f0_from_ior(m.specular_ior)
    diffuse = eval_diffuse(m, v_local, l_local, f_avg_diel)
    specular = eval_specular(m, v_local, l_local, h_local, ax, ay)
    if m.coat_weight > 0.0:
        ax_c, ay_c = roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy)
        coat = eval_coat(m, v_local, l_local, h_local, ax_c, ay_c)
    else:
        coat = splat(0.0)
    fuzz = eval_fuzz(m, v_local, l_local, h_local) if m.fuzz_weight > 0.0 else splat(0.0)
    coat_atten = coat_attenuation(m, v_local[2], l_local[2])
    base_atten = clamp(1.0 - m.fuzz_weight, 0.0, 1.0)
    return add(fuzz, mul(add(coat, mul(coat_atten, add(diffuse, specular))), base_atten))


def pdf_all(m, pmf, v_local, l_local, entering):  # :689-722
    if v_local[2] <= 0.0:
        return 0.0
    if l_local[2] <= 0.0:
        if not transmission_is_continuous(m):
            return 0.0
        return pmf[4] * eval_transmission(m, v_local, l_local, entering)[1]
    h_local = normalize(add(v_local, l_local))
    ax, ay = roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy)
    ax_c, ay_c = roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy)
    pdf_cosine = max(l_local[2], 0.0) / PI
    pdf_specular = pdf_vndf_ggx_aniso_local(v_local, h_local, ax, ay)
    pdf_coat = pdf_vndf_ggx_aniso_local(v_local, h_local, ax_c, ay_c)
    return pmf[0] * pdf_cosine + pmf[1] * pdf_specular + pmf[2] * pdf_coat + pmf[3] * pdf_cosine


def material_eval(m, ray_dir, normal, front_face, wi):
    """`OpenPBR::eval` (openpbr.rs:1138-1158): (brdf * |cos|, pdf) or None."""
    t, b = tangent_frame(normal)
    v_local = to_tangent(neg(normalize(ray_dir)), t, b, normal)
    if v_local[2] <= 0.0:
        return None
    l_local = to_tangent(normalize(wi), t, b, normal)
    pmf = lobe_pmf(m)
    pdf = max(pdf_all(m, pmf, v_local, l_local, front_face), 1e-4)
    return mul(eval_all(m, v_local, l_local, front_face), abs(l_local[2])), pdf


def material_scatter(m, ray_dir, normal, front_face, s):
    """`OpenPBR::scatter_importance` given the 4-D block `s` (openpbr.rs:1026-1136): dict(dir, value, pdf, delta) or None.
    Ray origins / medium tagging are the integrator's business and are not restated here."""
    t, b = tangent_frame(normal)
    v_local = to_tangent(neg(normalize(ray_dir)), t, b, normal)
    if v_local[2] <= 0.0:
        return None
    dir_uv = (s[1], s[2])
    pmf = lobe_pmf(m)
    lobe = lobe_pick(pmf, s[0])
    if lobe == TRANSMISSION:
        if transmission_is_continuous(m):
            l_local = sample_transmission_rough(m, v_local, front_face, s[3], dir_uv)
            if l_local is None:
                return None
            pdf = max(pdf_all(m, pmf, v_local, l_local, front_face), 1e-4)
            brdf = eval_all(m, v_local, l_local, front_face)
            return dict(dir=from_tangent(l_local, t, b, normal), value=mul(brdf, abs(l_local[2])), pdf=pdf, delta=False, lobe=lobe)
        # thin-walled delta transmission (:779-812)
        d = normalize(ray_dir)
        cos_i = clamp(dot(neg(d), normal), 0.0, 1.0)
        eta = max(m.specular_ior, 1e-4)
        sin2_t = (1.0 - cos_i * cos_i) / (eta * eta)
        if sin2_t >= 1.0:
            throughput = splat(0.0)
        else:
            cos_t = math.sqrt(1.0 - sin2_t)
            f = fresnel_dielectric(cos_i, 1.0, eta)
            window_transmittance = (1.0 - f) / (1.0 + f)
            path_length = 1.0 / max(cos_t, 1e-4)
            tint = vpow(vclamp01(m.transmission_color), path_length)
            throughput = mul(tint, window_transmittance * m.transmission_weight)
        return dict(dir=d, value=div(throughput, max(pmf[4], 1e-4)), pdf=1.0, delta=True, lobe=lobe)
    if lobe in (DIFFUSE, FUZZ):
        l_local = cosine_hemisphere(dir_uv)
    else:
        if lobe == SPECULAR:
            ax, ay = roughness_to_alpha_aniso(m.specular_roughness, m.specular_roughness_anisotropy)
        else:
            ax, ay = roughness_to_alpha_aniso(m.coat_roughness, m.coat_roughness_anisotropy)
        h_local = sample_vndf_ggx_aniso_local(v_local, ax, ay, dir_uv)
        l_local = sub(mul(h_local, 2.0 * dot(v_local, h_local)), v_local)
        if l_local[2] <= 0.0:
            return None
    pdf = max(pdf_all(m, pmf, v_local, l_local, front_face), 1e-4)
    brdf = eval_all(m, v_local, l_local, front_face)
    return dict(dir=from_tangent(l_local, t, b, normal), value=mul(brdf, max(l_local[2], 0.0)), pdf=pdf, delta=False, lobe=lobe)


def emitted_directional(m, cos_theta_o):  # :1211-1218
    uncoated = mul(m.emission_color, m.emission_luminance)
    if m.coat_weight <= 0.0:
        return uncoated
    dark = coat_darkening_factor(m.base_color, m.coat_ior, m.coat_darkening)
    return mul(mul(uncoated, coat_passage(m, cos_theta_o)), dark)
