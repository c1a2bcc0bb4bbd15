"""Known-answer / property tests of crust-core's shading code, restated against the host compile of
crust-render_b200/csrc/shade/*.h (through the oracle's C hooks). Each test names the reference test
it follows: crates/crust-core/src/material/openpbr.rs:1236-2246, light.rs:448-639,
filter.rs:262-412, tracer.rs:1616-1670, crates/utils/src/common.rs. These pin the leaf math the CUDA
integrator and the CPU oracle share; the sampler values themselves are unpinned (openqmc-rs is not
vendored), so sample-driven tests use this repo's sampler exactly like the reference uses its own."""
import math

import numpy as np
import pytest

from crust_render_b200.render_types import DistantLight, DomeLight, Emissive, OpenPBR, RectLight, SphereLight

PI = math.pi


def n3(v):
    v = np.asarray(v, np.float32)
    return v / np.float32(np.sqrt((v * v).sum(dtype=np.float32)))


def glass(ior, **kw):  # OpenPBR::glass, openpbr.rs:197-204
    return OpenPBR(**{**dict(transmission_weight=1.0, specular_ior=ior, specular_roughness=0.01), **kw})


def blocks(orc, n, start=0):
    """`S::next()` of the reference tests (openpbr.rs:1247-1256): a fresh sampler per draw."""
    return [orc.sampler_draw(0, 0, 0, start + i + 1)[0] for i in range(n)]


R_IN = n3((-0.3, 0.2, -1.0))
Z = (0.0, 0.0, 1.0)


def refract_dir(v, n, eta):
    v, n = np.asarray(v, np.float64), np.asarray(n, np.float64)
    cos_i = float(v @ n)
    sin2_t = eta * eta * (1.0 - cos_i * cos_i)
    assert sin2_t < 1.0
    return (-v * eta + n * (eta * cos_i - math.sqrt(1.0 - sin2_t))).astype(np.float32)


# ---- openpbr.rs --------------------------------------------------------------------------------
def _eval_consistency(orc, m, n, tol_pdf=1e-3, max_w=10.0, check_value=True):
    stats = dict(checked=0, transmitted=0, reflected=0)
    for s4 in blocks(orc, n):
        smp = orc.material_scatter(m, R_IN, Z, True, s4)
        if smp is None:
            continue
        assert not smp["delta"]
        wi = n3(smp["dir"])
        stats["transmitted" if wi[2] < 0 else "reflected"] += 1
        ev = orc.material_eval(m, R_IN, Z, True, wi)
        assert ev is not None
        value, pdf = ev
        if check_value:
            assert np.abs(value - smp["value"]).max() < 1e-3 * (1.0 + abs(smp["value"].max()))
        assert abs(pdf - smp["pdf"]) < tol_pdf * (1.0 + smp["pdf"])
        w = (smp["value"] / smp["pdf"]).max()
        assert np.isfinite(w) and 0.0 <= w < max_w
        stats["checked"] += 1
    return stats


def test_eval_matches_scatter_importance(orc):  # openpbr.rs:1259-1285
    assert _eval_consistency(orc, OpenPBR(), 128)["checked"] > 32


def test_glass_transmission_is_continuous_and_eval_consistent(orc):  # :1288-1355
    st = _eval_consistency(orc, glass(1.5, specular_roughness=0.25), 256)
    assert st["transmitted"] > 64
    _eval_consistency(orc, glass(1.5), 128, tol_pdf=0.05, check_value=False)


def test_near_smooth_refraction_matches_snell(orc):  # :1358-1384
    m = glass(1.5)
    dir_in = n3((0.4, 0.0, -1.0))
    expected = n3(refract_dir(-dir_in, Z, 1.0 / 1.5))
    checked = 0
    for s4 in blocks(orc, 128):
        smp = orc.material_scatter(m, dir_in, Z, True, s4)
        if smp is not None and n3(smp["dir"])[2] < 0:
            assert float(n3(smp["dir"]) @ expected) > 0.995
            checked += 1
    assert checked > 32


def test_thin_walled_transmission_stays_delta(orc):  # :1387-1410
    m = glass(1.5, geometry_thin_walled=1)
    deltas = 0
    for s4 in blocks(orc, 128):
        smp = orc.material_scatter(m, R_IN, Z, True, s4)
        if smp is not None and smp["dir"][2] < 0:
            assert smp["delta"]
            deltas += 1
    assert deltas > 16
    value, _ = orc.material_eval(m, R_IN, Z, True, (0, 0, -1))
    assert np.all(value == 0.0)


def _thin_throughput(orc, m, d):
    """sample_transmission_thin (openpbr.rs:779-812) = the delta sample's value times the lobe-selection probability."""
    smp = orc.material_scatter(m, d, Z, True, (0.9999, 0.5, 0.5, 0.5))
    assert smp is not None and smp["delta"]
    return smp["value"] * max(orc.material_pmf(m)[4], 1e-4)


def test_thin_wall_window_transmittance_matches_formula(orc):  # :1413-1446
    m = glass(1.5, geometry_thin_walled=1)
    d = n3((0.6, 0.0, -1.0))
    thr = _thin_throughput(orc, m, d)
    f = orc.shade_fn("fresnel_dielectric", [-d[2], 1.0, 1.5])[0]
    expected = (1.0 - f) / (1.0 + f)
    assert np.abs(thr - expected).max() < 1e-5
    assert abs(2.0 * f / (1.0 + f) + expected - 1.0) < 1e-6
    assert _thin_throughput(orc, m, n3((6.0, 0.0, -1.0)))[0] < _thin_throughput(orc, m, (0, 0, -1))[0]


def test_thin_wall_tint_darkens_with_angle(orc):  # :1449-1477
    m = glass(1.5, geometry_thin_walled=1, transmission_color=(0.4, 0.8, 0.9))
    tn = _thin_throughput(orc, m, (0, 0, -1))
    assert abs(tn[0] / tn[2] - 0.4 / 0.9) < 1e-4
    to = _thin_throughput(orc, m, n3((2.0, 0.0, -1.0)))
    assert to[0] / to[2] < tn[0] / tn[2]


def test_thin_wall_reflection_boosted_by_internal_bounces(orc):  # :1480-1499
    thick = glass(1.5, specular_roughness=0.25)
    thin = glass(1.5, specular_roughness=0.25, geometry_thin_walled=1)
    v, l = n3((0.4, 0.0, 1.0)), n3((-0.4, 0.0, 1.0))
    assert orc.shade_fn("eval_all", [*v, *l, 1], thin)[0] > orc.shade_fn("eval_all", [*v, *l, 1], thick)[0] * 1.5


def test_fresnel_dielectric_sanity(orc):  # :1502-1513
    assert abs(orc.shade_fn("fresnel_dielectric", [1.0, 1.0, 1.5])[0] - 0.04) < 1e-3
    assert orc.shade_fn("fresnel_dielectric", [0.2, 1.5, 1.0])[0] == 1.0
    assert orc.shade_fn("fresnel_dielectric", [0.01, 1.0, 1.5])[0] > 0.9


def test_f0_from_ior(orc):  # :1516-1524
    assert abs(orc.shade_fn("f0_from_ior", [1.5])[0] - 0.04) < 1e-3
    assert abs(orc.shade_fn("f0_from_ior", [1.0])[0]) < 1e-6


def test_iso_matches_aniso_at_zero_and_reference_remap(orc):  # :1527-1532, 1961-1969
    ax, ay = orc.shade_fn("roughness_to_alpha_aniso", [0.4, 0.0])
    assert abs(ax - 0.16) < 1e-5 and abs(ay - 0.16) < 1e-5
    r, a = 0.5, 0.8
    ax, ay = orc.shade_fn("roughness_to_alpha_aniso", [r, a])
    expect = r * r * math.sqrt(2.0 / (1.0 + (1 - a) ** 2))
    assert abs(ax - expect) < 1e-6 and abs(ay - (1 - a) * expect) < 1e-6


def test_sheen_nonneg(orc):  # :1535-1544
    for nv in (0.1, 0.3, 0.5, 0.9):
        for nl in (0.1, 0.3, 0.5, 0.9):
            for nh in (0.1, 0.5, 0.9):
                v = orc.shade_fn("sheen_charlie", [nv, nl, nh, 0.4])[0]
                assert np.isfinite(v) and v >= 0.0


def test_defaults_match_spec(orc):  # :1547-1555 — the C default record equals the Python mirror
    from crust_render_b200.render_types import MATERIAL_DTYPE
    rec = np.zeros((), dtype=MATERIAL_DTYPE)
    orc.lib().orc_material_default(rec.ctypes.data)
    assert rec.tobytes() == OpenPBR().tobytes()
    assert rec["base_weight"] == 1.0 and rec["specular_ior"] == 1.5 and rec["coat_ior"] == np.float32(1.6)
    assert rec["thin_film_ior"] == np.float32(1.4) and rec["transmission_dispersion_abbe_number"] == 20.0
    assert tuple(rec["subsurface_radius_scale"]) == (1.0, 0.5, 0.25)


def test_coat_darkening_identity_at_zero(orc):  # :1558-1562
    assert np.linalg.norm(orc.shade_fn("coat_darkening_factor", [0.7, 0.3, 0.2, 1.6, 0.0]) - 1.0) < 1e-4


def test_thin_film_bounded(orc):  # :1565-1570, 1972-1982
    r = orc.shade_fn("thin_film_fresnel", [1.0, 1.0, 1.4, 1.5, 500.0])
    assert np.all((r >= 0) & (r <= 1))
    f0 = (0.9, 0.7, 0.4)
    r = orc.shade_fn("thin_film_fresnel_metal", [0.8, 1.0, 1.4, *f0, 500.0])
    assert np.all((r >= 0) & (r <= 1))
    plain = orc.shade_fn("fresnel_f82_tint", [0.8, *f0, 1, 1, 1])
    assert np.abs(r - plain).max() > 1e-3


def test_coat_lobe_contributes_when_enabled(orc):  # :1573-1598
    m = OpenPBR(coat_weight=1.0, coat_roughness=0.05, coat_ior=1.5, base_color=(0.5, 0.5, 0.5))
    d = n3((-0.5, 0.0, -1.0))
    assert any((s := orc.material_scatter(m, d, Z, False, s4)) is not None and (s["value"] ** 2).sum() > 0 for s4 in blocks(orc, 128))


def test_dispersive_ior(orc):  # :1601-1660
    assert tuple(orc.shade_fn("dispersive_ior", [1.5, 30.0, 0.0])) == (1.5, 1.5, 1.5)
    v = orc.shade_fn("dispersive_ior", [1.5, 30.0, 1.0])
    assert v[2] > v[1] > v[0]
    n_d, v_d = 1.5, 30.0
    assert abs(orc.shade_fn("cauchy_ior", [n_d, v_d, 587.6])[0] - n_d) < 1e-6
    abbe = (n_d - 1.0) / (orc.shade_fn("cauchy_ior", [n_d, v_d, 486.1])[0] - orc.shade_fn("cauchy_ior", [n_d, v_d, 656.3])[0])
    assert abs(abbe - v_d) < 0.05
    full, half = orc.shade_fn("dispersive_ior", [1.5, 40.0, 1.0]), orc.shade_fn("dispersive_ior", [1.5, 40.0, 0.5])
    assert abs((full[2] - full[0]) / (half[2] - half[0]) - 2.0) < 1e-3 and abs(full[1] - 1.5) < 0.01
    n, inv = orc.shade_fn("dispersive_ior", [1.5, 30.0, 1.0]), orc.shade_fn("dispersive_ior", [1.0 / 1.5, 30.0, 1.0])
    assert np.abs(inv - 1.0 / n).max() < 1e-5 and inv[2] < inv[1] < inv[0]


def test_dispersive_glass_is_continuous_and_eval_consistent(orc):  # :1663-1708
    st = _eval_consistency(orc, glass(1.5, specular_roughness=0.25, transmission_dispersion_scale=1.0), 256, max_w=30.0)
    assert st["transmitted"] > 64


def test_dispersion_separates_channels(orc):  # :1711-1735
    m = glass(1.5, transmission_dispersion_scale=1.0)
    d = n3((0.6, 0.0, -1.0))
    eta_g = orc.shade_fn("transmission_iors", [0], m)[1]
    l_green = n3(refract_dir(-d, Z, 1.0 / eta_g))
    v, pdf = orc.material_eval(m, d, Z, True, l_green)
    assert pdf > 0 and v[1] > 0 and v[1] > v[0] and v[1] > v[2]


def test_eon(orc):  # :1738-1805
    rho = (0.8, 0.5, 0.3)
    f = orc.shade_fn("eon_diffuse", [*rho, 0.0, *n3((0.3, 0.1, 0.9)), *n3((-0.2, 0.4, 0.8))])
    assert np.abs(f - np.asarray(rho) / PI).max() < 1e-4
    v, l = n3((0.5, -0.1, 0.6)), n3((-0.3, 0.2, 0.9))
    for r in (0.2, 0.6, 1.0):
        a = orc.shade_fn("eon_diffuse", [0.9, 0.6, 0.2, r, *v, *l])
        b = orc.shade_fn("eon_diffuse", [0.9, 0.6, 0.2, r, *l, *v])
        assert np.abs(a - b).max() < 1e-5
    for i in range(1, 21):
        for r in (0.0, 0.3, 0.7, 1.0):
            assert abs(orc.shade_fn("eon_albedo_exact", [i / 20.0, r])[0] - orc.shade_fn("eon_albedo_approx", [i / 20.0, r])[0]) < 0.01


def test_eon_preserves_energy_at_high_roughness(orc):  # :1778-1805 (coarser quadrature: 48 x 48)
    nt = nphi = 48
    for view_z in (0.95, 0.6, 0.25):
        v = (math.sqrt(1 - view_z * view_z), 0.0, view_z)
        total = 0.0
        for it in range(nt):
            th = (it + 0.5) / nt * (PI / 2)
            for ip in range(nphi):
                ph = (ip + 0.5) / nphi * 2 * PI
                l = (math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th))
                total += orc.shade_fn("eon_diffuse", [1, 1, 1, 1.0, *v, *l])[0] * math.cos(th) * math.sin(th)
        total *= (PI / 2) / nt * (2 * PI) / nphi
        assert 0.97 <= total <= 1.03


def test_deep_transmission_interface_is_untinted(orc):  # :1808-1844
    color = (0.9, 0.4, 0.2)
    shallow = glass(1.5, specular_roughness=0.25, transmission_color=color)
    deep = glass(1.5, specular_roughness=0.25, transmission_color=color, transmission_depth=1.0)
    d = n3((0.4, 0.0, -1.0))
    wi = n3(refract_dir(-d, Z, 1.0 / 1.5))
    vd, _ = orc.material_eval(deep, d, Z, True, wi)
    assert vd[1] > 0 and abs(vd[0] - vd[1]) < 1e-5 and abs(vd[1] - vd[2]) < 1e-5
    vs, _ = orc.material_eval(shallow, d, Z, True, wi)
    assert abs(vs[0] / vs[1] - color[0] / color[1]) < 1e-3


def test_f82_metal_edge_tint(orc):  # :1847-1863
    f0, tint = (0.9, 0.6, 0.3), (1.0, 0.5, 0.25)
    assert np.abs(orc.shade_fn("fresnel_f82_tint", [1.0, *f0, *tint]) - f0).max() < 1e-5
    assert np.abs(orc.shade_fn("fresnel_f82_tint", [0.0, *f0, *tint]) - 1.0).max() < 1e-5
    w = orc.shade_fn("fresnel_f82_tint", [1 / 7, *f0, *tint])
    wo = orc.shade_fn("fresnel_f82_tint", [1 / 7, *f0, 1, 1, 1])
    assert abs(w[1] / wo[1] - tint[1]) < 1e-3 and abs(w[2] / wo[2] - tint[2]) < 1e-3 and abs(w[0] - wo[0]) < 1e-5


def test_coat_color_and_passages(orc):  # :1866-1930
    m = OpenPBR(coat_weight=1.0, coat_color=(0.9, 0.2, 0.2), coat_darkening=0.0)
    v = n3((0.3, 0.0, 1.0))
    c = orc.shade_fn("eval_coat", [*v, *v, *v, 0.1, 0.1], m)
    assert abs(c[0] - c[1]) < 1e-6 and abs(c[1] - c[2]) < 1e-6
    at = orc.shade_fn("coat_attenuation", [1.0, 1.0], m)
    assert abs(at[0] / at[1] - 0.9 / 0.2) < 1e-3
    m2 = OpenPBR(coat_weight=1.0, coat_color=(0.9, 0.4, 0.16), coat_darkening=0.0)
    f0 = orc.shade_fn("f0_from_ior", [1.6])[0]
    assert np.abs(orc.shade_fn("coat_attenuation", [1.0, 1.0], m2) - np.asarray((0.9, 0.4, 0.16)) * (1 - f0) ** 2).max() < 1e-4
    m3 = OpenPBR(coat_weight=1.0, coat_color=(0.9, 0.3, 0.3))
    pn, pg = orc.shade_fn("coat_passage", [1.0], m3), orc.shade_fn("coat_passage", [0.2], m3)
    assert pg[0] < pn[0] and pg[1] < pn[1] and pg[1] / pg[0] < pn[1] / pn[0]
    assert orc.shade_fn("coat_attenuation", [0.2, 0.2], m3)[1] < orc.shade_fn("coat_attenuation", [1.0, 1.0], m3)[1]


def test_coated_emission_dims_and_tints(orc):  # :1933-1958
    un = OpenPBR(emission_luminance=100.0)
    assert np.array_equal(orc.material_emitted_directional(un, 0.3), orc.shade_fn("material_emitted", [0], un))
    co = OpenPBR(emission_luminance=100.0, coat_weight=1.0, coat_color=(1.0, 0.2, 0.2))
    en = orc.material_emitted_directional(co, 1.0)
    assert en[0] < 100.0 and abs(en[1] / en[0] - math.sqrt(0.2)) < 1e-3
    assert orc.material_emitted_directional(co, 0.05)[0] < en[0]


def _interior_medium(orc, m):  # sample_interior_medium, :1984-2000
    for s4 in blocks(orc, 256):
        smp = orc.material_scatter(m, R_IN, Z, True, s4)
        if smp is not None and smp["dir"][2] < 0:
            return orc.material_medium(m) if smp["enters_medium"] else None
    raise AssertionError("material never transmitted")


def test_transmission_scatter_wires_into_interior_medium(orc):  # :2003-2031
    m = glass(1.5, specular_roughness=0.25, transmission_color=(0.5, 0.7, 0.9), transmission_depth=2.0, transmission_scatter=(0.2, 0.2, 0.1),
              transmission_scatter_anisotropy=0.5)
    sa, ss, g = _interior_medium(orc, m)
    assert ss.max() > 1e-6
    assert np.abs(ss - np.asarray((0.2, 0.2, 0.1)) / 2).max() < 1e-5
    assert np.abs(sa + ss - (-np.log(np.asarray((0.5, 0.7, 0.9), np.float32)) / 2)).max() < 1e-5
    assert g == 0.5


def test_medium_constructors(orc):  # :2034-2066, 2116-2131
    md = orc.shade_fn("medium_from_transmission", [1, 1, 1, 1.0, 0.1, 0.2, 0.4, 0.0])
    sa = md[0:3]
    assert sa.min() >= 0 and abs(sa[2]) < 1e-6 and sa[0] > sa[1] > sa[2]
    albedo = lambda md: md[3:6] / (md[0:3] + md[3:6])  # noqa: E731  Medium::albedo
    m = orc.shade_fn("medium_from_subsurface", [0.5, 0.5, 0.5, 1.0, 1, 1, 1, 0.0])
    a = albedo(m)
    assert abs(a[0] - 0.9117) < 5e-3
    hi = albedo(orc.shade_fn("medium_from_subsurface", [0.95] * 3 + [1.0, 1, 1, 1, 0.0]))
    lo = albedo(orc.shade_fn("medium_from_subsurface", [0.05] * 3 + [1.0, 1, 1, 1, 0.0]))
    assert hi[0] > 0.99 and lo[0] < a[0] < hi[0]
    assert albedo(orc.shade_fn("medium_from_subsurface", [0.5] * 3 + [1.0, 1, 1, 1, 0.9]))[0] > a[0]
    assert abs((m[0:3] + m[3:6])[0] - 1.0) < 1e-5
    z = orc.shade_fn("medium_from_transmission", [0.5, 0.5, 0.5, 0.0, 0, 0, 0, 0.0])
    assert np.linalg.norm(orc.shade_fn("medium_transmittance", [*z, 1.0]) - 1.0) < 1e-4
    md = orc.shade_fn("medium_from_transmission", [0.5, 0.7, 0.9, 1.0, 0, 0, 0, 0.0])
    assert np.all(orc.shade_fn("medium_transmittance", [*md, 2.0]) < orc.shade_fn("medium_transmittance", [*md, 0.1]))


def test_interior_medium_blends_subsurface_into_glass(orc):  # :2069-2099
    m = glass(1.5, specular_roughness=0.25, transmission_depth=1.0, subsurface_weight=1.0, subsurface_color=(0.5, 0.5, 0.5), subsurface_radius=0.1,
              subsurface_radius_scale=(1, 1, 1), subsurface_scatter_anisotropy=0.3, transmission_weight=0.5)
    sa, ss, g = _interior_medium(orc, m)
    assert ss.max() > 1e-6 and abs(g - 0.3) < 1e-5
    full = orc.shade_fn("medium_from_subsurface", [0.5, 0.5, 0.5, 0.1, 1, 1, 1, 0.3])
    assert np.abs(ss - full[3:6] * 0.5).max() < 1e-3


def test_inert_glass_attaches_no_medium(orc):  # :2102-2113
    assert _interior_medium(orc, glass(1.5, specular_roughness=0.25)) is None


def test_subsurface_shifts_diffuse_color(orc):  # :2134-2171
    a = OpenPBR(base_color=(0.9, 0.9, 0.9), subsurface_color=(0.9, 0.1, 0.1), subsurface_weight=1.0, base_diffuse_roughness=0.5)
    b = OpenPBR(base_color=(0.9, 0.9, 0.9), base_diffuse_roughness=0.5)
    sa, sb = np.zeros(3), np.zeros(3)
    bl = blocks(orc, 1024)
    for k in range(512):
        s = orc.material_scatter(a, (0, 0, -1), Z, True, bl[2 * k])
        sa += s["value"] if s else 0
        s = orc.material_scatter(b, (0, 0, -1), Z, True, bl[2 * k + 1])
        sb += s["value"] if s else 0
    assert sa[1] < sb[1] and sa[2] < sb[2]


def test_hg_isotropic_at_g_zero(orc):  # :2174-2188
    total = np.zeros(3)
    for i in range(2048):
        total += orc.shade_fn("sample_henyey_greenstein", [0, 0, 1, 0.0, ((i * 13 + 7) % 1024) / 1024.0, ((i * 31 + 5) % 1024) / 1024.0])
    assert np.linalg.norm(total / 2048) < 0.05


def test_thin_walled_transmission_scatters_downward(orc):  # :2191-2216
    m = OpenPBR(transmission_weight=1.0, transmission_color=(0.7, 0.9, 0.7), geometry_thin_walled=1)
    got = False
    for s4 in blocks(orc, 64):
        s = orc.material_scatter(m, (0, -1, 0), (0, 1, 0), True, s4)
        if s is not None and s["dir"][1] < 0:
            assert s["delta"]
            got = True
            break
    assert got


def test_scatter_importance_finite(orc):  # :2219-2245
    m = OpenPBR(base_color=(0.7, 0.3, 0.2), base_metalness=0.3, fuzz_weight=0.2, fuzz_color=(0.9, 0.9, 0.9))
    for s4 in blocks(orc, 64):
        s = orc.material_scatter(m, n3((0, -1, -1)), (0, 1, 0), False, s4)
        if s is not None:
            assert np.isfinite(s["pdf"]) and s["pdf"] > 0 and np.all(np.isfinite(s["value"])) and np.all(s["value"] >= 0)


def test_emissive_neither_scatters_nor_evaluates(orc):  # emissive.rs:26-39, material.rs:64-69
    e = Emissive((3, 2, 1))
    assert orc.material_scatter(e, R_IN, Z, True, (0.5, 0.5, 0.5, 0.5)) is None
    assert orc.material_eval(e, R_IN, Z, True, Z) is None
    assert tuple(orc.material_emitted_directional(e, 0.3)) == (3.0, 2.0, 1.0)


# ---- light.rs:448-639 ----------------------------------------------------------------------------
def test_sphere_shape_samples_lie_on_surface(orc):
    light = SphereLight((1, 2, 3), 0.5, (1, 1, 1), 0)
    origin = np.zeros(3, np.float32)
    for u, v in ((0.0, 0.0), (0.25, 0.75), (0.99, 0.5), (0.5, 0.01)):
        s = orc.light_sample_li(light, origin, u, v)
        p = origin + s["direction"] * s["distance"]
        assert abs(np.linalg.norm(p - np.asarray((1, 2, 3))) - 0.5) < 1e-5


def test_rect_shape(orc):
    light = RectLight((-1, 5, -2), (2, 0, 0), (0, 0, 4), (0, -2, 0), (1, 1, 1), 3)
    assert tuple(light["n"]) == (0.0, -1.0, 0.0)
    s = orc.light_sample_li(light, (0, 0, 0), 0.5, 0.5)
    assert np.linalg.norm(s["direction"] * s["distance"] - np.asarray((0, 5, 0))) < 1e-5
    # area 8 => pdf = d^2 / (cos * 8 + 1e-4) straight below the centre
    assert abs(s["pdf"] - 25.0 / (8.0 + 1e-4)) < 1e-4


def test_area_light_pdf_is_positive_facing_side(orc):
    light = SphereLight((0, 5, 0), 1.0, (10, 10, 10), 0)
    pdf = orc.light_pdf_at_point(light, (0, 0, 0), (0, 4, 0))
    assert np.isfinite(pdf) and pdf > 0
    s = orc.light_sample_li(light, (0, 0, 0), 0.3, 0.7)
    assert abs(np.linalg.norm(s["direction"]) - 1) < 1e-5 and 0 < s["distance"] < np.inf
    assert tuple(s["radiance"]) == (10.0, 10.0, 10.0) and s["pdf"] > 0
    fp = orc.light_pdf_at_point(light, (0, 0, 0), s["direction"] * s["distance"])
    assert abs(s["pdf"] - fp) <= 1e-3 * max(s["pdf"], fp)
    assert orc.light_escaped(light, (0, 1, 0)) is None


def test_distant_light_cone_is_consistent(orc):
    d = n3((0.3, -1.0, 0.2))
    light = DistantLight(d, (2, 2, 2), 10.0)
    rng = np.random.default_rng(7)
    for _ in range(500):
        s = orc.light_sample_li(light, (0, 0, 0), float(rng.random()), float(rng.random()))
        assert abs(np.linalg.norm(s["direction"]) - 1) < 1e-5 and np.isinf(s["distance"])
        esc = orc.light_escaped(light, s["direction"])
        assert esc is not None
        assert np.array_equal(esc[0], s["radiance"]) and abs(esc[1] - s["pdf"]) < 1e-3 * s["pdf"]
    assert orc.light_escaped(light, d) is None
    outside = n3(orc.shade_fn("align_to_normal", [math.sin(math.radians(20)), 0, math.cos(math.radians(20)), *(-d)]))
    assert orc.light_escaped(light, outside) is None


def test_distant_light_irradiance_is_angle_invariant(orc):
    e = np.asarray((3, 2, 1), np.float32)
    for angle in (0.0, 0.53, 5.0, 30.0):
        s = orc.light_sample_li(DistantLight((0, -1, 0), e, angle), (0, 0, 0), 0.4, 0.6)
        assert np.linalg.norm(s["radiance"] / s["pdf"] - e) < 1e-3 * np.linalg.norm(e)


def test_distant_light_zero_angle_stays_finite(orc):
    light = DistantLight((0, -1, 0), (1, 1, 1), 0.0)
    s = orc.light_sample_li(light, (0, 0, 0), 0.5, 0.5)
    assert np.isfinite(s["pdf"]) and s["pdf"] > 0 and np.all(np.isfinite(s["radiance"]))
    off = n3(orc.shade_fn("align_to_normal", [math.sin(math.radians(1)), 0, math.cos(math.radians(1)), 0, 1, 0]))
    assert orc.light_escaped(light, off) is None
    assert orc.light_pdf_at_point(light, (0, 0, 0), (0, 1, 0)) == 0.0


def test_dome_light_uniform(orc):  # light.rs:320-389 without a map
    light = DomeLight((0.5, 0.6, 0.7))
    s = orc.light_sample_li(light, (0, 0, 0), 0.25, 0.5)
    assert abs(s["pdf"] - 1 / (4 * PI)) < 1e-7 and np.isinf(s["distance"]) and abs(np.linalg.norm(s["direction"]) - 1) < 1e-5
    r, pdf = orc.light_escaped(light, s["direction"])
    assert np.allclose(r, (0.5, 0.6, 0.7)) and abs(pdf - s["pdf"]) < 1e-7


def test_light_pick(orc):  # LightList::pick, light.rs:421-430
    assert orc.light_pick(0, 0.5) == -1
    assert [orc.light_pick(4, u) for u in (0.0, 0.249, 0.25, 0.999, 1.0)] == [0, 0, 1, 3, 3]


# ---- filter.rs:262-412 ---------------------------------------------------------------------------
FILTERS = [(0, 0.5), (1, 1.0), (2, 1.5), (3, 1.5), (4, 2.0)]


def test_box_half_radius_is_the_identity_jitter(orc):
    u = np.random.default_rng(7).random(10000, dtype=np.float32)
    x, w = orc.filter_sample_n(0, 0.5, u)
    assert np.array_equal(x.view(np.uint32), u.view(np.uint32)) and np.all(w == 1.0)


def test_samples_stay_in_support_and_monotone(orc):
    u = np.minimum(np.arange(1001, dtype=np.float32) / np.float32(1000.0), np.float32(0.999999))
    for kind, r in FILTERS:
        x, _ = orc.filter_sample_n(kind, r, u)
        assert np.all((x >= 0.5 - r - 1e-3) & (x <= 0.5 + r + 1e-3)) and np.all(np.diff(x) >= 0)


def test_weights_average_to_one(orc):
    for kind, r in FILTERS:
        u = np.random.default_rng(11).random(200000, dtype=np.float32)
        _, w = orc.filter_sample_n(kind, r, u)
        assert abs(w.astype(np.float64).mean() - 1.0) < 5e-3
        assert bool((w < 0).any()) == (kind == 4)


def test_weighted_histogram_reproduces_the_kernel(orc):
    H = 16
    for kind, r in FILTERS:
        u = np.random.default_rng(23).random(400000, dtype=np.float32)
        x, w = orc.filter_sample_n(kind, r, u)
        b = np.minimum(((x - np.float32(0.5) + np.float32(r)) / np.float32(2 * r) * H).astype(np.int64), H - 1)
        hist = np.bincount(b, weights=w.astype(np.float64), minlength=H) / len(u)
        mid = np.asarray([orc.filter_eval(kind, r, -r + (k + 0.5) * (2 * r / H)) for k in range(H)], np.float64)
        assert np.abs(hist - mid / mid.sum()).max() < 0.01


# ---- tracer.rs:1616-1670 -------------------------------------------------------------------------
def test_strategy_weights_partition_unity(orc):
    for s in range(4):
        for lp, bp in ((0.5, 0.5), (1e-4, 1e4), (1e4, 1e-4), (3.0, 0.2), (0.05, 40.0)):
            assert abs(orc.mis_weight(s, True, lp, bp) + orc.mis_weight(s, False, bp, lp) - 1.0) < 1e-3


def test_single_strategy_modes_and_power_sharpens_balance(orc):
    assert orc.mis_weight(2, True, 1.0, 100.0) == 1.0 and orc.mis_weight(2, False, 100.0, 1.0) == 0.0
    assert orc.mis_weight(3, True, 100.0, 1.0) == 0.0 and orc.mis_weight(3, False, 1.0, 100.0) == 1.0
    assert orc.mis_weight(0, True, 10.0, 1.0) > orc.mis_weight(1, True, 10.0, 1.0)


# ---- utils/src/common.rs + camera.rs ---------------------------------------------------------------
def test_warps_and_heuristics(orc):
    assert abs(orc.shade_fn("power_heuristic", [1.0, 1.0])[0] - 0.5) < 1e-6
    assert abs(orc.shade_fn("balance_heuristic", [3.0, 1.0])[0] - 0.75) < 1e-6
    for u, v in ((0.1, 0.2), (0.9, 0.5), (0.5, 0.99)):
        d = orc.shade_fn("cosine_hemisphere", [u, v])
        assert abs(np.linalg.norm(d) - 1) < 1e-5 and d[2] >= 0
        k = orc.shade_fn("concentric_disk", [u, v])
        assert np.linalg.norm(k[:2]) <= 1 + 1e-6 and k[2] == 0
    w = orc.shade_fn("align_to_normal", [0, 0, 1, *n3((1, 2, 3))])
    assert np.abs(w - n3((1, 2, 3))).max() < 1e-6


def test_camera_centre_ray_points_at_lookat(orc):  # camera.rs:27-84
    from crust_render_b200.render_types import Camera
    cam = Camera((0, 1, 5), (0, 1, 0), (0, 1, 0), 40.0, 16 / 9, 0.0, 5.0)
    o, d = orc.camera_ray(cam, 0.5, 0.5)
    assert np.allclose(o, (0, 1, 5)) and np.allclose(n3(d), (0, 0, -1), atol=1e-6)
    o2, d2 = orc.camera_ray(cam, 0.0, 0.0)
    half_h = 5.0 * math.tan(math.radians(20.0))
    assert np.allclose(d2, (-half_h * 16 / 9, -half_h, -5.0), rtol=1e-5)
    lens = Camera((0, 1, 5), (0, 1, 0), (0, 1, 0), 40.0, 1.0, 0.5, 5.0)
    o3, d3 = orc.camera_ray(lens, 0.5, 0.5, 0.9, 0.5)
    assert np.linalg.norm(o3 - np.asarray((0, 1, 5))) <= 0.25 + 1e-6 and np.allclose(o3 + d3, (0, 1, 0), atol=1e-5)  # in-focus point is fixed


# ---- environment.rs:206-405 ----------------------------------------------------------------------------
def spotty_map(orc, w, h, bx, by):
    px = np.full((h, w, 3), 0.05, np.float32)
    px[by, bx] = 500.0
    return orc.EnvironmentMap(w, h, px)


def test_env_direction_and_uv_round_trip_and_conventions(orc):
    E = orc.EnvironmentMap
    for d in ((0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1), n3((0.3, 0.5, -0.8)), n3((-0.6, -0.2, 0.7))):
        u, v = E.direction_to_uv(d)
        assert np.linalg.norm(E.uv_to_direction(u, v) - np.asarray(d, np.float32)) < 1e-4
    u, v = E.direction_to_uv((0, 0, -1))
    assert abs(u - 0.5) < 1e-5 and abs(v - 0.5) < 1e-5
    assert E.direction_to_uv((0, 1, 0))[1] < 1e-5 and E.direction_to_uv((0, -1, 0))[1] > 1 - 1e-5


def test_env_sample_and_pdf_agree(orc):
    m = spotty_map(orc, 64, 32, 40, 8)
    rng = np.random.default_rng(11)
    for _ in range(3000):
        s = m.sample(float(rng.random()), float(rng.random()))
        if s is None:
            continue
        q = m.pdf(s[0])
        assert abs(s[2] - q) <= 1e-3 * max(s[2], q)


def test_env_pdf_is_normalized_and_concentrates_on_bright_texels(orc):
    m = spotty_map(orc, 32, 16, 20, 6)
    rng = np.random.default_rng(5)
    total, n, hits, checked = 0.0, 200000, 0, 0
    for k in range(n):
        s = m.sample(float(rng.random()), float(rng.random()))
        if s is not None and s[2] > 0:
            total += 1.0 / s[2]
            if k < 20000:
                u, v = orc.EnvironmentMap.direction_to_uv(s[0])
                hits += int(min(int(u * 32), 31) == 20 and min(int(v * 16), 15) == 6)
                checked += 1
    assert abs(total / n - 4 * PI) < 0.05 * 4 * PI
    assert hits / checked > 0.5


def test_env_uniform_map_has_uniform_solid_angle_pdf_and_poles_are_not_oversampled(orc):
    w, h = 32, 16
    m = orc.EnvironmentMap(w, h, np.ones((h, w, 3), np.float32))
    for y in range(h):
        d = orc.EnvironmentMap.uv_to_direction(0.5 / w, (y + 0.5) / h)
        assert abs(m.pdf(d) - 1 / (4 * PI)) < 0.01 / (4 * PI)
    m2 = orc.EnvironmentMap(64, 32, np.ones((32, 64, 3), np.float32))
    at = lambda y: m2.pdf(orc.EnvironmentMap.uv_to_direction(0.5 / 64, (y + 0.5) / 32))  # noqa: E731
    assert abs(at(0) - at(16)) < 0.02 * at(16)


def test_env_black_map_declines_to_sample(orc):
    m = orc.EnvironmentMap(8, 4, np.zeros((4, 8, 3), np.float32))
    assert m.sample(0.5, 0.5) is None and m.pdf((0, 1, 0)) == 0.0
    # rejects_malformed_buffers (:385-389)
    assert not orc.EnvironmentMap(0, 4, []).valid and not orc.EnvironmentMap(4, 4, np.ones(9, np.float32)).valid
