"""The host build of the product's shading headers (`csrc/shade/*.h`, compiled by g++ into the C++ oracle and by nvcc into
the device kernels) against a witness that shares no source with them: `oracle/openpbr_ref.py`, a float64 restatement
written separately from the reference's `material/openpbr.rs` and `material/brdf.rs`. Same cases and tolerances as the
device-side tests/test_gpu_shade_independent.py. A transliteration slip in `brdf.h` / `material.h` — invisible to every
GPU-vs-C++-oracle image test because both compile the same headers — fails here. Also restates the reference unit tests
of `brdf.rs` / `openpbr.rs` that are pure functions of the witness (sanity of the witness itself)."""
import math

import numpy as np
import pytest

import openpbr_ref as R
import shade_cases as S


def _host_eval(orc):
    def f(rec, ray_dir, normal, ff, wi):
        n = len(ray_dir)
        value, pdf, some = np.zeros((n, 3)), np.zeros(n), np.zeros(n, bool)
        for i in range(n):
            r = orc.material_eval(rec, ray_dir[i], normal[i], bool(ff[i]), wi[i])
            if r is not None:
                value[i], pdf[i], some[i] = r[0], r[1], True
        return value, pdf, some
    return f


def _host_scatter(orc):
    def f(rec, ray_dir, normal, ff, s4):
        n = len(ray_dir)
        d, value, pdf, delta, some = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros(n, bool), np.zeros(n, bool)
        for i in range(n):
            r = orc.material_scatter(rec, ray_dir[i], normal[i], bool(ff[i]), s4[i])
            if r is not None:
                d[i], value[i], pdf[i], delta[i], some[i] = r["dir"], r["value"], r["pdf"], r["delta"], True
        return d, value, pdf, delta, some
    return f


@pytest.mark.parametrize("name", sorted(S.MATERIALS))
def test_host_eval_matches_the_independent_restatement(orc, name):
    S.check_eval(name, _host_eval(orc), n=400)


@pytest.mark.parametrize("name", sorted(S.MATERIALS))
def test_host_scatter_matches_the_independent_restatement(orc, name):
    S.check_scatter(name, _host_scatter(orc), n=400)


def test_witness_reproduces_reference_unit_tests():
    """brdf.rs / openpbr.rs unit tests that are pure functions, evaluated on the witness: f0_from_ior_glass / _air
    (openpbr.rs:1516-1525), fresnel_dielectric_sanity (:1502-1513), eon_reduces_to_lambert_at_zero_roughness (:1738-1749),
    eon_is_reciprocal (:1751-1761), cauchy_fit_hits_abbe_definition (:1616-1630), dispersive_ior_no_scale_is_flat (:1601-1607),
    dispersive_ior_blue_bends_more (:1609-1614), coat_darkening_identity_at_zero (:1558-1563), defaults (:1547-1556)."""
    assert abs(R.f0_from_ior(1.5) - 0.04) < 1e-6 and R.f0_from_ior(1.0) == 0.0
    assert abs(R.fresnel_dielectric(1.0, 1.0, 1.5) - 0.04) < 1e-6 and R.fresnel_dielectric(0.0, 1.0, 1.5) == 1.0
    assert R.fresnel_dielectric(0.1, 1.5, 1.0) == 1.0  # total internal reflection
    v, l = (0.0, 0.6, 0.8), (0.6, 0.0, 0.8)
    assert all(abs(c - 0.5 / math.pi) < 1e-6 for c in R.eon_diffuse((0.5, 0.5, 0.5), 0.0, v, l))
    a, b = R.eon_diffuse((0.7, 0.4, 0.2), 0.8, v, l), R.eon_diffuse((0.7, 0.4, 0.2), 0.8, l, v)
    assert all(abs(x - y) < 1e-9 for x, y in zip(a, b))
    n_d, v_d = 1.5168, 64.17
    n_f, n_c = R.cauchy_ior(n_d, v_d, R.FRAUNHOFER_F_NM), R.cauchy_ior(n_d, v_d, R.FRAUNHOFER_C_NM)
    assert abs((n_d - 1.0) / (n_f - n_c) - v_d) < 1e-6 and abs(R.cauchy_ior(n_d, v_d, R.FRAUNHOFER_D_NM) - n_d) < 1e-12
    assert R.dispersive_ior(1.5, 20.0, 0.0) == (1.5, 1.5, 1.5)
    r, g, b_ = R.dispersive_ior(1.5, 20.0, 1.0)
    assert r < g < b_
    assert R.coat_darkening_factor((0.3, 0.5, 0.7), 1.6, 0.0) == (1.0, 1.0, 1.0)
    m = R.M()
    pmf = R.lobe_pmf(m)
    assert abs(sum(pmf) - 1.0) < 1e-12 and pmf[4] == 0.0 and m.specular_ior == 1.5 and m.coat_ior == 1.6 and m.base_color == (0.8, 0.8, 0.8)
