"""The unit tests of crates/crust-core/src/volume.rs:542-800 (and medium.rs:186-245 for the phase function),
restated against the host compile of crust-render_b200/csrc/shade/volume.h through the oracle's C hooks.
The reference walks with `openqmc::pcg::Rng::new(0xC0FFEE)`; openqmc-rs is not vendored, so the stream here is
this repo's PCG (values unpinned) — the tests are statistical or exact-analytic exactly like the originals."""
import math

import numpy as np
import pytest

from crust_render_b200.render_types import DensityField, VolumeRegion

I4 = np.eye(4, dtype=np.float32)
X_RAY = ((-2.0, 0.0, 0.0), (1.0, 0.0, 0.0))  # volume.rs:561-563


def unit_region(sigma_s, sigma_a, g, field):  # volume.rs:548-559
    return VolumeRegion(I4, (0.5, 0.5, 0.5), sigma_s, sigma_a, g, (0, 0, 0), 1.0, field)


def translation(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def test_homogeneous_transmittance_is_exact_beer_lambert(orc):  # volume.rs:566-586
    v = orc.Volumes([unit_region((0.7,) * 3, (0.2, 0.4, 0.9), 0.0, DensityField.Homogeneous())])
    s = orc.Rng(0xC0FFEE)
    tr = v.transmittance(*X_RAY, 1e-3, 10.0, s)
    expect = np.exp(-(np.float32(0.7) + np.array([0.2, 0.4, 0.9], np.float32)))
    assert np.abs(tr - expect).max() < 1e-5
    tr2 = v.transmittance(*X_RAY, 1e-3, 10.0, s)  # the fast path is deterministic
    assert (tr == tr2).all()


def test_ratio_tracking_matches_analytic_on_grid(orc):  # volume.rs:589-616
    d = 0.6
    v = orc.Volumes([unit_region((0.3, 0.5, 0.8), (0.4,) * 3, 0.0, DensityField.Grid(4, 4, 4, np.full(64, d, np.float32)))])
    s = orc.Rng(0xC0FFEE)
    mean = v.transmittance(*X_RAY, 1e-3, 10.0, s, n=20_000).mean(axis=0)
    expect = np.exp(-(np.array([0.3, 0.5, 0.8]) + 0.4) * d)
    assert np.abs(mean - expect).max() < 0.01


def test_delta_tracking_scatter_probability_matches_analytic(orc):  # volume.rs:619-645
    sigma = 1.3
    v = orc.Volumes([unit_region((sigma,) * 3, (0, 0, 0), 0.0, DensityField.Homogeneous())])
    s = orc.Rng(0xC0FFEE)
    ev = v.sample_interaction(*X_RAY, 1e-3, 10.0, s, n=20_000)
    sc = ev[ev[:, 0] > 0]
    assert np.abs(sc[:, 5:8] - 1.0).max() < 1e-5  # gray pure scattering: the event weight is exactly 1
    observed = len(sc) / len(ev)
    assert abs(observed - (1.0 - math.exp(-sigma))) < 0.01
    # the collision lies inside the slab and on the ray
    assert (sc[:, 1] > 1.5).all() and (sc[:, 1] < 2.5).all()
    assert np.abs(sc[:, 2] - (-2.0 + sc[:, 1])).max() < 1e-5
    assert (sc[:, 11] == 1).all() and np.abs(sc[:, 12] - 1.0).max() < 1e-6


def test_emission_walk_matches_analytic_slab(orc):  # volume.rs:648-683
    sigma_a, le = 0.8, np.array([4.0, 1.5, 0.3])
    v = orc.Volumes([VolumeRegion(I4, (0.5,) * 3, (0, 0, 0), (sigma_a,) * 3, 0.0, le, 1.0, DensityField.Homogeneous())])
    s = orc.Rng(0xC0FFEE)
    ev = v.sample_interaction(*X_RAY, 1e-3, 10.0, s, n=40_000)
    mean = ev[:, 8:11].astype(np.float64).mean(axis=0)
    expect = le * (1.0 - math.exp(-sigma_a))
    assert np.abs((mean - expect) / expect).max() < 0.03


def test_grid_trilinear_exact_at_centers(orc):  # volume.rs:686-714
    data = np.arange(8, dtype=np.float32)
    v = orc.Volumes([unit_region((1,) * 3, (0,) * 3, 0.0, DensityField.Grid(2, 2, 2, data))])
    for z in range(2):
        for y in range(2):
            for x in range(2):
                u = (0.25 + 0.5 * x, 0.25 + 0.5 * y, 0.25 + 0.5 * z)
                assert abs(v.field_density(0, u) - data[x + 2 * (y + 2 * z)]) < 1e-6
    assert abs(v.field_density(0, (0.5, 0.5, 0.5)) - 3.5) < 1e-6
    # edge clamp: within half a voxel of the boundary the edge value holds (volume.rs:152-155)
    assert abs(v.field_density(0, (0.0, 0.0, 0.0)) - 0.0) < 1e-6 and abs(v.field_density(0, (1.0, 1.0, 1.0)) - 7.0) < 1e-6
    assert abs(v.region(0)["majorant"] - 7.0) < 1e-6  # max_channel(sigma_t) * max(data) (volume.rs:80-82, 220)


def test_noise_deterministic_and_bounded_by_majorant(orc):  # volume.rs:717-748
    v = orc.Volumes([unit_region((1,) * 3, (0,) * 3, 0.0, DensityField.Noise(4.0, 4, 0.5, 2.0, 0.3, 42))])
    runs = []
    for _ in range(2):
        vals = []
        for i in range(1000):
            u = (orc.vol_hash3(i, 1, 2, 7), orc.vol_hash3(i, 3, 4, 7), orc.vol_hash3(i, 5, 6, 7))
            d = v.field_density(0, u)
            assert 0.0 <= d <= 1.0
            vals.append(d)
        runs.append(vals)
    assert runs[0] == runs[1]
    assert any(d > 0.0 for d in runs[0])


def test_hash3_is_the_reference_integer_mix(orc):  # volume.rs:88-99, recomputed with Python integers
    def ref(ix, iy, iz, seed):
        M = 0xFFFFFFFF
        h = ((ix & M) * 0x8DA6B343 & M) ^ ((iy & M) * 0xD8163841 & M) ^ ((iz & M) * 0xCB1AB31F & M) ^ (seed * 0x9E3779B9 & M)
        h ^= h >> 15
        h = h * 0x2C1B3C6D & M
        h ^= h >> 12
        h = h * 0x297A2D39 & M
        h ^= h >> 15
        return np.float32(h >> 8) / np.float32(1 << 24)

    for args in [(0, 0, 0, 0), (1, 2, 3, 4), (-1, -7, 12, 42), (123456, -98765, 4242, 0xDEADBEEF), (2**31 - 1, -(2**31), 5, 7)]:
        assert orc.vol_hash3(*args) == float(ref(*args))


def test_oriented_box_interval_in_world_units(orc):  # volume.rs:751-786
    xf = translation(5, 0, 0) @ np.diag([2.0, 1.0, 1.0, 1.0]).astype(np.float32)
    v = orc.Volumes([VolumeRegion(xf, (1, 1, 1), (1, 1, 1), (0, 0, 0), 0.0, (0, 0, 0), 1.0, DensityField.Homogeneous())])
    t0, t1 = v.intersect(0, *X_RAY)
    assert abs(t0 - 5.0) < 1e-4 and abs(t1 - 9.0) < 1e-4
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    rot = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)  # Mat4::from_rotation_z
    v = orc.Volumes([VolumeRegion(rot, (1, 1, 1), (1, 1, 1), (0, 0, 0), 0.0, (0, 0, 0), 1.0, DensityField.Homogeneous())])
    t0, t1 = v.intersect(0, *X_RAY)
    r2 = math.sqrt(2.0)
    assert abs(t0 - (2.0 - r2)) < 1e-4 and abs(t1 - (2.0 + r2)) < 1e-4
    # world bounds of the rotated cube (volume.rs:207-219)
    info = v.region(0)
    assert np.allclose(info["aabb_min"], (-r2, -r2, -1), atol=1e-5) and np.allclose(info["aabb_max"], (r2, r2, 1), atol=1e-5)


def test_overlapping_regions_compose_exactly(orc):  # volume.rs:789-809
    a = unit_region((0,) * 3, (0.5,) * 3, 0.0, DensityField.Homogeneous())
    b = VolumeRegion(translation(0.25, 0, 0), (0.5,) * 3, (0,) * 3, (0.75,) * 3, 0.0, (0,) * 3, 1.0, DensityField.Homogeneous())
    v = orc.Volumes([a, b])
    tr = v.transmittance(*X_RAY, 1e-3, 10.0, orc.Rng(0xC0FFEE))
    assert abs(tr[0] - math.exp(-0.5 - 0.75)) < 1e-5


def test_phase_mix_pdf_matches_single_lobe(orc):  # volume.rs:812-818
    for mu in (-0.9, -0.2, 0.0, 0.5, 0.99):
        assert abs(orc.phase_single_pdf(0.4, mu) - orc.hg_phase(mu, 0.4)) < 1e-7


def test_hg_phase_normalizes_over_sphere(orc):  # medium.rs:186-203
    for g in (-0.7, 0.0, 0.4, 0.9):
        n = 20_000
        total = sum(orc.hg_phase(-1.0 + 2.0 * (i + 0.5) / n, g) for i in range(0, n, 1))
        assert abs(2.0 * math.pi * total * (2.0 / n) - 1.0) < 1e-3


def test_region_constructor_folds_density_scale_and_clamps_g(orc):  # volume.rs:205-206, 220, 227
    v = orc.Volumes([VolumeRegion(I4, (0.5,) * 3, (0.2, 0.4, 0.6), (0.1, 0.1, 0.1), 1.5, (0,) * 3, 3.0, DensityField.Homogeneous())])
    info = v.region(0)
    assert np.allclose(info["sigma_s"], (0.6, 1.2, 1.8), rtol=1e-6) and np.allclose(info["sigma_a"], (0.3,) * 3, rtol=1e-6)
    assert abs(info["g"] - 0.99) < 1e-7 and abs(info["majorant"] - 2.1) < 1e-6 and info["homogeneous"]


def test_density_is_zero_outside_the_box_and_local_inside(orc):  # volume.rs:233-241
    xf = translation(1, 2, 3) @ np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    data = np.arange(8, dtype=np.float32)
    v = orc.Volumes([VolumeRegion(xf, (0.5,) * 3, (1,) * 3, (0,) * 3, 0.0, (0,) * 3, 1.0, DensityField.Grid(2, 2, 2, data))])
    assert v.density(0, (1, 2, 3)) == pytest.approx(3.5, abs=1e-6)  # the box centre
    assert v.density(0, (1.5, 2.5, 3.5)) == pytest.approx(7.0, abs=1e-6)  # voxel (1,1,1) centre: local +0.25 -> world +0.5
    assert v.density(0, (2.01, 2, 3)) == 0.0 and v.density(0, (1, 0.9, 3)) == 0.0


def test_mat4_inverse_inverts(orc):  # glam Mat4::inverse (volume.rs:222)
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = rng.normal(size=(3, 3)).astype(np.float32) + 2.0 * np.eye(3, dtype=np.float32)
        m[:3, 3] = rng.normal(size=3).astype(np.float32) * 5.0
        inv = orc.mat4_inverse(m)
        assert np.abs(inv @ m - np.eye(4)).max() < 2e-5
        assert np.allclose(inv[3], (0, 0, 0, 1), atol=1e-6)
