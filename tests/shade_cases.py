"""Materials, query generators and the comparison used by the two independent-witness shading tests
(tests/test_shade_independent.py on the CPU, tests/test_gpu_shade_independent.py on the device)."""
import math
import zlib

import numpy as np

import openpbr_ref as R

REL, ABS = 2e-3, 1e-6

SHOWCASE = {  # samples/openpbr_showcase.usda:38-145 (crust-render_b200/scenes.py openpbr_showcase)
    "chrome": dict(base_color=(0.95, 0.95, 0.95), base_metalness=1.0, specular_roughness=0.02),
    "plastic": dict(base_color=(0.8, 0.1, 0.1)),
    "car_paint": dict(base_color=(0.8, 0.05, 0.05), base_metalness=0.9, specular_roughness=0.25, coat_weight=1.0, coat_roughness=0.02, coat_ior=1.5),
    "velvet": dict(base_color=(0.35, 0.15, 0.55), base_diffuse_roughness=0.9, specular_roughness=0.9, fuzz_weight=1.0, fuzz_roughness=0.3),
    "dispersive_glass": dict(specular_roughness=0.02, specular_ior=1.55, transmission_weight=1.0, transmission_color=(0.7, 0.9, 0.75),
                             transmission_depth=1.0, transmission_dispersion_scale=0.6),
    "wax": dict(base_color=(0.9, 0.75, 0.55), specular_roughness=0.5, subsurface_weight=1.0, subsurface_color=(0.95, 0.55, 0.35), subsurface_radius=1.5),
    "soap_bubble": dict(base_color=(0.8, 0.8, 0.8), specular_roughness=0.02, specular_ior=1.33, transmission_weight=0.4,
                        transmission_color=(0.95, 0.95, 1.0), thin_film_weight=1.0, thin_film_thickness=0.55, geometry_thin_walled=True),
    "default_grey": dict(),
}
VEACH = {  # samples/veach_mis.usda:80-140
    f"plate_{r}": dict(base_color=(0.9, 0.9, 0.9), base_metalness=1.0, specular_roughness=r) for r in (0.01, 0.05, 0.12, 0.25)
}
VEACH["backdrop"] = dict(base_color=(0.1, 0.1, 0.1), specular_weight=0.0)
EXTRA = {
    "brushed_metal": dict(base_color=(0.9, 0.6, 0.3), base_metalness=1.0, specular_roughness=0.35, specular_roughness_anisotropy=0.7,
                          specular_color=(0.8, 0.9, 1.0), specular_weight=0.8),
    "film_on_metal": dict(base_color=(0.8, 0.7, 0.6), base_metalness=1.0, specular_roughness=0.2, thin_film_weight=0.7, thin_film_thickness=0.4, thin_film_ior=1.8),
    "film_under_coat": dict(base_color=(0.2, 0.3, 0.7), specular_roughness=0.3, thin_film_weight=1.0, thin_film_thickness=0.3, coat_weight=0.6,
                            coat_roughness=0.1, coat_color=(0.9, 0.6, 0.3), coat_darkening=0.5, coat_roughness_anisotropy=0.4),
    "tinted_glass": dict(specular_roughness=0.15, specular_ior=1.5, transmission_weight=0.8, transmission_color=(0.8, 0.95, 0.8), base_color=(0.3, 0.2, 0.9)),
    "inverted_ior": dict(specular_roughness=0.1, specular_ior=0.75, transmission_weight=1.0, transmission_dispersion_scale=1.5, transmission_dispersion_abbe_number=30.0),
    "half_metal_fuzz": dict(base_color=(0.6, 0.6, 0.2), base_metalness=0.5, base_weight=0.7, specular_roughness=0.4, fuzz_weight=0.4, fuzz_color=(0.9, 0.2, 0.2),
                            fuzz_roughness=0.7, base_diffuse_roughness=0.5, subsurface_weight=0.3),
}
MATERIALS = {**SHOWCASE, **VEACH, **EXTRA}


def _unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def _directions(rng, n):
    """Unit vectors over the whole sphere, plus near-grazing and near-normal ones."""
    d = rng.normal(size=(n, 3))
    d[: n // 8, 2] *= 0.02  # grazing
    d[n // 8: n // 4, :2] *= 0.02  # along the normal
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def _frame_queries(rng, n):
    """n (ray_dir, normal, front_face, wi) tuples: a random shading normal, a view direction above it, a light direction anywhere."""
    normal = _directions(rng, n)
    v = _directions(rng, n)
    flip = np.einsum("ij,ij->i", v, normal) < 0
    v[flip] -= 2.0 * np.einsum("ij,ij->i", v[flip], normal[flip])[:, None] * normal[flip]  # mirror into the upper hemisphere
    wi = _directions(rng, n)
    ray_dir = -v * rng.uniform(0.5, 3.0, (n, 1))  # unnormalised, like real rays (camera.rs:83)
    ff = rng.random(n) < 0.6
    return ray_dir, normal, ff, wi * rng.uniform(0.5, 2.0, (n, 1))


def _close(a, b, what, rel=REL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (rel * np.abs(b) + ABS)
    worst = np.unravel_index(np.argmax(err), err.shape) if err.size else None
    assert err.size == 0 or err.max() <= 0, f"{what}: device {a[worst]} vs witness {b[worst]} at {worst}"




def record_of(params):
    from crust_render_b200.render_types import OpenPBR
    return OpenPBR(**{k: (int(v) if k == "geometry_thin_walled" else v) for k, v in params.items()})


def check_eval(name, evaluate, n=1500):
    """`evaluate(record, ray_dir, normal, front_face, wi) -> (value[n,3], pdf[n], some[n])` of the implementation under test
    against the witness, for material `name`."""
    params = MATERIALS[name]
    m = R.M(**params)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    ray_dir, normal, ff, wi = _frame_queries(rng, n)
    # what the implementation sees is the f32 rounding of these inputs: hand the witness the same numbers
    ray_dir, normal, wi = (a.astype(np.float32) for a in (ray_dir, normal, wi))
    value, pdf, some = evaluate(record_of(params), ray_dir, normal, ff, wi)
    want_v, want_p, keep = [], [], []
    for i in range(n):
        r = R.material_eval(m, tuple(map(float, ray_dir[i])), tuple(map(float, normal[i])), bool(ff[i]), tuple(map(float, wi[i])))
        assert (r is not None) == bool(some[i]), (name, i)
        if r is None:
            continue
        # spots where f32 cancellation dominates (the GGX peak of a 1e-4-alpha lobe met within 1e-3 rad) are excluded by
        # magnitude: they carry pdf > 1e5
        if r[1] > 1e5:
            continue
        keep.append(i)
        want_v.append(r[0])
        want_p.append(r[1])
    assert len(keep) > n // 2
    _close(np.asarray(value)[keep], want_v, f"{name} eval value")
    _close(np.asarray(pdf)[keep], want_p, f"{name} eval pdf")


def check_scatter(name, scatter, n=1500):
    """`scatter(record, ray_dir, normal, front_face, s4) -> (dir[n,3], value[n,3], pdf[n], delta[n], some[n])` against the witness."""
    params = MATERIALS[name]
    m = R.M(**params)
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    ray_dir, normal, ff, _ = _frame_queries(rng, n)
    s4 = rng.random((n, 4)).astype(np.float32)
    s4[:, 0] = ((np.arange(n) + rng.random(n)) / n).astype(np.float32)  # stratified lobe pick: every lobe with selection mass gets its share
    ray_dir, normal = ray_dir.astype(np.float32), normal.astype(np.float32)
    d, value, pdf, delta, some = scatter(record_of(params), ray_dir, normal, ff, s4)
    pmf = R.lobe_pmf(m)
    edges = np.cumsum(pmf)
    checked, lobes = 0, set()
    for i in range(n):
        if np.min(np.abs(edges - float(s4[i, 0]))) < 1e-5:
            continue  # a pick within f32 rounding of a lobe boundary may fall either side
        nrm = tuple(map(float, normal[i]))
        r = R.material_scatter(m, tuple(map(float, ray_dir[i])), nrm, bool(ff[i]), tuple(map(float, s4[i])))
        if r is None or not some[i]:
            if (r is None) != (not some[i]):
                # only a direction within rounding of the horizon / of the critical angle may be accepted by one side alone
                assert r is not None and (abs(R.dot(r["dir"], nrm)) < 1e-3 or r["lobe"] == R.TRANSMISSION), (name, i)
            continue
        assert bool(delta[i]) == r["delta"], (name, i)
        if r["pdf"] > 1e5:
            continue
        ang = math.acos(max(-1.0, min(1.0, float(np.dot(_unit(d[i]), _unit(r["dir"]))))))
        # a needle lobe (alpha 1e-4) sampled near the rim of the disk (u1 -> 1) is ill-conditioned in f32: sin/cos of 2 ulp move
        # the half vector by ~1e-3 rad there; a sign or argument slip moves it by 0.1-1 rad
        assert ang < 5e-3, (name, i, ang, r["lobe"])
        if r["delta"]:
            _close(value[i], r["value"], f"{name} delta throughput", rel=4e-3)
            assert pdf[i] == 1.0
        else:
            # value and pdf are those of eval() at the sampled direction (openpbr.rs:1259-1284 eval_matches_scatter_importance).
            # They are compared at the direction the implementation under test returned, not at the witness's own sample:
            # across a needle lobe (alpha 1e-4 rad wide) the value changes by percents within the 1e-3 rad the two samples
            # may differ by. Even so the f32 direction carries ~1e-7 rad of rounding, so needles (pdf > 1e4) are left out.
            e = R.material_eval(m, tuple(map(float, ray_dir[i])), nrm, bool(ff[i]), tuple(map(float, d[i])))
            if e is None or e[1] > 1e4 or r["pdf"] > 1e4:
                continue
            _close(value[i], e[0], f"{name} scatter value (lobe {r['lobe']})", rel=1e-2)
            _close(pdf[i], e[1], f"{name} scatter pdf (lobe {r['lobe']})", rel=1e-2)
        checked += 1
        lobes.add(r["lobe"])
    assert checked >= 5, checked  # the 0.01-roughness plate (alpha 1e-4) is all needle: few samples stay below the pdf cut
    expect = {k for k in range(5) if pmf[k] > 0.02}
    assert expect <= lobes, (name, expect, lobes)
