"""The reference's scene-level known-answer tests, written once against the
crust_rt-shaped API and run twice: on the CPU oracle (test_oracle_scene.py,
pins the oracle) and on the CUDA backend (test_gpu_scene.py, parity).

Each function mirrors one #[test] of /root/reference/crates/crust-rt/src; the
citation is in its docstring. `api` is either `orc` or `crust_render_b200`.
"""
import math

import numpy as np

from crust_render_b200.rt_types import (MASK_ALL, MASK_CAMERA, MASK_SHADOW, Affine3A, CubicCurveSegment, CurveSegment,
                                        Geometry, Ray)
from crust_render_b200 import workloads

INF = float("inf")
X, Y, Z = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)


def neg(v):
    return tuple(-c for c in v)


def close(a, b, eps):
    return np.allclose(np.asarray(a, np.float32), np.asarray(b, np.float32), atol=eps, rtol=0)


def unit_sphere_scene(api):
    b = api.SceneBuilder()
    b.attach(Geometry.Sphere((0, 0, 0), 1.0))
    return b.commit()


def kat_doc_example(api):
    """lib.rs:9-23"""
    b = api.SceneBuilder()
    ball = b.attach(Geometry.Sphere((0, 0, 0), 1.0))
    scene = b.commit()
    hit = scene.intersect(Ray((0, 0, -5), Z), 0.001, INF)
    assert hit is not None and hit.geom_id == ball and abs(hit.t - 4.0) < 1e-4
    assert not scene.occluded(Ray((0, 0, -5), Z), 0.001, 3.9)


def kat_reserved_slots(api):
    """scene.rs:498-537"""
    b = api.SceneBuilder()
    a = b.attach(Geometry.Sphere((-5, 0, 0), 1.0))
    placeholder = b.attach(api.SceneBuilder.empty_geometry())
    c = b.attach(Geometry.Sphere((5, 0, 0), 1.0))
    assert (a, placeholder, c) == (0, 1, 2)
    scene = b.commit()
    assert scene.geometry_count() == 3 and scene.primitive_count() == 2
    b = api.SceneBuilder()
    b.attach(Geometry.Sphere((-5, 0, 0), 1.0))
    slot = b.attach(api.SceneBuilder.empty_geometry())
    b.set_geometry(slot, Geometry.Sphere((0, 0, 0), 1.0))
    scene = b.commit()
    assert scene.primitive_count() == 2
    hit = scene.intersect(Ray((0, 0, -8), Z), 1e-4, 3.4028235e38)
    assert hit is not None and hit.geom_id == slot


def kat_ids_map_back(api):
    """scene.rs:540-579"""
    b = api.SceneBuilder()
    ball = b.attach(Geometry.Sphere((-3, 0, 0), 1.0))
    quad = b.attach(Geometry.TriangleMesh([(2, -1, -1), (2, -1, 1), (2, 1, 1), (2, 1, -1)], [(0, 1, 2), (0, 2, 3)]))
    scene = b.commit()
    assert scene.geometry_count() == 2 and scene.primitive_count() == 3
    h = scene.intersect(Ray((-3, 0, -5), Z), 0.001, INF)
    assert h is not None and h.geom_id == ball
    h = scene.intersect(Ray((0, 0.5, -0.5), X), 0.001, INF)
    assert h is not None and h.geom_id == quad and h.prim_id == 1


def kat_front_face(api):
    """scene.rs:582-596"""
    scene = unit_sphere_scene(api)
    outside = scene.intersect(Ray((0, 0, -5), Z), 0.001, 100.0)
    assert outside.front_face and close(outside.normal, neg(Z), 1e-4)
    inside = scene.intersect(Ray((0, 0, 0), Z), 0.001, 100.0)
    assert (not inside.front_face) and close(inside.normal, neg(Z), 1e-4)


def kat_motion_flag(api):
    """scene.rs:603-652"""
    assert not unit_sphere_scene(api).has_motion()
    b = api.SceneBuilder()
    b.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.from_translation((3, 0, 0))))
    assert not b.commit().has_motion()
    mid = api.SceneBuilder()
    mid.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.IDENTITY, Affine3A.from_translation((4, 0, 0))))
    mid = mid.commit()
    assert mid.has_motion()
    root = api.SceneBuilder()
    root.attach(Geometry.Instance(mid, Affine3A.from_translation((0, 7, 0))))
    assert root.commit().has_motion()


def _nested(api):
    leaf = unit_sphere_scene(api)
    mid = api.SceneBuilder()
    for x in (-2.0, 2.0):
        mid.attach(Geometry.Instance(leaf, Affine3A.from_translation((x, 0, 0))))
    mid = mid.commit()
    root = api.SceneBuilder()
    for y in (-5.0, 5.0):
        root.attach(Geometry.Instance(mid, Affine3A.from_translation((0, y, 0))))
    return root.commit()


def kat_instances_nest(api):
    """scene.rs:655-710"""
    scene = _nested(api)
    assert scene.primitive_count() == 2
    for x, y in [(-2.0, -5.0), (2.0, -5.0), (-2.0, 5.0), (2.0, 5.0)]:
        ray = Ray((x, y, -8), Z)
        hit = scene.intersect(ray, 0.001, 100.0)
        assert hit is not None and abs(hit.t - 7.0) < 1e-3
        assert close(hit.normal, neg(Z), 1e-4)
        assert scene.occluded(ray, 0.001, 100.0)
    assert scene.intersect(Ray((0, 0, -8), Z), 0.001, 100.0) is None


def kat_unique_breakdown(api):
    """scene.rs:716-748"""
    scene = _nested(api)
    top = scene.primitive_breakdown()
    assert top.instances == 2 and top.spheres == 0
    uniq = scene.unique_primitive_breakdown()
    assert uniq.spheres == 1 and uniq.instances == 4


def kat_nested_compose(api):
    """scene.rs:755-803"""
    mid = api.SceneBuilder()
    mid.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.from_scale((2, 1, 1))))
    mid = mid.commit()
    root = api.SceneBuilder()
    root.attach(Geometry.Instance(mid, Affine3A.from_rotation_z(math.pi / 2)))
    scene = root.commit()
    ay = scene.intersect(Ray((0, -8, 0), Y), 0.001, 100.0)
    assert ay is not None and abs(ay.t - 6.0) < 1e-3
    ax = scene.intersect(Ray((-8, 0, 0), X), 0.001, 100.0)
    assert ax is not None and abs(ax.t - 7.0) < 1e-3
    assert close(ay.normal, neg(Y), 1e-4)


def kat_nested_masks(api):
    """scene.rs:808-837"""
    mid = api.SceneBuilder()
    mid.attach_masked(Geometry.Instance(unit_sphere_scene(api), Affine3A.IDENTITY), MASK_SHADOW)
    mid = mid.commit()
    root = api.SceneBuilder()
    root.attach_masked(Geometry.Instance(mid, Affine3A.IDENTITY), MASK_SHADOW | MASK_CAMERA)
    scene = root.commit()
    ray = Ray((0, 0, -5), Z)
    assert scene.intersect(ray.with_mask(MASK_CAMERA), 0.001, 100.0) is None
    assert scene.intersect(ray.with_mask(MASK_SHADOW), 0.001, 100.0) is not None


def kat_masks_filter(api):
    """scene.rs:840-855"""
    b = api.SceneBuilder()
    b.attach_masked(Geometry.Sphere((0, 0, 0), 1.0), MASK_SHADOW)
    scene = b.commit()
    ray = Ray((0, 0, -5), Z)
    assert scene.intersect(ray.with_mask(MASK_CAMERA), 0.001, 100.0) is None
    assert scene.intersect(ray.with_mask(MASK_SHADOW), 0.001, 100.0) is not None
    assert not scene.occluded(ray.with_mask(MASK_CAMERA), 0.001, 100.0)
    assert scene.occluded(ray.with_mask(MASK_SHADOW), 0.001, 100.0)


def kat_smooth_normals(api):
    """scene.rs:858-882"""
    def nrm(v):
        v = np.asarray(v, np.float32)
        return v / np.float32(np.sqrt((v * v).sum(dtype=np.float32)))

    b = api.SceneBuilder()
    b.attach(Geometry.TriangleMesh([(-1, -1, 2), (1, -1, 2), (0, 1, 2)], [(0, 1, 2)],
                                   normals=[nrm((-0.5, 0, -1)), nrm((0.5, 0, -1)), nrm((0, 0.5, -1))]))
    scene = b.commit()
    hit = scene.intersect(Ray((0.6, -0.7, 0), Z), 0.001, 100.0)
    assert hit is not None and hit.normal[0] > 0.1 and hit.normal[2] < 0.0


def kat_translated_instance(api):
    """scene.rs:885-899"""
    b = api.SceneBuilder()
    b.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.from_translation((3, 0, 0))))
    scene = b.commit()
    ray = Ray((3, 0, -5), Z)
    hit = scene.intersect(ray, 0.001, INF)
    assert hit is not None and abs(hit.t - 4.0) < 1e-4 and close(hit.normal, neg(Z), 1e-4)
    assert scene.occluded(ray, 0.001, INF) and not scene.occluded(ray, 0.001, 3.9)


def kat_nonuniform_scale_normals(api):
    """scene.rs:902-925"""
    b = api.SceneBuilder()
    b.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.from_scale((2, 1, 1))))
    scene = b.commit()
    hit = scene.intersect(Ray((1, 5, 0), neg(Y)), 0.001, INF)
    e = np.array([0.5, math.sqrt(3.0), 0.0])
    assert hit is not None and close(hit.normal, e / np.linalg.norm(e), 1e-3)


def kat_rotated_instance(api):
    """scene.rs:928-955"""
    inner = api.SceneBuilder()
    inner.attach(Geometry.TriangleMesh([(-1, -1, 0), (1, -1, 0), (0, 1, 0)], [(0, 1, 2)]))
    b = api.SceneBuilder()
    b.attach(Geometry.Instance(inner.commit(), Affine3A.from_rotation_y(math.pi / 2) * Affine3A.from_translation((0, 0, 2))))
    scene = b.commit()
    hit = scene.intersect(Ray((5, 0, 0), neg(X)), 0.001, INF)
    assert hit is not None and abs(hit.t - 3.0) < 1e-4 and close(hit.normal, X, 1e-4)


def kat_motion_blur(api):
    """scene.rs:958-981"""
    b = api.SceneBuilder()
    b.attach(Geometry.Instance(unit_sphere_scene(api), Affine3A.IDENTITY, Affine3A.from_translation((4, 0, 0))))
    scene = b.commit()
    assert scene.intersect(Ray((0, 0, -5), Z).with_time(0.0), 0.001, INF) is not None
    assert scene.intersect(Ray((4, 0, -5), Z).with_time(1.0), 0.001, INF) is not None
    assert scene.intersect(Ray((0, 0, -5), Z).with_time(1.0), 0.001, INF) is None
    hit = scene.intersect(Ray((2, 0, -5), Z).with_time(0.5), 0.001, INF)
    assert hit is not None and abs(hit.t - 4.0) < 1e-4
    bb = scene.bounds()
    assert bb.minimum[0] <= -1.0 and bb.maximum[0] >= 5.0


def kat_instance_ids(api):
    """scene.rs:984-1015"""
    inner = api.SceneBuilder()
    inner.attach(Geometry.TriangleMesh([(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)], [(0, 1, 2), (0, 2, 3)]))
    b = api.SceneBuilder()
    b.attach(Geometry.Sphere((0, -100, 0), 1.0))
    inst = b.attach(Geometry.Instance(inner.commit(), Affine3A.from_translation((0, 0, 5))))
    scene = b.commit()
    hit = scene.intersect(Ray((-0.5, 0.5, 0), Z), 0.001, INF)
    assert hit is not None and hit.geom_id == inst and hit.prim_id == 1


def kat_empty_scene(api):
    """bvh.rs:1663-1668"""
    scene = api.SceneBuilder().commit()
    assert scene.intersect(Ray((0, 0, 0), X), 0.001, INF) is None
    assert not scene.occluded(Ray((0, 0, 0), X), 0.001, INF)
    assert scene.bounds() is None


def kat_bounds_cover(api):
    """bvh.rs:1671-1676"""
    bb = workloads.sphere_grid_prims(api, 3).commit().bounds()
    assert np.all(bb.minimum <= -0.5) and np.all(bb.maximum >= 6.5)


def kat_curve_scene(api):
    """curve.rs:228-306 expressed through RoundCurves geometry (prim.rs:186-200)."""
    b = api.SceneBuilder()
    g = b.attach(Geometry.RoundCurves([CurveSegment((0, 0, 0), (4, 0, 0), 0.5, 0.5), CurveSegment((0, 5, 0), (4, 5, 0), 0.5, 0.1)]))
    scene = b.commit()
    h = scene.intersect(Ray((-3, 0, 0), X), 0.001, INF)  # axial: cap at x = -0.5
    assert h is not None and abs(h.t - 2.5) < 1e-4 and close(h.normal, neg(X), 1e-4) and h.geom_id == g and h.prim_id == 0
    h = scene.intersect(Ray((2, 3, 0), neg(Y)), 0.001, INF)  # side hit at radius
    assert h is not None and abs(h.t - 2.5) < 1e-3 and close(h.normal, Y, 1e-3)
    assert scene.intersect(Ray((2, 3, 0), neg(Y)), 0.001, 2.0) is None
    assert scene.intersect(Ray((2, 3, 2), neg(Y)), 0.001, INF) is None
    h = scene.intersect(Ray((2, 3, 0), (0, -2, 0)), 0.001, INF)  # unnormalised dir: half the t
    assert h is not None and abs(h.t - 1.25) < 1e-3
    thick = scene.intersect(Ray((0.5, 8, 0), neg(Y)), 0.001, INF)
    thin = scene.intersect(Ray((3.5, 8, 0), neg(Y)), 0.001, INF)
    assert thick.prim_id == 1 and thin.prim_id == 1
    assert (3.0 - thick.t) > (3.0 - thin.t) + 0.2


def kat_cubic_scene(api):
    """curve.rs:348-414 expressed through CubicCurves geometry (prim.rs:230-244)."""
    K = 0.55228475
    straight = [(0, 0, 0), (4 / 3, 0, 0), (8 / 3, 0, 0), (4, 0, 0)]
    arc = [(11, 0, 10), (11, K, 10), (10 + K, 1, 10), (10, 1, 10)]
    b = api.SceneBuilder()
    b.attach(Geometry.CubicCurves([CubicCurveSegment(straight, 0.5, 0.5), CubicCurveSegment(arc, 0.05, 0.05)]))
    scene = b.commit()
    h = scene.intersect(Ray((2, 3, 0), neg(Y)), 0.001, INF)
    assert h is not None and abs(h.t - 2.5) < 1e-3 and close(h.normal, Y, 1e-3) and h.prim_id == 0
    assert scene.intersect(Ray((2, 3, 2), neg(Y)), 0.001, INF) is None
    assert scene.intersect(Ray((2, 3, 0), neg(Y)), 0.001, 2.0) is None
    s = math.sqrt(2.0) / 2.0
    h = scene.intersect(Ray((10 + s, s, 20), neg(Z)), 0.001, INF)  # the true arc point must hit
    assert h is not None and h.prim_id == 1
    assert scene.intersect(Ray((10.5, 0.5, 20), neg(Z)), 0.001, INF) is None  # the chord midpoint must miss


ALL = [v for k, v in sorted(globals().items()) if k.startswith("kat_")]
