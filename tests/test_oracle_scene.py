"""Pins oracle/ against the reference's scene-, curve- and BVH-level known-answer
tests (crates/crust-rt/src/{scene.rs:481-1016, curve.rs:219-415, bvh.rs:1377-1766})."""
import numpy as np
import pytest

import kats
from crust_render_b200 import workloads
from crust_render_b200.rt_types import Geometry, Ray

INF = float("inf")


@pytest.mark.parametrize("kat", kats.ALL, ids=lambda f: f.__name__)
def test_scene_kat(orc, kat):
    kat(orc)


# ---- curve.rs:219-415 on the bare functions -------------------------------------
def test_capsule_axial_hit_at_cap(orc):  # curve.rs:228-242
    t, n = orc.rounded_cone_intersect((-3, 0, 0), (1, 0, 0), (0, 0, 0), (4, 0, 0), 0.5, 0.5, 0.001, INF)
    assert abs(t - 2.5) < 1e-4 and np.allclose(n, (-1, 0, 0), atol=1e-4)


def test_capsule_perpendicular_hit_at_radius(orc):  # curve.rs:245-257
    t, n = orc.rounded_cone_intersect((2, 3, 0), (0, -1, 0), (0, 0, 0), (4, 0, 0), 0.5, 0.5, 0.001, INF)
    assert abs(t - 2.5) < 1e-3 and np.allclose(n, (0, 1, 0), atol=1e-3)


def test_cone_radius_shrinks_along_axis(orc):  # curve.rs:260-277
    p0, p1 = (0, 0, 0), (4, 0, 0)
    t_thick = orc.rounded_cone_intersect((0.5, 3, 0), (0, -1, 0), p0, p1, 0.5, 0.1, 0.001, INF)[0]
    t_thin = orc.rounded_cone_intersect((3.5, 3, 0), (0, -1, 0), p0, p1, 0.5, 0.1, 0.001, INF)[0]
    thick, thin = 3.0 - t_thick, 3.0 - t_thin
    assert thick > thin + 0.2 and thick <= 0.5 + 1e-3 and thin >= 0.1 - 1e-3


def test_cone_respects_t_range_and_misses(orc):  # curve.rs:280-290
    p0, p1 = (0, 0, 0), (4, 0, 0)
    assert orc.rounded_cone_intersect((2, 3, 0), (0, -1, 0), p0, p1, 0.5, 0.5, 0.001, 2.0) is None
    assert orc.rounded_cone_intersect((2, 3, 2), (0, -1, 0), p0, p1, 0.5, 0.5, 0.001, INF) is None
    t, _ = orc.rounded_cone_intersect((2, 3, 0), (0, -2, 0), p0, p1, 0.5, 0.5, 0.001, INF)
    assert abs(t - 1.25) < 1e-3


def test_degenerate_swallowed_sphere_still_hits(orc):  # curve.rs:293-306
    t, _ = orc.rounded_cone_intersect((0, 0, -5), (0, 0, 1), (0, 0, 0), (0.1, 0, 0), 1.0, 0.05, 0.001, INF)
    assert abs(t - 4.0) < 1e-3


K = 0.55228475
STRAIGHT = [(0, 0, 0), tuple(np.float32(4) * np.float32(1 / 3) * np.array([1, 0, 0], np.float32)),
            tuple(np.float32(4) * np.float32(2 / 3) * np.array([1, 0, 0], np.float32)), (4, 0, 0)]
ARC = [(1, 0, 0), (1, K, 0), (K, 1, 0), (0, 1, 0)]


def test_flat_span_needs_no_subdivision(orc):  # curve.rs:324-331
    assert orc.flatness_depth(STRAIGHT, 1.0) == 0


def test_curved_span_requires_subdivision(orc):  # curve.rs:334-345
    assert orc.flatness_depth(ARC, 0.1) > 0


def test_straight_span_matches_rounded_cone(orc):  # curve.rs:348-362
    cubic = orc.cubic_curve_intersect((2, 3, 0), (0, -1, 0), STRAIGHT, 0.5, 0.5, 0.001, INF)
    cone = orc.rounded_cone_intersect((2, 3, 0), (0, -1, 0), STRAIGHT[0], STRAIGHT[3], 0.5, 0.5, 0.001, INF)
    assert abs(cubic[0] - cone[0]) < 1e-4 and np.allclose(cubic[1], cone[1], atol=1e-4)


def test_quarter_circle_arc_is_followed_not_its_chord(orc):  # curve.rs:365-397
    s = np.float32(np.sqrt(2.0) / 2.0)
    assert orc.cubic_curve_intersect((s, s, 10), (0, 0, -1), ARC, 0.05, 0.05, 0.001, INF) is not None
    assert orc.cubic_curve_intersect((0.5, 0.5, 10), (0, 0, -1), ARC, 0.05, 0.05, 0.001, INF) is None


def test_cubic_wide_miss_and_range(orc):  # curve.rs:400-414
    assert orc.cubic_curve_intersect((2, 3, 2), (0, -1, 0), STRAIGHT, 0.5, 0.5, 0.001, INF) is None
    assert orc.cubic_curve_intersect((2, 3, 0), (0, -1, 0), STRAIGHT, 0.5, 0.5, 0.001, 2.0) is None


# ---- bvh.rs:1377-1766 -------------------------------------------------------------
ORIGINS = [(-5.0, 4.5, 4.5), (20.0, 3.0, 3.0), (4.5, -5.0, 4.5), (0.0, 0.0, -10.0), (5.0, 5.2, -3.0)]


def _norm(v):
    v = np.asarray(v, np.float32)
    return v * (np.float32(1.0) / np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])))


DIRS = [(1.0, 0.0, 0.0), _norm((-1.0, 0.05, 0.02)), (0.0, 1.0, 0.0), _norm((0.3, 0.3, 1.0)), (0.0, 0.0, -1.0), (0.577, 0.577, 0.577)]


def _shards_scene(api, n):
    b = api.SceneBuilder()
    b.attach(Geometry.TriangleMesh(*workloads.diagonal_shards(n)))
    return b.commit()


def _assert_matches_linear(scene):  # bvh.rs:1472-1515
    for o in ORIGINS:
        for d in DIRS:
            ray = Ray(o, d)
            a = scene.intersect(ray, 0.001, INF)
            b = scene.linear_scan(ray, 0.001, INF)
            assert (a is None) == (b is None), f"hit disagreement for {o} {d}"
            if a is not None:
                assert abs(a.t - float(b["t"])) < 1e-4 and a.geom_id == int(b["geom_id"])
            assert scene.occluded(ray, 0.001, INF) == (b is not None)


def test_matches_linear_scan(orc):  # bvh.rs:1518-1521
    _assert_matches_linear(workloads.sphere_grid_prims(orc, 4).commit())


def test_spatial_splits_match_linear_scan(orc):  # bvh.rs:1524-1527
    _assert_matches_linear(_shards_scene(orc, 64))


def test_spatial_splits_duplicate_references(orc):  # bvh.rs:1532-1541
    s = _shards_scene(orc, 64)
    assert len(s.tree()["refs"]) > s.primitive_count()


def test_triangles_are_packed_into_simd_lanes(orc):  # bvh.rs:1545-1566
    t = _shards_scene(orc, 64).tree()
    assert t["packets"] > 0 and t["indices"] == 0
    t = workloads.sphere_grid_prims(orc, 4).commit().tree()
    assert t["packets"] == 0 and t["indices"] == len(t["refs"])
    b = orc.SceneBuilder()
    b.attach(Geometry.TriangleMesh(*workloads.diagonal_shards(16)))
    for x in range(2):
        for y in range(2):
            for z in range(2):
                b.attach(Geometry.Sphere(np.array([x, y, z], np.float32) * 3.0, 0.5))
    s = b.commit()
    t = s.tree()
    assert t["packets"] > 0 and t["indices"] > 0 and len(t["refs"]) >= s.primitive_count()


def test_packets_are_well_filled(orc):  # bvh.rs:1570-1579
    t = _shards_scene(orc, 256).tree()
    assert t["packet_lanes"] / t["packets"] >= 2.5


def test_hit_any_matches_hit(orc):  # bvh.rs:1582-1607
    scene = workloads.sphere_grid_prims(orc, 4).commit()
    for o in [(-5.0, 4.5, 4.5), (20.0, 3.0, 3.0), (4.5, 4.5, 4.5)]:
        for d in [(1.0, 0.0, 0.0), _norm((-1.0, 0.05, 0.02)), _norm((0.3, 0.3, 1.0))]:
            for t_max in (0.5, 3.0, INF):
                ray = Ray(o, d)
                assert scene.occluded(ray, 0.001, t_max) == (scene.intersect(ray, 0.001, t_max) is not None)


def test_wide_node_is_two_cache_lines(orc):  # bvh.rs:1612-1616
    assert orc.lib().orc_sizeof_wide_node() == 128


def test_build_is_deterministic(orc):  # bvh.rs:1620-1638 — also across thread counts (rayon::join equivalence)
    a = workloads.sphere_grid_prims(orc, 6).commit(host_threads=1).tree()
    b = workloads.sphere_grid_prims(orc, 6).commit(host_threads=8).tree()
    for k in ("nodes", "tri_counts", "other_counts", "refs"):
        assert np.array_equal(a[k], b[k])
    s1 = orc.SceneBuilder()
    s1.attach(Geometry.TriangleMesh(*workloads.uv_sphere((0, 0, 0), 1.0, 160, 80)))
    s2 = orc.SceneBuilder()
    s2.attach(Geometry.TriangleMesh(*workloads.uv_sphere((0, 0, 0), 1.0, 160, 80)))
    a, b = s1.commit(host_threads=1).tree(), s2.commit(host_threads=8).tree()
    for k in ("nodes", "tri_counts", "other_counts", "refs"):
        assert np.array_equal(a[k], b[k])


def test_collapse_widens_the_tree(orc):  # bvh.rs:1642-1660
    t = workloads.sphere_grid_prims(orc, 6).commit().tree()
    n_leaf_slots = sum(bin((int(f) >> 4) & 0xF).count("1") for f in t["nodes"][:, 28])
    assert n_leaf_slots > 0 and n_leaf_slots == len(t["tri_counts"])
    assert len(t["nodes"]) * 2 < max(n_leaf_slots, 2) * 2 - 1


def test_eight_wide_packets_would_not_reduce_vector_rounds(orc):  # bvh.rs:1737-1765
    f = np.float32
    segs, rings = 80, 40
    r = np.arange(rings + 1, dtype=f)
    s = np.arange(segs + 1, dtype=f)
    phi = (r / f(rings)) * f(np.pi)
    th = (s / f(segs)) * f(2 * np.pi)
    v = np.stack([np.outer(np.sin(phi, dtype=f), np.cos(th, dtype=f)), np.outer(np.cos(phi, dtype=f), np.ones(segs + 1, f)),
                  np.outer(np.sin(phi, dtype=f), np.sin(th, dtype=f))], -1).reshape(-1, 3).astype(f)
    _, idx = workloads.uv_sphere((0, 0, 0), 1.0, segs, rings)
    b = orc.SceneBuilder()
    b.attach(Geometry.TriangleMesh(v, idx))
    per_leaf = b.commit().tree()["tri_counts"].astype(np.int64)
    assert per_leaf.max() > 0 and per_leaf.max() <= 4
    assert ((per_leaf + 3) // 4).sum() == ((per_leaf + 7) // 8).sum()
