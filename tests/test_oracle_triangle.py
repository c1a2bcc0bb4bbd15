"""Pins oracle/ against the reference's known-answer tests for the Woop
watertight triangle path: crates/crust-rt/src/triangle.rs:431-699."""
import numpy as np

from crust_render_b200.rt_types import MASK_ALL, MASK_CAMERA, MASK_SHADOW

INF = float("inf")


def test_interior_hit_has_expected_t_and_barycentrics(orc):  # triangle.rs:440-453
    r = orc.triangle_intersect((2, 1, 0), (0, 0, 1), (0, 0, 5), (4, 0, 5), (0, 4, 5), 0.001, INF)
    assert r is not None
    t, u, v = r
    assert abs(t - 5.0) < 1e-5 and abs(u - 0.5) < 1e-5 and abs(v - 0.25) < 1e-5


def test_respects_t_range(orc):  # triangle.rs:456-469
    tri = ((-1, -1, 5), (1, -1, 5), (0, 1, 5))
    o, z = (0, 0, 0), (0, 0, 1)
    assert orc.triangle_intersect(o, z, *tri, 0.001, 4.9) is None
    assert orc.triangle_intersect(o, z, *tri, 5.1, 100.0) is None
    assert orc.triangle_intersect(o, z, *tri, 0.001, INF) is not None
    assert orc.triangle_intersect(o, (0, 0, -1), *tri, 0.001, INF) is None


def test_shared_edge_is_watertight(orc):  # triangle.rs:476-497 — 10 001 rays across the shared diagonal
    f = np.float32
    p00, p10 = np.array([-1.3371, -0.7713, 3.7], f), np.array([1.9241, -1.1157, 4.3], f)
    p11, p01 = np.array([1.6083, 1.4127, 3.9], f), np.array([-0.9743, 0.8291, 4.1], f)
    tris = [(p00, p10, p11), (p00, p11, p01)]
    origin = np.array([0.1731, -0.0913, 0.0], f)
    for i in range(10001):
        s = f(i) / f(10000.0)
        target = p00 + s * (p11 - p00)
        d = target - origin
        hits = sum(orc.triangle_intersect(origin, d, *t, 0.001, INF) is not None for t in tris)
        assert hits >= 1, f"pinhole at s={s}"


def test_shared_vertex_is_covered(orc):  # triangle.rs:501-518
    hub = (0.2137, 0.5391, 5.1)
    rim = [(1.3, 0.4, 5.0), (0.3, 1.7, 5.3), (-1.1, 0.6, 4.9), (-0.2, -1.2, 5.2)]
    hits = sum(orc.triangle_intersect((0, 0, 0), hub, hub, rim[k], rim[(k + 1) % 4], 0.001, INF) is not None for k in range(4))
    assert hits >= 1


def test_simd_matches_scalar_bitwise(orc):  # triangle.rs:525-594 — same LCG, same case list
    f = np.float32
    state = [0x12345678]

    def nxt():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xFFFFFFFF
        return f(f(state[0] >> 8) / f(1 << 24) - f(0.5))

    def v3():
        return np.array([nxt(), nxt(), nxt()], f)

    compared = hits = 0
    for _ in range(2000):
        tris = []
        for _i in range(4):
            base = v3() * f(4.0)
            tris.append((base + v3(), base + v3(), base + v3()))
        origin = v3() * f(6.0)
        aimed = nxt() > 0.0
        if aimed:
            pick = int(f((nxt() + f(0.5)) * f(4.0))) % 4
            v0, v1, v2 = tris[pick]
            a, b = nxt() + f(0.5), nxt() + f(0.5)
            if a + b > 1.0:
                a, b = f(1.0) - a, f(1.0) - b
            d = (v0 + (v1 - v0) * a + (v2 - v0) * b) - origin
        else:
            d = v3()
        l2 = f(f(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        if l2 < 1e-8:
            continue
        d = d * (f(1.0) / np.sqrt(l2, dtype=f))
        out = orc.tri4_intersect(origin, d, tris, [MASK_ALL] * 4, MASK_ALL, 0.001, INF)
        for lane, (v0, v1, v2) in enumerate(tris):
            scalar = orc.triangle_intersect(origin, d, v0, v1, v2, 0.001, INF)
            if out["fallback"] & (1 << lane):
                continue
            compared += 1
            simd_hit = bool(out["hits"] & (1 << lane))
            assert simd_hit == (scalar is not None), f"hit disagreement lane {lane}"
            if simd_hit:
                hits += 1
                got = np.array([out["t"][lane], out["u"][lane], out["v"][lane]], f)
                assert got.tobytes() == np.array(scalar, f).tobytes()
    assert compared > 7000 and hits > 50


def test_simd_respects_t_range_like_scalar(orc):  # triangle.rs:599-626
    a, b, c = (-1, -1, 5), (1, -1, 5), (0, 1, 5)
    cases = [(a, b, c), (a, c, b)]
    for t_min, t_max in [(0.001, 4.9), (5.1, 100.0), (0.001, INF), (4.9, 5.1)]:
        for dz in (1.0, -1.0):
            o = (0, 0, 0.0 if dz > 0 else 10.0)
            out = orc.tri4_intersect(o, (0, 0, dz), cases, [MASK_ALL] * 2, MASK_ALL, t_min, t_max)
            for lane, tri in enumerate(cases):
                sc = orc.triangle_intersect(o, (0, 0, dz), *tri, t_min, t_max)
                assert bool(out["hits"] & (1 << lane)) == (sc is not None)


def test_simd_inactive_lanes_never_hit(orc):  # triangle.rs:630-644
    out = orc.tri4_intersect((0, 0, 0), (0, 0, 1), [((-1, -1, 5), (1, -1, 5), (0, 1, 5))], [MASK_ALL], MASK_ALL, 0.001, INF)
    assert out["active"] == 0b0001 and out["hits"] == 0b0001 and out["fallback"] & ~1 == 0


def test_simd_masks_gate_lanes(orc):  # triangle.rs:648-681
    def tri(z):
        return ((-1, -1, z), (1, -1, z), (0, 1, z))

    cases, masks = [tri(5.0), tri(6.0), tri(7.0)], [MASK_CAMERA, MASK_SHADOW, MASK_ALL]
    run = lambda m: orc.tri4_intersect((0, 0, 0), (0, 0, 1), cases, masks, m, 0.001, INF)["hits"]  # noqa: E731
    assert run(MASK_CAMERA) == 0b101 and run(MASK_SHADOW) == 0b110 and run(MASK_ALL) == 0b111 and run(1 << 20) == 0b100


def test_axis_aligned_rays_hit(orc):  # triangle.rs:686-698
    tri = ((-1, -1, 2), (1, -1, 2), (0, 1, 2))
    for dz in (1.0, -1.0):
        r = orc.triangle_intersect((0, 0, 0.0 if dz > 0 else 4.0), (0, 0, dz), *tri, 0.001, INF)
        assert r is not None and abs(r[0] - 2.0) < 1e-5


def test_clip_triangle_aabb_shrinks_every_axis(orc):  # triangle.rs:361-429 behaviour (no reference test; sanity)
    v0, v1, v2 = (0, 0, 0), (10, 10, 10.2), (10, 10.3, 10)
    lo, hi = orc.clip_triangle_aabb(v0, v1, v2, 0, 2.0, 4.0)
    assert lo[0] == 2.0 and hi[0] == 4.0
    assert lo[1] >= 1.9 and hi[1] <= 4.2 and lo[2] >= 1.9 and hi[2] <= 4.2
    assert orc.clip_triangle_aabb(v0, v1, v2, 0, 11.0, 12.0) is None
