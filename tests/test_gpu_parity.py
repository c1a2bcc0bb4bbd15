"""-m gpu: CUDA traversal vs the CPU oracle on identical seeded ray batches.

Bar (BASELINE.json north_star): per-ray geom_id/prim_id agreement >= 99.99 %,
hit t within 1e-5 relative. Because the tree, the visiting order and the f32
arithmetic are reproduced exactly, these tests demand the stronger property the
design aims for — bit-identical (t, u, v, ids) — and report the weaker one on
failure.
"""
import numpy as np
import pytest

from crust_render_b200 import workloads
from crust_render_b200.rt_types import (MASK_CAMERA, MASK_SHADOW, Affine3A, CubicCurveSegment, CurveSegment, Geometry,
                                        make_rays)

pytestmark = pytest.mark.gpu
INF = np.float32(np.inf)


def _compare_closest(p_scene, o_scene, rays, time=None, mask=None, exact=True, normals_tol=1e-5):
    ph, pf = p_scene.intersect_n(rays, time, mask)
    oh, of = o_scene.intersect_n(rays, time, mask, threads=0)
    n = len(rays)
    ids_ok = (ph["geom_id"] == oh["geom_id"]) & (ph["prim_id"] == oh["prim_id"])
    agree = ids_ok.mean()
    both = (ph["geom_id"] != 0xFFFFFFFF) & (oh["geom_id"] != 0xFFFFFFFF)
    rel = np.abs(ph["t"][both] - oh["t"][both]) / np.maximum(np.abs(oh["t"][both]), 1e-30)
    assert agree >= 0.9999, f"id agreement {agree:.6f} over {n} rays"
    assert rel.size == 0 or rel.max() <= 1e-5, f"max relative t error {rel.max():.3e}"
    if exact:
        assert np.array_equal(ph["geom_id"], oh["geom_id"]) and np.array_equal(ph["prim_id"], oh["prim_id"])
        for k in ("t", "u", "v"):
            assert ph[k].tobytes() == oh[k].tobytes(), f"{k} not bit-identical"
        assert np.array_equal(pf, of)
    m = both & ids_ok
    assert np.abs(ph["normal"][m] - oh["normal"][m]).max(initial=0.0) <= normals_tol
    return int(both.sum())


def _compare_occluded(p_scene, o_scene, rays, time=None, mask=None):
    a = p_scene.occluded_n(rays, time, mask)
    b = o_scene.occluded_n(rays, time, mask, threads=0)
    assert np.array_equal(a, b), f"{int((a != b).sum())} of {len(rays)} occlusion answers differ"
    return int(a.sum())


def _pair(gpu, orc, make):
    return make(gpu).commit(), make(orc).commit()


def test_reference_probe_batches(gpu, orc):
    """ray_throughput.rs: the three probe scenes x the 4096-ray LCG batch, closest + occluded."""
    for make, extent in ((workloads.triangle_scene, 6.0), (workloads.sphere_grid_scene, 14.0), (workloads.instance_scene, 7.0)):
        p, o = _pair(gpu, orc, make)
        rays = workloads.ray_batch(4096, extent)
        hits = _compare_closest(p, o, rays)
        assert hits > 100
        assert _compare_occluded(p, o, rays) == hits  # hit_any == hit.is_some() (bvh.rs:1583-1607)


def test_tri_spheres_one_million_rays(gpu, orc):
    p, o = _pair(gpu, orc, workloads.triangle_scene)
    rays = workloads.ray_batch(1 << 20, 6.0, seed=0x1234567)
    assert _compare_closest(p, o, rays) > 10000
    _compare_occluded(p, o, rays)
    # finite ranges: occlusion with t_max short of / beyond the first hit
    rays["tmax"] = np.float32(9.0)
    _compare_closest(p, o, rays)
    _compare_occluded(p, o, rays)


def test_mixed_t_min_signs_and_all_shear_frames(gpu, orc):
    """Both push-order paths of the closest-hit node visit in the same warps — the unsigned-key ranks need t_min >= +0, rays
    with a negative (or -0.0) t_min take the float / class ranks — and every one of the six (kz, sign d[kz]) shear frames of the
    triangle test, whose hot path leaves out the reference's kx / ky swap (traverse.cuh tri_eval): bit-identical to the oracle."""
    for make, extent in ((workloads.triangle_scene, 6.0), (workloads.instance_scene, 7.0)):
        p, o = _pair(gpu, orc, make)
        rays = workloads.ray_batch(1 << 18, extent, seed=0xBEEF)
        rng = np.random.default_rng(11)
        n = len(rays)
        rays["tmin"] = rng.choice(np.array([-50.0, -1.0, -0.0, 0.0, 1e-3, 2.0], np.float32), n)  # negative: hits behind the origin count
        assert _compare_closest(p, o, rays) > 5000
        _compare_occluded(p, o, rays)
        # dominant axis and its sign forced per ray: the six frames in equal shares, neighbours in a warp differ
        d = rays["d"].copy()
        k = np.arange(n) % 3
        sgn = np.where((np.arange(n) // 3) % 2 == 0, 1.0, -1.0).astype(np.float32)
        d[np.arange(n), k] = (np.abs(d).max(1) * np.float32(1.5) + np.float32(0.1)) * sgn
        rays["d"] = d
        _compare_closest(p, o, rays)
        _compare_occluded(p, o, rays)


def _copies_scene(api, top_level_triangles: bool):
    """One 576-triangle sphere instanced 64 times one level deep: translated copies (identity linear part — the child keeps the
    parent's direction-derived ray setup, traverse.cu), rotations by 0 (cos = 1, sin = 0: identity values reached through a
    product), true rotations, scales, a moving copy — interleaved so neighbouring rays of a warp take different paths."""
    inner = api.SceneBuilder()
    v, i = workloads.uv_sphere(np.zeros(3, np.float32), 1.0, 24, 12)
    inner.attach(Geometry.TriangleMesh(v, i))
    inner = inner.commit()
    b = api.SceneBuilder()
    k = 0
    for x in range(4):
        for y in range(4):
            for z in range(4):
                t = Affine3A.from_translation(np.array([x, y, z], np.float32) * np.float32(2.5))
                kind = k % 5
                if kind == 0:
                    xf, end = t, None
                elif kind == 1:
                    xf, end = t * Affine3A.from_rotation_y(0.0), None
                elif kind == 2:
                    xf, end = t * Affine3A.from_rotation_z(0.3 + 0.1 * k), None
                elif kind == 3:
                    xf, end = t * Affine3A.from_scale(np.array([1.0, 0.5, 1.5], np.float32)), None
                else:
                    xf, end = t, Affine3A.from_translation(np.array([x, y, z], np.float32) * np.float32(2.5) + np.float32(0.4))
                b.attach(Geometry.Instance(inner, xf, end))
                k += 1
    if top_level_triangles:
        fv = np.array([[-3, -1.5, -3], [12, -1.5, -3], [12, -1.5, 12], [-3, -1.5, 12]], np.float32)
        b.attach(Geometry.TriangleMesh(fv, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)))
    return b


@pytest.mark.parametrize("top_level_triangles", [False, True])
def test_translated_copies_next_to_rotated_and_moving_instances(gpu, orc, top_level_triangles):
    p, o = _pair(gpu, orc, lambda api: _copies_scene(api, top_level_triangles))
    rays = workloads.ray_batch(1 << 18, 11.0, seed=0xC0B1E5)
    rays["o"] += np.float32(3.75)
    time = np.random.default_rng(3).uniform(0.0, 1.0, len(rays)).astype(np.float32)
    time[::3] = 0.0  # the reference takes the cached transform at time 0 (prim.rs:345-350)
    assert _compare_closest(p, o, rays, time=time) > 20000
    _compare_occluded(p, o, rays, time=time)
    d = rays["d"].copy()
    d[::2, 1] = np.float32(-0.0)  # signed zeros in the direction: the copy's direction is compared bit for bit
    d[1::4, 0] = np.float32(0.0)
    rays["d"] = d
    _compare_closest(p, o, rays, time=time)
    _compare_occluded(p, o, rays, time=time)


def test_stress_scene_with_spatial_splits(gpu, orc):
    """gen_stress_scene.py at 1/9 size: flat spheres + thin shards (SBVH duplicates) + floor + slab."""
    make = lambda api: workloads.stress_scene(api, grid=2, shards=30000)[0]  # noqa: E731
    p, o = _pair(gpu, orc, make)
    ext = workloads.scene_extent(p.bounds())
    rays = workloads.ray_batch(1 << 19, ext, seed=99)
    assert _compare_closest(p, o, rays) > 1000
    _compare_occluded(p, o, rays)
    # rays started on the floor going up (bounce-like, t_min = 0.001), and axis-aligned ones
    rng = np.random.default_rng(3)
    n = 1 << 18
    org = np.stack([rng.uniform(-16, 16, n), np.zeros(n), rng.uniform(-16, 16, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 1] = np.abs(d[:, 1])
    d[: n // 8] = (0, 1, 0)
    rays = make_rays(org, d, 0.001, np.inf)
    _compare_closest(p, o, rays)
    _compare_occluded(p, o, rays)


def test_watertight_edges_and_vertices_on_gpu(gpu, orc):
    """Rays aimed exactly at mesh vertices and edge midpoints: the f64 tie-break lanes
    (triangle.rs:141-145, bvh.rs:551-561) must resolve identically. (The reference's slab
    test is not conservative — bvh.rs:790-808 has no ulp padding — so a ray through a vertex
    that is also a box corner can miss in the BVH while a linear scan hits; the oracle
    reproduces that, and the GPU must reproduce the oracle.)"""
    v, i = workloads.uv_sphere((0, 0, 0), 1.0, 40, 20)
    make = lambda api: (lambda b: (b.attach(Geometry.TriangleMesh(v, i)), b)[1])(api.SceneBuilder())  # noqa: E731
    p, o = _pair(gpu, orc, make)
    tri = v[i]  # [m,3,3]
    targets = np.concatenate([tri[:, 0], (tri[:, 0] + tri[:, 1]) * np.float32(0.5), (tri[:, 1] + tri[:, 2]) * np.float32(0.5)])
    for origin in ((0.0, 0.0, 0.0), (3.0, 0.5, -2.0)):
        org = np.tile(np.asarray(origin, np.float32), (len(targets), 1))
        rays = make_rays(org, targets - org, 0.001, np.inf)
        hits = _compare_closest(p, o, rays)
        if origin == (0.0, 0.0, 0.0):
            assert hits >= 0.98 * len(rays)


def test_all_primitive_kinds_masks_motion_nesting(gpu, orc):
    rng = np.random.default_rng(11)
    cache = {}

    def make(api):
        r = np.random.default_rng(5)
        ball = api.SceneBuilder()
        ball.attach(Geometry.TriangleMesh(*workloads.uv_sphere((0, 0, 0), 1.0, 24, 12)))
        ball.attach_masked(Geometry.Sphere((0, 1.5, 0), 0.5), MASK_CAMERA | MASK_SHADOW)
        ball = ball.commit()
        hair = api.SceneBuilder()
        hair.attach(Geometry.RoundCurves([CurveSegment(r.uniform(-1, 1, 3), r.uniform(-1, 1, 3), 0.08, 0.02) for _ in range(60)]))
        hair.attach(Geometry.CubicCurves([CubicCurveSegment(r.uniform(-1.5, 1.5, (4, 3)), 0.06, 0.03) for _ in range(40)]))
        hair = hair.commit()
        mid = api.SceneBuilder()
        mid.attach(Geometry.Instance(ball, Affine3A.from_translation((-2, 0, 0))))
        mid.attach(Geometry.Instance(hair, Affine3A.from_translation((2, 0, 0)) * Affine3A.from_scale((1, 2, 1))))
        mid.attach_masked(Geometry.Instance(ball, Affine3A.from_translation((0, 3, 0))), MASK_SHADOW)
        mid = mid.commit()
        cache[api.__name__] = (ball, hair, mid)
        b = api.SceneBuilder()
        b.attach(Geometry.TriangleMesh([(-30, -3, -30), (30, -3, -30), (30, -3, 30), (-30, -3, 30)], [(0, 1, 2), (0, 2, 3)]))
        for k in range(12):
            rot = Affine3A.from_rotation_y(0.5 * k) * Affine3A.from_rotation_x(0.1 * k)
            t = Affine3A.from_translation((6.0 * (k % 4) - 9.0, 0.0, 6.0 * (k // 4) - 6.0)) * rot
            end = Affine3A.from_translation((1.0, 0.5, 0.0)) * t if k % 3 == 0 else None
            b.attach(Geometry.Instance(mid, t, end))
        for k in range(20):
            b.attach(Geometry.Sphere(r.uniform(-12, 12, 3), float(r.uniform(0.2, 1.0))))
        b.attach(Geometry.RoundCurves([CurveSegment(r.uniform(-12, 12, 3), r.uniform(-12, 12, 3), 0.2, 0.1) for _ in range(30)]))
        return b

    p, o = _pair(gpu, orc, make)
    assert p.has_motion() and o.has_motion()
    n = 1 << 18
    rays = workloads.ray_batch(n, 16.0, seed=2024)
    time = rng.uniform(0, 1, n).astype(np.float32)
    time[: n // 4] = 0.0
    mask = rng.choice(np.array([0xFFFFFFFF, MASK_CAMERA, MASK_SHADOW, 4], np.uint32), n)
    # curves / instance normals go through sqrt+div chains: ids and t must still be exact
    assert _compare_closest(p, o, rays, time, mask) > 5000
    _compare_occluded(p, o, rays, time, mask)


def _deep_stack_scene(api, levels=8, quads=32768):
    """Eight nested levels; each holds `quads` parallel quads stacked along +z (all crossed by a ray along the z axis) and,
    squeezed in front of the first quad, an instance of the next level scaled down to the quad spacing — so the instance
    sits at the bottom of its level's tree and every level leaves its ~3 far siblings per BVH4 level on the traversal
    stack when the ray descends: about 8 x 25 = 200 live entries (a build with the capacity forced back to 128 fails this test: CRT_STACK_CAP_ALL), more than the 128 the first kernel kept."""
    dz = np.float32(0.01)
    k = np.arange(1, quads + 1, dtype=np.float32)
    z = np.float32(10.0) + k * dz
    corners = np.array([[-1, -2], [3, -2], [3, 2], [-1, 2]], np.float32)  # the z axis crosses the interior of the first triangle
    v = np.zeros((quads, 4, 3), np.float32)
    v[:, :, :2] = corners
    v[:, :, 2] = z[:, None]
    idx = (np.arange(quads, dtype=np.uint32)[:, None] * 4 + np.array([0, 1, 2, 0, 2, 3], np.uint32)).reshape(-1, 3)
    s = np.float32(dz / (z[-1] - np.float32(10.0) + dz))  # the child's z extent -> one quad spacing
    child = None
    for _ in range(levels):
        b = api.SceneBuilder()
        b.attach(api.Geometry.TriangleMesh(v.reshape(-1, 3), idx))
        if child is not None:
            xf = Affine3A([s, 0, 0, 0, s, 0, 0, 0, s, 0, 0, float(np.float32(10.0) - np.float32(10.0) * s)])
            b.attach(api.Geometry.Instance(child, xf))
        child = b.commit()
    return child


def test_traversal_stack_deeper_than_128_entries(gpu, orc):
    """crust-rt's traversal stack spills to the heap and never drops an entry (bvh.rs:711-755). The device stack is
    bounded, but by more than any committed tree can need (traverse.cu StackCap): a scene built to keep ~150 entries
    alive must give the oracle's hits bit for bit and leave the sticky overflow bit clear."""
    p, o = _deep_stack_scene(gpu), _deep_stack_scene(orc)
    n = 2048
    rng = np.random.default_rng(5)
    org = np.zeros((n, 3), np.float32)
    org[:, 2] = -5.0
    d = np.zeros((n, 3), np.float32)
    d[:, 2] = 1.0
    d[n // 2:, :2] = rng.normal(scale=0.02, size=(n - n // 2, 2)).astype(np.float32)  # half the batch leaves the axis
    rays = make_rays(org, d, 0.001, INF)
    hits = _compare_closest(p, o, rays)
    assert hits == n
    assert _compare_occluded(p, o, rays) == n
    assert gpu.device_status() == 0


def test_ragged_and_empty_batches(gpu, orc):
    p, o = _pair(gpu, orc, workloads.sphere_grid_scene)
    for n in (0, 1, 31, 33, 127, 129, 4097):
        rays = workloads.ray_batch(max(n, 1), 14.0)[:n]
        ph, _ = p.intersect_n(rays)
        oh, _ = o.intersect_n(rays, threads=1)
        assert ph.tobytes() == oh.tobytes()
    empty = gpu.SceneBuilder().commit()
    h, _ = empty.intersect_n(workloads.ray_batch(100, 1.0))
    assert np.all(h["geom_id"] == 0xFFFFFFFF) and np.all(np.isinf(h["t"]))
    assert not empty.occluded_n(workloads.ray_batch(100, 1.0)).any()


def test_device_pointer_entry_points(gpu, orc):
    """crt_intersect_n / crt_occluded_n with DEVICE buffers on a torch stream."""
    import torch

    p, o = _pair(gpu, orc, workloads.triangle_scene)
    rays = workloads.ray_batch(1 << 16, 6.0, seed=7)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.empty((len(rays), 8), dtype=torch.float32, device="cuda")
    d_ff = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    d_occ = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    before = gpu.launch_count()
    with torch.cuda.stream(s):
        p.intersect_ptr(len(rays), d_rays.data_ptr(), 0, 0, d_hits.data_ptr(), d_ff.data_ptr(), s.cuda_stream)
        p.occluded_ptr(len(rays), d_rays.data_ptr(), 0, 0, d_occ.data_ptr(), s.cuda_stream)
    s.synchronize()
    assert gpu.launch_count() == before + 2
    oh, of = o.intersect_n(rays, threads=0)
    got = d_hits.cpu().numpy().view(gpu.HIT_DTYPE).reshape(-1)
    assert got.tobytes() == oh.tobytes()
    assert np.array_equal(d_ff.cpu().numpy(), of)
    assert np.array_equal(d_occ.cpu().numpy(), o.occluded_n(rays, threads=0))
    assert gpu.device_status() == 0


def test_full_size_k1_mesh_properties(gpu, orc):
    """BASELINE's 10 M-triangle stress mesh (K1) at full size: the oracle checks a strided sample, and the
    whole batch is checked through size-independent properties — bit determinism, invariance under a permutation of the
    rays (persistent warps fetch rays dynamically, so this exercises the refill / phase-balancing scheduling),
    hit_any == hit.is_some() (bvh.rs:1583-1607), a finite range equal to the open query's answer, and the hit point lying on
    the reported triangle's plane."""
    b_gpu, n_tris = workloads.stress_scene(gpu, grid=32, shards=30000, jitter=True)
    assert n_tris > 10_000_000
    p = b_gpu.commit()
    ext = workloads.scene_extent(p.bounds())
    n = 1 << 22
    rays = workloads.ray_batch(n, ext, seed=0xC0FFEE)
    h, ff = p.intersect_n(rays)
    h2, ff2 = p.intersect_n(rays)
    assert h.tobytes() == h2.tobytes() and ff.tobytes() == ff2.tobytes()
    perm = np.random.default_rng(11).permutation(n)
    hp, ffp = p.intersect_n(rays[perm])
    assert hp.tobytes() == h[perm].tobytes() and ffp.tobytes() == ff[perm].tobytes()
    hit = h["geom_id"] != 0xFFFFFFFF
    assert 0.2 < hit.mean() < 0.99
    occ = p.occluded_n(rays)
    assert np.array_equal(occ.astype(bool), hit)
    # a range that ends just beyond the found hit finds the same hit; one that ends just short of it finds nothing nearer
    r2 = rays.copy()
    r2["tmax"] = np.where(hit, h["t"] * np.float32(1.0 + 1e-6), np.float32(1.0))
    hb, _ = p.intersect_n(r2)
    same = hit & (hb["geom_id"] == h["geom_id"]) & (hb["prim_id"] == h["prim_id"]) & (hb["t"] == h["t"])
    assert same.sum() == hit.sum()
    r2["tmax"] = np.where(hit, h["t"] * np.float32(1.0 - 1e-5), np.float32(1.0))
    hc, _ = p.intersect_n(r2)
    assert not (hc["geom_id"][hit] != 0xFFFFFFFF).any() or (hc["t"][hit][hc["geom_id"][hit] != 0xFFFFFFFF] < h["t"][hit][hc["geom_id"][hit] != 0xFFFFFFFF]).all()
    assert (hc["geom_id"][hit] != 0xFFFFFFFF).mean() < 1e-3  # closest means closest: (almost) nothing lies in front of it
    # the oracle on a strided sample of the same batch
    b_orc, _ = workloads.stress_scene(orc, grid=32, shards=30000, jitter=True)
    o = b_orc.commit()
    sub = rays[:: n // (1 << 16)]
    oh, of = o.intersect_n(sub, threads=0)
    ph = h[:: n // (1 << 16)]
    assert ph.tobytes() == oh.tobytes() and np.array_equal(ff[:: n // (1 << 16)], of)
    assert np.array_equal(occ[:: n // (1 << 16)], o.occluded_n(sub, threads=0))
    # ... and the five ray sets of bench.py (LCG box rays, camera rays, cosine bounce rays, RectLight shadow rays)
    # at 2^22 rays each, every ray against the oracle, bit for bit
    import bench
    sets, _ = bench.build_ray_sets(workloads, p.bounds(), 22, 0, lambda r: p.intersect_n(r)[0])
    for name in bench.SETS:
        set_rays, set_mask = sets[name]
        assert len(set_rays) >= (1 << 22) * 0.9
        if name.startswith("closest"):
            gh, gf = p.intersect_n(set_rays, None, set_mask)
            ch, cf = o.intersect_n(set_rays, None, set_mask, threads=0)
            assert gh.tobytes() == ch.tobytes() and np.array_equal(gf, cf), name
        else:
            assert np.array_equal(p.occluded_n(set_rays, None, set_mask), o.occluded_n(set_rays, None, set_mask, threads=0)), name


def test_l2_window_over_breadth_first_top_levels(gpu, orc):
    """BASELINE north_star: "L2-persistence hints on top-level nodes". A tree larger than the L2 set-aside gets its top levels
    re-ordered breadth-first at the head of the device node array (child words renumbered; the host tree keeps the reference's
    DFS order) and every launch attaches an access-policy window over them: results stay bit-identical to the oracle."""
    set_aside, applied0, _ = gpu.l2_persist_info()
    assert set_aside > 0
    # 383 k triangles: the whole 8 MB node array fits the window; the re-ordered case is the 215 MB tree of
    # test_full_size_k1_mesh_properties
    make = lambda api: workloads.stress_scene(api, grid=6, shards=30000)[0]  # noqa: E731
    p, o = _pair(gpu, orc, make)
    bi = p.build_info()
    rays = workloads.ray_batch(1 << 18, workloads.scene_extent(p.bounds()), seed=5)
    _compare_closest(p, o, rays)
    _compare_occluded(p, o, rays)
    _, applied1, refused = gpu.l2_persist_info()
    assert applied1 > applied0 and refused == 0, (applied0, applied1, refused, bi.node_bytes)


def test_device_traversal_counters_match_the_reference_counters(gpu, orc):
    """The counting build of the kernels (crt_trace_counters_*) against the oracle's restatement of the reference's `stats`
    counters (bvh.rs:31-136) on the same batch. The device culls a popped entry whose stored entry distance already exceeds
    `closest` without fetching the node; the reference fetches it and finds nothing — so nodes fetched + entries culled
    equals the reference's nodes popped, and leaves / triangle lanes / instance descents are equal one to one."""
    for make, extent in ((workloads.triangle_scene, 6.0), (workloads.instance_scene, 7.0)):
        p, o = _pair(gpu, orc, make)
        rays = workloads.ray_batch(1 << 16, extent, seed=99)
        _, _, st = o.intersect_n(rays, threads=0, stats=True)
        gpu.trace_counters_read(reset=True)
        gpu.trace_counters_enable(True)
        try:
            hits, _ = p.intersect_n(rays)
            c = gpu.trace_counters_read(reset=True)
            occ = p.occluded_n(rays)
            ca = gpu.trace_counters_read(reset=True)
        finally:
            gpu.trace_counters_enable(False)
        assert c["rays"] == len(rays) and ca["rays"] == len(rays)
        assert c["nodes_fetched"] + c["entries_culled_on_pop"] == st["nodes_top"] + st["nodes_inst"]
        assert c["nodes_fetched"] < st["nodes_top"] + st["nodes_inst"]  # culling on pop does save fetches
        assert c["leaves_entered"] == st["leaves_top"] + st["leaves_inst"]
        assert c["triangle_records"] == st["tri_lanes"]
        assert c["instance_descents"] <= st["instance_tests"] and c["instance_descents"] + c["other_records"] >= st["instance_tests"]
        assert 0 < ca["nodes_fetched"] <= c["nodes_fetched"] + c["entries_culled_on_pop"]  # any-hit stops at the first hit
        # counting launches give the same answers as the plain ones
        h2, _ = p.intersect_n(rays)
        assert h2.tobytes() == hits.tobytes() and np.array_equal(p.occluded_n(rays), occ)


def test_scene_export_import_round_trip(gpu, orc):
    """crt_scene_export / crt_scene_import (build once, share across GPUs): a committed scene rebuilt from its blob — through
    host memory and through device memory — answers every query bit for bit like the original, for a flat SBVH scene and for
    an instancing hierarchy (children imported first and handed in by index), and reports the same metadata."""
    import torch

    def clone(scene, children=(), device_blob=False):
        n = scene.export_size()
        if device_blob:
            blob = torch.empty(n, dtype=torch.uint8, device="cuda")
            scene.export_to(blob.data_ptr(), n, True, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            return gpu.Scene.import_from(blob.data_ptr(), n, True, children)
        blob = np.empty(n, dtype=np.uint8)
        scene.export_to(blob.ctypes.data, n, False)
        return gpu.Scene.import_from(blob.ctypes.data, n, False, children)

    # flat, with spatial-split duplicates and the breadth-first top levels
    flat = workloads.stress_scene(gpu, grid=3, shards=4000)[0].commit()
    rays = workloads.ray_batch(1 << 17, workloads.scene_extent(flat.bounds()), seed=3)
    h0, f0 = flat.intersect_n(rays)
    for dev in (False, True):
        c = clone(flat, device_blob=dev)
        h1, f1 = c.intersect_n(rays)
        assert h1.tobytes() == h0.tobytes() and np.array_equal(f0, f1)
        assert np.array_equal(c.occluded_n(rays), flat.occluded_n(rays))
        assert c.primitive_count() == flat.primitive_count() and c.geometry_count() == flat.geometry_count()
        assert np.array_equal(c.bounds().minimum, flat.bounds().minimum) and c.memory_footprint().total() == flat.memory_footprint().total()
    # instancing: two levels, a moving instance, spheres next to meshes; the imported parent points at the imported children
    inner = gpu.SceneBuilder()
    v, i = workloads.uv_sphere(np.zeros(3, np.float32), 1.0, 16, 8)
    inner.attach(Geometry.TriangleMesh(v, i))
    inner.attach(Geometry.Sphere((0.0, 1.5, 0.0), 0.4))
    inner = inner.commit()
    mid = gpu.SceneBuilder()
    mid.attach(Geometry.Instance(inner, Affine3A.from_translation((-1.5, 0, 0))))
    mid.attach(Geometry.Instance(inner, Affine3A.from_translation((1.5, 0, 0)), Affine3A.from_translation((1.5, 1.0, 0))))
    mid = mid.commit()
    top = gpu.SceneBuilder()
    for k in range(5):
        top.attach(Geometry.Instance(mid, Affine3A.from_translation((0.0, 0.0, 4.0 * (k - 2)))))
    top.attach(Geometry.Instance(inner, Affine3A.from_translation((0.0, 4.0, 0.0))))
    top.attach(Geometry.Sphere((0.0, -3.0, 0.0), 1.0))
    top = top.commit()
    rays = workloads.ray_batch(1 << 16, 9.0, seed=4)
    time = np.random.default_rng(1).random(len(rays)).astype(np.float32)
    h0, f0 = top.intersect_n(rays, time)
    assert (h0["geom_id"] != 0xFFFFFFFF).mean() > 0.05
    inner2 = clone(inner, device_blob=True)
    mid2 = clone(mid, [inner2])
    top2 = clone(top, [mid2, inner2], device_blob=True)
    del inner, mid  # the clones must not depend on the originals' device memory
    h1, f1 = top2.intersect_n(rays, time)
    assert h1.tobytes() == h0.tobytes() and np.array_equal(f0, f1)
    assert np.array_equal(top2.occluded_n(rays, time), top.occluded_n(rays, time))
    assert top2.has_motion() and top2.unique_primitive_breakdown().triangles == top.unique_primitive_breakdown().triangles
    with pytest.raises(gpu.CrtError):
        clone(top, [mid2])  # wrong number of children
