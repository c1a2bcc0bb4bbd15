"""-m gpu: the device-side BVH build (csrc/device_build.cu, CRT_COMMIT_DEVICE_BUILD) against the oracle.
The GPU-built tree is not the reference's SAH/SBVH tree, so the comparison is on query results: the closest hit
of every ray must be the oracle's (same t, u, v, ids bit for bit — the primitive tests are the same code), except
where two triangles tie at exactly equal t; occlusion must agree everywhere."""
import numpy as np
import pytest

from crust_render_b200 import workloads as W
from crust_render_b200.rt_types import Geometry, make_rays

pytestmark = pytest.mark.gpu


def commit_both(gpu, orc, meshes, mask=None):
    bg, bo = gpu.SceneBuilder(), orc.SceneBuilder()
    for k, (v, i) in enumerate(meshes):
        m = 0xFFFFFFFF if mask is None else mask[k]
        bg.attach_masked(Geometry.TriangleMesh(v, i), m)
        bo.attach_masked(Geometry.TriangleMesh(v, i), m)
    return bg.commit(flags=gpu.COMMIT_DEVICE_BUILD), bo.commit()


def agree(g, o):
    same = (g["geom_id"] == o["geom_id"]) & (g["prim_id"] == o["prim_id"]) & (g["t"] == o["t"]) & (g["u"] == o["u"]) & (g["v"] == o["v"])
    tie = (~same) & (g["t"] == o["t"]) & (g["geom_id"] != 0xFFFFFFFF) & (o["geom_id"] != 0xFFFFFFFF)
    return same, tie


def test_probe_scene_matches_the_oracle(gpu, orc):
    """ray_throughput.rs's tri_spheres (43 200 triangles) x 262 144 LCG rays."""
    meshes = [W.uv_sphere(c, 0.45, 24, 12) for c in [(x * 1.2, y * 1.2, z * 1.2) for x in range(-2, 3) for y in range(-1, 2) for z in range(-2, 3)]]
    sg, so = commit_both(gpu, orc, meshes)
    info = sg.build_info()
    assert info.built_on_device == 1 and info.primitives == sum(len(i) for _, i in meshes) and info.leaf_refs >= info.primitives
    assert info.leaves * 4 >= info.primitives and info.wide_nodes > 0
    rays = W.ray_batch(1 << 18, W.scene_extent(so.bounds()))
    hg, fg = sg.intersect_n(rays)
    ho, fo = so.intersect_n(rays, threads=0)
    same, tie = agree(hg, ho)
    assert (same | tie).all() and same.mean() >= 0.9999
    assert np.array_equal(fg[same], fo[same])
    assert np.array_equal(sg.occluded_n(rays), so.occluded_n(rays, threads=0))
    bg, bo = sg.bounds(), so.bounds()
    assert np.allclose(bg.minimum, bo.minimum) and np.allclose(bg.maximum, bo.maximum)


def test_every_triangle_is_reachable(gpu, orc):
    """One ray per triangle, aimed at its centroid from outside the convex mesh: the hit must be that triangle."""
    v, i = W.uv_sphere((0.0, 0.0, 0.0), 1.0, 96, 48)
    sg, _ = commit_both(gpu, orc, [(v, i)])
    tri = v[i.astype(np.int64)]
    c = tri.mean(axis=1)
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    keep = area > 1e-7  # the pole rows hold zero-area triangles
    d = -c / np.linalg.norm(c, axis=1, keepdims=True)
    rays = make_rays(c[keep] - 2.0 * d[keep], d[keep])
    h, _ = sg.intersect_n(rays)
    assert np.array_equal(h["prim_id"], np.nonzero(keep)[0].astype(np.uint32)) and np.all(h["geom_id"] == 0)


def test_stress_scene_with_masks_and_bad_indices(gpu, orc):
    meshes = W.stress_meshes(grid=2, shards=3000, jitter=True)
    v, i = meshes[0]
    bad = i.copy()
    bad[5] = (0, 1, 10 ** 6)            # out-of-range index: skipped, still consumes prim_id 5 (scene.rs:251-253)
    meshes[0] = (v, bad)
    masks = [0xFFFFFFFF] * len(meshes)
    masks[1] = 0x2                      # shadow rays only
    sg, so = commit_both(gpu, orc, meshes, masks)
    assert sg.build_info().built_on_device == 1 and sg.primitive_count() == so.primitive_count()
    rays = W.ray_batch(1 << 17, W.scene_extent(so.bounds()), seed=99)
    for m in (None, np.full(len(rays), 0x1, np.uint32), np.full(len(rays), 0x2, np.uint32)):
        hg, _ = sg.intersect_n(rays, None, m)
        ho, _ = so.intersect_n(rays, None, m, threads=0)
        same, tie = agree(hg, ho)
        assert (same | tie).all() and same.mean() >= 0.9999
        assert np.array_equal(sg.occluded_n(rays, None, m), so.occluded_n(rays, None, m, threads=0))


@pytest.mark.parametrize("n_tris", [0, 1, 3, 4, 5, 9])
def test_tiny_and_empty_scenes(gpu, orc, n_tris):
    rng = np.random.default_rng(n_tris)
    v = rng.uniform(-1, 1, (max(3 * n_tris, 3), 3)).astype(np.float32)
    i = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    sg, so = commit_both(gpu, orc, [(v, i)])
    rays = W.ray_batch(4096, 2.0, seed=5)
    hg, _ = sg.intersect_n(rays)
    ho, _ = so.intersect_n(rays, threads=0)
    same, tie = agree(hg, ho)
    assert (same | tie).all()
    assert np.array_equal(sg.occluded_n(rays), so.occluded_n(rays, threads=0))


def test_deterministic_and_falls_back_for_vertex_normals(gpu):
    meshes = W.stress_meshes(grid=2, shards=2000, jitter=False)
    rays = W.ray_batch(1 << 16, 30.0)
    outs = []
    for _ in range(2):
        b = gpu.SceneBuilder()
        for v, i in meshes:
            b.attach(Geometry.TriangleMesh(v, i))
        s = b.commit(flags=gpu.COMMIT_DEVICE_BUILD)
        bi = s.build_info()
        outs.append((s.intersect_n(rays)[0].tobytes(), bi.wide_nodes, bi.leaves))
    assert outs[0] == outs[1]
    # the one thing left to the host build: meshes with per-vertex shading normals
    b = gpu.SceneBuilder()
    v, i = meshes[0]
    b.attach(Geometry.TriangleMesh(v, i, normals=np.tile(np.array([[0, 1, 0]], np.float32), (len(v), 1))))
    b.attach(Geometry.Sphere((0, 0, 0), 1.0))
    s = b.commit(flags=gpu.COMMIT_DEVICE_BUILD)
    assert s.build_info().built_on_device == 0 and s.primitive_breakdown().spheres == 1


def _mixed_scene(api, flags=0):
    """Every primitive kind in one scene: two meshes, spheres, round curves, cubic spans, instances of a mesh BLAS (one of them
    moving), instances of a scene that itself instances (nesting 2), all with masks."""
    from crust_render_b200.rt_types import Affine3A, CubicCurveSegment, CurveSegment
    kw = {"flags": flags} if flags else {}
    blas = api.SceneBuilder()
    blas.attach(Geometry.TriangleMesh(*W.uv_sphere((0, 0, 0), 0.5, 16, 8)))
    blas.attach(Geometry.Sphere((0.0, 0.8, 0.0), 0.25))
    blas = blas.commit(**kw)
    mid = api.SceneBuilder()
    mid.attach(Geometry.Instance(blas, Affine3A.from_translation((-0.7, 0, 0))))
    mid.attach(Geometry.Instance(blas, Affine3A.from_translation((0.7, 0, 0))))
    mid = mid.commit(**kw)
    b = api.SceneBuilder()
    b.attach(Geometry.TriangleMesh(*W.uv_sphere((-3, 0, 0), 1.0, 24, 12)))
    b.attach_masked(Geometry.TriangleMesh(*W.uv_sphere((3, 0, 0), 1.0, 24, 12)), 0x5)
    rng = np.random.default_rng(3)
    for c in rng.uniform(-4, 4, (40, 3)):
        b.attach(Geometry.Sphere(c.astype(np.float32), 0.3))
    b.attach(Geometry.RoundCurves([CurveSegment((-4, -3, k * 0.5), (4, -3, k * 0.5 + 0.2), 0.08, 0.03) for k in range(-6, 7)]))
    b.attach(Geometry.CubicCurves([CubicCurveSegment(((-4, 3, k), (-1, 4.5, k), (1, 1.5, k), (4, 3, k)), 0.1, 0.05) for k in (-2.0, 0.0, 2.0)]))
    for k in range(12):
        t = (float(rng.uniform(-4, 4)), float(rng.uniform(-4, 4)), float(rng.uniform(-4, 4)))
        end = Affine3A.from_translation((t[0], t[1] + 0.8, t[2])) if k == 3 else None
        b.attach_masked(Geometry.Instance(blas, Affine3A.from_translation(t), end), 0xFFFFFFFF if k % 4 else 0x3)
    for k in range(4):
        b.attach(Geometry.Instance(mid, Affine3A.from_translation((0.0, 0.0, -4.5 + 3.0 * k)) * Affine3A.from_scale((1.0, 1.0 + 0.25 * k, 1.0))))
    return b.commit(**kw)


def test_every_primitive_kind_builds_on_the_device(gpu, orc):
    """Spheres, curve segments, cubic spans and (nested, moving, masked) instances go through the device build too: their
    boxes and leaf records are prepared on the host exactly as the host build prepares them and are sorted and grouped on
    the device together with the triangles. Closest hits are the oracle's except exact-t ties, occlusion is equal."""
    sg = _mixed_scene(gpu, gpu.COMMIT_DEVICE_BUILD)
    so = _mixed_scene(orc)
    bi = sg.build_info()
    pb, po = sg.primitive_breakdown(), so.primitive_breakdown()
    assert bi.built_on_device == 1
    assert (pb.triangles, pb.spheres, pb.curve_segments, pb.cubic_curve_spans, pb.instances) == (po.triangles, po.spheres, po.curve_segments, po.cubic_curve_spans, po.instances)
    assert sg.has_motion() and sg.geometry_count() == so.geometry_count() and sg.primitive_count() == so.primitive_count()
    assert np.allclose(sg.bounds().minimum, so.bounds().minimum) and np.allclose(sg.bounds().maximum, so.bounds().maximum)
    eg, eh = sg.primitive_extents(), _mixed_scene(gpu).primitive_extents()  # scene.rs:446-455: count, scene / mean / max diagonal
    assert eg[0] == eh[0] and abs(eg[1] - eh[1]) <= 1e-5 * eh[1] and abs(eg[2] - eh[2]) <= 1e-4 * eh[2] and abs(eg[3] - eh[3]) <= 1e-5 * eh[3]
    rays = W.ray_batch(1 << 17, 9.0, seed=21)
    time = np.random.default_rng(2).random(len(rays)).astype(np.float32)
    mask = np.where(np.arange(len(rays)) % 3 == 0, 0x1, 0xFFFFFFFF).astype(np.uint32)
    hg, fg = sg.intersect_n(rays, time, mask)
    ho, fo = so.intersect_n(rays, time, mask, threads=0)
    same, tie = agree(hg, ho)
    assert (same | tie).all() and same.mean() >= 0.9995 and (hg["geom_id"] != 0xFFFFFFFF).mean() > 0.3
    assert np.array_equal(fg[same], fo[same])
    assert np.abs(hg["normal"][same] - ho["normal"][same]).max() <= 1e-5
    assert np.array_equal(sg.occluded_n(rays, time, mask), so.occluded_n(rays, time, mask, threads=0))
    # a device-built instancing scene exports / imports like a host-built one
    import torch
    n = sg.export_size()
    blob = torch.empty(n, dtype=torch.uint8, device="cuda")
    sg.export_to(blob.data_ptr(), n, True, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    clone = gpu.Scene.import_from(blob.data_ptr(), n, True, sg._children)
    hc, _ = clone.intersect_n(rays, time, mask)
    assert hc.tobytes() == hg.tobytes()
    assert gpu.device_status() == 0


def test_forty_thousand_instance_tlas_builds_on_the_device(gpu, orc):
    """BASELINE config 4: the TLAS over PointInstancedMedCity's 40 000 instances (and its prototype BLASes) built on the device."""
    from crust_render_b200 import scenes
    import crust_render_b200.rt as rt_mod
    prev = rt_mod._commit_override
    rt_mod._commit_override = lambda b, flags, threads: b.commit_local(flags | gpu.COMMIT_DEVICE_BUILD, threads)
    try:
        wd, cam, st = scenes.medcity(gpu, width=96, height=54, spp=8, variance_threshold=0.0)
    finally:
        rt_mod._commit_override = prev
    wh = scenes.medcity(gpu, width=96, height=54, spp=8, variance_threshold=0.0)[0]
    assert wd.scene.build_info().built_on_device == 1 and wd.scene.primitive_breakdown().instances == 40000
    rays = W.ray_batch(1 << 16, W.scene_extent(wh.scene.bounds()), seed=8)
    hd, _ = wd.scene.intersect_n(rays)
    hh, _ = wh.scene.intersect_n(rays)
    same, tie = agree(hd, hh)
    assert (same | tie).all() and same.mean() >= 0.99  # the city repeats buildings at identical places: exact-t ties between instances
    a, sa = gpu.Renderer(cam, wh, st).render_with_stats()
    b, sb_ = gpu.Renderer(cam, wd, st).render_with_stats()
    rel = np.abs(a[..., :3] - b[..., :3]) / (np.abs(a[..., :3]) + 1e-3)
    assert (rel.max(axis=-1) > 1e-3).mean() < 0.01 and sa.camera_rays == sb_.camera_rays


def test_render_on_a_device_built_scene(gpu, orc):
    """The wavefront integrator over a GPU-built tree renders the same image (ties aside) as over the host tree."""
    from crust_render_b200.render_types import Camera, LightList, OpenPBR, RenderSettings
    meshes = W.stress_meshes(grid=2, shards=1000, jitter=False)

    def world(flags):
        sb = gpu.SceneBuilder()
        for v, i in meshes:
            sb.attach(Geometry.TriangleMesh(v, i))
        scene = sb.commit(flags=flags)
        mats = np.zeros(1, dtype=OpenPBR().dtype)
        mats[0] = OpenPBR()
        return gpu.World(scene, mats, np.zeros(len(meshes), np.uint32), LightList())
    cam = Camera((0, 9, 20), (0, 0, 0), (0, 1, 0), 40.0, 16 / 9)
    st = RenderSettings(4, 4, 160, 90, variance_threshold=0.0)
    a, sa = gpu.Renderer(cam, world(0), st).render_with_stats()
    b, sb_ = gpu.Renderer(cam, world(gpu.COMMIT_DEVICE_BUILD), st).render_with_stats()
    rel = np.abs(a[..., :3] - b[..., :3]) / (np.abs(a[..., :3]) + 1e-3)
    assert (rel.max(axis=-1) > 1e-3).mean() < 0.002 and sa.camera_rays == sb_.camera_rays


def test_sah_built_top_keeps_every_hit():
    """CRT_DEV_SAH_TOP (read once per process, hence the subprocess): the binary nodes above clusters of 64 references are
    rebuilt by a host SAH and spliced into the radix tree; the probe scene's hits must still be the oracle's."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import numpy as np
import crust_render_b200 as gpu
import orc
from crust_render_b200 import workloads as W
from test_gpu_device_build import commit_both, agree
gpu.init(0)
meshes = [W.uv_sphere(c, 0.45, 24, 12) for c in [(x * 1.2, y * 1.2, z * 1.2) for x in range(-2, 3) for y in range(-1, 2) for z in range(-2, 3)]]
plain, so = commit_both(gpu, orc, meshes)
rays = W.ray_batch(1 << 17, W.scene_extent(so.bounds()))
hg, _ = plain.intersect_n(rays)
ho, _ = so.intersect_n(rays, threads=0)
same, tie = agree(hg, ho)
assert (same | tie).all() and same.mean() >= 0.9999, same.mean()
assert np.array_equal(plain.occluded_n(rays), so.occluded_n(rays, threads=0))
info = plain.build_info()
assert info.built_on_device == 1
print("ok", int(info.wide_nodes), float(same.mean()))
''' % (root, root, root)
    env = dict(os.environ, CRT_DEV_SAH_TOP="64", CRT_DEVICE_BUILD_TRACE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ok" in r.stdout and "SAH top" in r.stderr, r.stderr[-2000:]
