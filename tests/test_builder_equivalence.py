"""Host logic, no GPU: the product's own commit() (crust-render_b200/csrc/host_build.cpp,
built independently of the oracle) must produce the reference's tree — node for
node, leaf for leaf — as pinned by the oracle restatement of bvh.rs:300-327,
1059-1375. Commits use CRT_COMMIT_HOST_ONLY (no upload; queries would fail)."""
import numpy as np
import pytest

from crust_render_b200 import workloads
from crust_render_b200.rt_types import Affine3A, CubicCurveSegment, CurveSegment, Geometry


def _commit_pair(crt, orc, make, threads=0):
    flags = crt.COMMIT_KEEP_HOST_TREE | crt.COMMIT_HOST_ONLY
    return make(crt).commit(flags=flags, host_threads=threads), make(orc).commit(host_threads=threads)


def _assert_same_tree(p, o):
    a, b = p.debug_tree(), o.tree()
    assert a["binary_nodes"] == b["binary_nodes"]
    assert a["nodes"].shape == b["nodes"].shape
    # bounds compare as floats (+0 == -0), topology words bit for bit
    assert np.array_equal(a["nodes"][:, :24].view(np.float32), b["nodes"][:, :24].view(np.float32))
    assert np.array_equal(a["nodes"][:, 24:29], b["nodes"][:, 24:29])
    for k in ("tri_counts", "other_counts", "refs"):
        assert np.array_equal(a[k], b[k]), k


def _mesh(v, i):
    def make(api):
        b = api.SceneBuilder()
        b.attach(Geometry.TriangleMesh(v, i))
        return b
    return make


@pytest.mark.parametrize("n", [1, 3, 5, 64, 256, 2000])
def test_diagonal_shards(crt, orc, n):  # SBVH-heavy (bvh.rs:1443-1458)
    _assert_same_tree(*_commit_pair(crt, orc, _mesh(*workloads.diagonal_shards(n))))


@pytest.mark.parametrize("n", [1, 2, 4, 6])
def test_sphere_grid(crt, orc, n):
    _assert_same_tree(*_commit_pair(crt, orc, lambda api: workloads.sphere_grid_prims(api, n)))


def test_tri_spheres_43200(crt, orc):  # ray_throughput.rs:50-66
    _assert_same_tree(*_commit_pair(crt, orc, workloads.triangle_scene))


def test_parallel_build_is_thread_count_invariant(crt, orc):  # bvh.rs:1156-1168, 1620-1638
    v, i = workloads.uv_sphere((0, 0, 0), 1.0, 200, 100)  # 40 000 triangles: forks above 4096 refs
    p1, o1 = _commit_pair(crt, orc, _mesh(v, i), threads=1)
    p8, _ = _commit_pair(crt, orc, _mesh(v, i), threads=8)
    _assert_same_tree(p1, o1)
    _assert_same_tree(p8, o1)


@pytest.mark.parametrize("threads", [2, 3, 8])
def test_data_parallel_node_passes_do_not_change_the_tree(crt, orc, threads, monkeypatch):
    """Large nodes run their bounds / binning / spatial-split / partition passes over chunks of the references in parallel
    (host_build.cpp `Par`, normally from 2^19 references up); CRT_BUILD_PAR_MIN=64 forces that path on small scenes. The tree
    must still be the oracle's, whatever the chunking (odd thread counts give ragged chunks)."""
    monkeypatch.setenv("CRT_BUILD_PAR_MIN", "64")
    for make in (_mesh(*workloads.diagonal_shards(2000)), lambda api: workloads.stress_scene(api, grid=2, shards=3000)[0],
                 lambda api: workloads.sphere_grid_prims(api, 6)):
        p, o = _commit_pair(crt, orc, make, threads=threads)
        _assert_same_tree(p, o)


def test_stress_scene_small(crt, orc):  # scripts/gen_stress_scene.py at reduced size: spheres + shards + quads
    def make(api):
        return workloads.stress_scene(api, grid=2, shards=3000)[0]
    _assert_same_tree(*_commit_pair(crt, orc, make))


def test_mixed_kinds_and_degenerates(crt, orc):
    rng = np.random.default_rng(7)

    def make(api):
        b = api.SceneBuilder()
        v, i = workloads.diagonal_shards(40)
        i = i.copy()
        i[5] = (0, 1, 9999)  # out-of-range index: skipped, prim_id still consumed (scene.rs:251-253)
        b.attach(Geometry.TriangleMesh(v, i))
        b.attach(api.SceneBuilder.empty_geometry())
        for k in range(30):
            b.attach(Geometry.Sphere(rng.uniform(-5, 15, 3).astype(np.float32), float(rng.uniform(0.1, 1.0))))
        segs = [CurveSegment(rng.uniform(0, 10, 3), rng.uniform(0, 10, 3), 0.1, 0.05) for _ in range(25)]
        b.attach(Geometry.RoundCurves(segs))
        cub = [CubicCurveSegment(rng.uniform(0, 10, (4, 3)), 0.1, 0.2) for _ in range(17)]
        b.attach(Geometry.CubicCurves(cub))
        # coincident centroids: identical triangles -> median split path (bvh.rs:1148-1152)
        same = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (9, 1))
        b.attach(Geometry.TriangleMesh(same, np.arange(27, dtype=np.uint32).reshape(9, 3)))
        # zero-area triangle: flat-axis padding (aabb.rs:47-58)
        b.attach(Geometry.TriangleMesh([(3, 3, 3), (3, 3, 3), (3, 3, 3)], [(0, 1, 2)]))
        return b

    rng = np.random.default_rng(7)
    p = make(crt).commit(flags=crt.COMMIT_KEEP_HOST_TREE | crt.COMMIT_HOST_ONLY)
    rng = np.random.default_rng(7)
    o = make(orc).commit()
    _assert_same_tree(p, o)
    assert p.primitive_count() == o.primitive_count() == 39 + 30 + 25 + 17 + 9 + 1
    assert p.geometry_count() == o.geometry_count()


def test_instances_bounds_and_motion(crt, orc):
    def make(api, commit):
        inner = api.SceneBuilder()
        inner.attach(Geometry.TriangleMesh(*workloads.uv_sphere((0, 0, 0), 1.0, 24, 12)))
        inner = commit(inner)
        empty = commit(api.SceneBuilder())
        b = api.SceneBuilder()
        rot = Affine3A.from_rotation_y(0.7) * Affine3A.from_scale((1.5, 0.5, 2.0))
        for k in range(40):
            t = Affine3A.from_translation((k % 5 * 2.5, k // 5 * 2.5, (k % 3) * 1.7)) * rot
            end = Affine3A.from_translation((0.5, 0.25, 0)) * t if k % 7 == 0 else None
            b.attach(Geometry.Instance(inner, t, end))
        b.attach(Geometry.Instance(empty, Affine3A.IDENTITY))  # empty instanced scene: skipped (scene.rs:307-309)
        return b

    flags = crt.COMMIT_KEEP_HOST_TREE | crt.COMMIT_HOST_ONLY
    p = make(crt, lambda b: b.commit(flags=flags)).commit(flags=flags)
    o = make(orc, lambda b: b.commit()).commit()
    _assert_same_tree(p, o)
    assert p.has_motion() and o.has_motion()
    assert p.primitive_count() == o.primitive_count() == 40
    assert np.array_equal(p.bounds().minimum, o.bounds().minimum) and np.array_equal(p.bounds().maximum, o.bounds().maximum)
    pb, ob = p.unique_primitive_breakdown(), o.unique_primitive_breakdown()
    assert (pb.triangles, pb.instances) == (ob.triangles, ob.instances) == (576, 40)


def test_host_only_scene_refuses_queries(crt):
    """No CPU fallback: a scene without device data cannot be queried."""
    s = workloads.sphere_grid_prims(crt, 2).commit(flags=crt.COMMIT_HOST_ONLY)
    with pytest.raises(crt.CrtError) as e:
        s.intersect(crt.Ray((0, 0, -5), (0, 0, 1)), 0.001, 100.0)
    assert e.value.status == 4


def test_set_geometry_unknown_id_is_an_error(crt):  # scene.rs:196-200 (reference panics)
    b = crt.SceneBuilder()
    with pytest.raises(crt.CrtError) as e:
        b.set_geometry(3, Geometry.Sphere((0, 0, 0), 1.0))
    assert e.value.status == 2
    assert b.count() == 0


def test_builder_is_consumed_by_commit(crt):  # scene.rs:226 commit(self)
    b = crt.SceneBuilder()
    b.attach(Geometry.Sphere((0, 0, 0), 1.0))
    b.commit(flags=crt.COMMIT_HOST_ONLY)
    with pytest.raises(crt.CrtError):
        b.attach(Geometry.Sphere((0, 0, 0), 1.0))


@pytest.mark.parametrize("threads,depth", [(2, 1), (3, 2), (8, 3)])
def test_parallel_collapse_writes_the_sequential_arrays(crt, orc, threads, depth, monkeypatch):
    """The BVH4 collapse measures subtrees in a dry run and then writes them in parallel at their offsets (host_build.cpp
    Flattener; normally from 2^16 binary nodes up, subtrees five levels below the root). Forced onto small scenes of every
    record kind — triangles with and without normals, spheres, both curve kinds, masked / moving instances — the wide nodes,
    the leaf record stream, the motion pairs and the instance record offsets must equal the one-thread commit's word for
    word, and the tree the oracle's."""
    def mixed(api):
        b = api.SceneBuilder()
        v, i = workloads.uv_sphere((0, 0, 0), 1.0, 48, 24)
        b.attach(Geometry.TriangleMesh(v, i))
        v2, i2 = workloads.uv_sphere((2.5, 0.2, 0), 0.7, 24, 12)
        b.attach(Geometry.TriangleMesh(v2, i2, normals=(v2 - np.array([2.5, 0.2, 0], np.float32)) / np.float32(0.7)))
        for k in range(40):
            b.attach(Geometry.Sphere(np.array([k * 0.3 - 6, 2.0, (k % 5) * 0.4], np.float32), 0.1 + 0.01 * k))
        segs = [CurveSegment(np.array([k * 0.2, -2, 0], np.float32), np.array([k * 0.2 + 0.1, -1.5, 0.3], np.float32), 0.05, 0.02) for k in range(30)]
        b.attach(Geometry.RoundCurves(segs))
        cub = [CubicCurveSegment(np.array([[k, 3, 0], [k + 0.3, 3.5, 0.2], [k + 0.6, 3.2, -0.2], [k + 0.9, 3.6, 0]], np.float32), 0.04, 0.01) for k in range(12)]
        b.attach(Geometry.CubicCurves(cub))
        return b

    def instanced(api):
        inner = api.SceneBuilder()
        v, i = workloads.uv_sphere((0, 0, 0), 1.0, 12, 6)
        inner.attach(Geometry.TriangleMesh(v, i))
        inner = inner.commit(flags=crt.COMMIT_KEEP_HOST_TREE | crt.COMMIT_HOST_ONLY) if api is crt else inner.commit()
        b = api.SceneBuilder()
        for k in range(300):
            t = Affine3A.from_translation(np.array([k % 10, (k // 10) % 10, k // 100], np.float32) * np.float32(2.5))
            end = Affine3A.from_translation(np.array([k % 10, (k // 10) % 10, k // 100], np.float32) * np.float32(2.5) + np.float32(0.3)) if k % 3 == 0 else None
            g = Geometry.Instance(inner, t * Affine3A.from_rotation_y(0.1 * k), end)
            if k % 7 == 0:
                b.attach_masked(g, 0x5)
            else:
                b.attach(g)
        fv = np.array([[-3, -2, -3], [30, -2, -3], [30, -2, 30], [-3, -2, 30]], np.float32)
        b.attach(Geometry.TriangleMesh(fv, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)))
        return b

    for make in (mixed, instanced, _mesh(*workloads.diagonal_shards(3000))):
        seq, o = _commit_pair(crt, orc, make, threads=1)
        monkeypatch.setenv("CRT_FLATTEN_PAR_MIN", "1")
        monkeypatch.setenv("CRT_FLATTEN_TASK_DEPTH", str(depth))
        par, _ = _commit_pair(crt, orc, make, threads=threads)
        monkeypatch.delenv("CRT_FLATTEN_PAR_MIN")
        monkeypatch.delenv("CRT_FLATTEN_TASK_DEPTH")
        _assert_same_tree(seq, o)
        _assert_same_tree(par, o)
        a, b = seq.debug_host_arrays(), par.debug_host_arrays()
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert len(a["leafs"]) > 0
