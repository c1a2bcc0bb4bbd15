"""ORACLE BINDING — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper over oracle/liborc.so (the C++ restatement of crust-rt) exposing
the same duck-typed surface as ``crust_render_b200`` (SceneBuilder / Scene /
Geometry / Ray) so a test can run the reference's scenario against both and
compare. Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU arm import
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from crust_render_b200.rt_types import (  # value types only — no product code paths
    HIT_DTYPE,
    MASK_ALL,
    RAY_DTYPE,
    Geometry,
    PrimitiveBreakdown,
    Ray,
    RayHit,
    make_rays,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

STAT_NAMES = ["queries_top", "queries_inst", "nodes_top", "nodes_inst", "leaves_top", "leaves_inst", "packets_top",
              "packets_inst", "prim_tests_top", "prim_tests_inst", "tri_lanes", "sphere_tests", "curve_tests",
              "cubic_tests", "instance_tests", "stack_high_water"]


def build() -> str:
    """Compiles liborc.so with the committed Makefile (idempotent)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return os.path.join(_HERE, "liborc.so")


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, u32, u64, sz, f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float
    L.orc_builder_new.restype = vp
    L.orc_builder_free.argtypes = [vp]
    L.orc_builder_count.argtypes = [vp]
    L.orc_builder_count.restype = u64
    L.orc_attach_triangle_mesh.argtypes = [vp, vp, sz, vp, sz, vp, sz, u32]
    L.orc_attach_sphere.argtypes = [vp, f, f, f, f, u32]
    L.orc_attach_round_curves.argtypes = [vp, vp, sz, u32]
    L.orc_attach_cubic_curves.argtypes = [vp, vp, sz, u32]
    L.orc_attach_instance.argtypes = [vp, vp, vp, vp, u32]
    L.orc_attach_empty.argtypes = [vp, u32]
    for n in ("orc_attach_triangle_mesh", "orc_attach_sphere", "orc_attach_round_curves", "orc_attach_cubic_curves",
              "orc_attach_instance", "orc_attach_empty"):
        getattr(L, n).restype = u32
    L.orc_set_triangle_mesh.argtypes = [vp, u32, vp, sz, vp, sz, vp, sz]
    L.orc_set_sphere.argtypes = [vp, u32, f, f, f, f]
    L.orc_set_instance.argtypes = [vp, u32, vp, vp, vp]
    L.orc_commit.argtypes = [vp, C.c_int]
    L.orc_commit.restype = vp
    L.orc_scene_release.argtypes = [vp]
    L.orc_geometry_count.argtypes = [vp]
    L.orc_geometry_count.restype = u32
    L.orc_primitive_count.argtypes = [vp]
    L.orc_primitive_count.restype = u64
    L.orc_has_motion.argtypes = [vp]
    L.orc_bounds.argtypes = [vp, vp]
    L.orc_primitive_breakdown.argtypes = [vp, vp]
    L.orc_unique_primitive_breakdown.argtypes = [vp, vp]
    L.orc_tree_sizes.argtypes = [vp, vp]
    L.orc_tree_nodes.argtypes = [vp, vp]
    L.orc_tree_leaves.argtypes = [vp, vp, vp, vp]
    L.orc_packet_lanes.argtypes = [vp]
    L.orc_packet_lanes.restype = u64
    L.orc_intersect_n.argtypes = [vp, u64, vp, vp, vp, vp, vp, C.c_int, vp]
    L.orc_intersect_n.restype = C.c_double
    L.orc_occluded_n.argtypes = [vp, u64, vp, vp, vp, vp, C.c_int, vp]
    L.orc_occluded_n.restype = C.c_double
    L.orc_linear_scan.argtypes = [vp, vp, f, u32, vp]
    L.orc_test_triangle_intersect.argtypes = [vp, vp, vp, f, f, vp]
    L.orc_test_tri4_intersect.argtypes = [vp, vp, vp, vp, C.c_int, u32, f, f, vp, vp]
    L.orc_test_rounded_cone.argtypes = [vp, vp, vp, vp, f, f, f, f, vp]
    L.orc_test_cubic_curve.argtypes = [vp, vp, vp, f, f, f, f, vp]
    L.orc_test_flatness_depth.argtypes = [vp, f]
    L.orc_test_flatness_depth.restype = u32
    L.orc_test_clip_triangle_aabb.argtypes = [vp, C.c_int, f, f, vp]
    L.orc_sizeof_wide_node.restype = u32
    # integrator + shading hooks (crt_oracle_render.hpp)
    L.orc_world_new.argtypes = [vp, vp, sz, vp, sz, vp, sz]
    L.orc_world_new.restype = vp
    L.orc_world_free.argtypes = [vp]
    L.orc_world_set_environment.argtypes = [vp, u32, u32, u32, vp, vp]
    L.orc_world_set_volumes.argtypes = [vp, vp, sz]
    L.orc_volumes_new.argtypes = [vp, sz]
    L.orc_volumes_new.restype = vp
    L.orc_volumes_free.argtypes = [vp]
    L.orc_rng_new.argtypes = [u32]
    L.orc_rng_new.restype = u32
    L.orc_rng_next.argtypes = [vp]
    L.orc_rng_next.restype = f
    L.orc_volumes_region_info.argtypes = [vp, u32, vp]
    L.orc_volumes_transmittance_n.argtypes = [vp, vp, vp, f, f, vp, sz, vp]
    L.orc_volumes_sample_n.argtypes = [vp, vp, vp, f, f, vp, sz, vp]
    L.orc_volumes_field_density.argtypes = [vp, u32, vp]
    L.orc_volumes_field_density.restype = f
    L.orc_volumes_density.argtypes = [vp, u32, vp]
    L.orc_volumes_density.restype = f
    L.orc_volumes_intersect.argtypes = [vp, u32, vp, vp, vp]
    L.orc_vol_hash3.argtypes = [C.c_int32, C.c_int32, C.c_int32, u32]
    L.orc_vol_hash3.restype = f
    L.orc_hg_phase.argtypes = [f, f]
    L.orc_hg_phase.restype = f
    L.orc_phase_single_pdf.argtypes = [f, f]
    L.orc_phase_single_pdf.restype = f
    L.orc_mat4_inverse.argtypes = [vp, vp]
    L.orc_world_light_sample.argtypes = [vp, u32, vp, f, f, vp]
    L.orc_world_light_escaped.argtypes = [vp, u32, vp, vp, vp]
    L.orc_env_new.argtypes = [u32, u32, vp]
    L.orc_env_new.restype = vp
    L.orc_env_free.argtypes = [vp]
    L.orc_env_call.argtypes = [vp, C.c_int, vp, vp]
    L.orc_render.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
    L.orc_render.restype = C.c_double
    L.orc_material_default.argtypes = [vp]
    L.orc_material_pmf.argtypes = [vp, vp]
    L.orc_material_medium.argtypes = [vp, vp, vp, vp]
    L.orc_material_eval.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.orc_material_scatter.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    L.orc_material_emitted_directional.argtypes = [vp, f, vp]
    L.orc_light_sample_li.argtypes = [vp, vp, f, f, vp]
    L.orc_light_pdf_at_point.argtypes = [vp, vp, vp]
    L.orc_light_pdf_at_point.restype = f
    L.orc_light_escaped.argtypes = [vp, vp, vp, vp]
    L.orc_light_pick.argtypes = [C.c_int, f]
    L.orc_camera_ray.argtypes = [vp, f, f, f, f, vp, vp]
    L.orc_filter_sample.argtypes = [u32, f, f, vp, vp]
    L.orc_filter_eval.argtypes = [u32, f, f]
    L.orc_filter_eval.restype = f
    L.orc_sampler_draw.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int, vp, vp]
    L.orc_mis_weight.argtypes = [u32, C.c_int, f, f]
    L.orc_mis_weight.restype = f
    L.orc_shade_fn.argtypes = [C.c_char_p, vp, vp, vp]
    L.orc_filter_sample_n.argtypes = [u32, f, sz, vp, vp, vp]
    _lib = L
    return L


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))


class SceneBuilder:
    def __init__(self):
        self._h = C.c_void_p(lib().orc_builder_new())
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_builder_free(self._h)
            self._h = None

    def attach(self, g: Geometry) -> int:
        return self.attach_masked(g, MASK_ALL)

    def attach_masked(self, g: Geometry, mask: int) -> int:
        L, h = lib(), self._h
        if g.kind == "mesh":
            n = g.normals
            return L.orc_attach_triangle_mesh(h, _p(g.vertices), len(g.vertices), _p(g.indices), len(g.indices), _p(n),
                                              0 if n is None else len(n), mask)
        if g.kind == "sphere":
            c = g.center
            return L.orc_attach_sphere(h, float(c[0]), float(c[1]), float(c[2]), g.radius, mask)
        if g.kind == "curves":
            return L.orc_attach_round_curves(h, _p(g.segments), len(g.segments), mask)
        if g.kind == "cubics":
            return L.orc_attach_cubic_curves(h, _p(g.segments), len(g.segments), mask)
        if g.kind == "instance":
            self._keep.append(g.scene)
            return L.orc_attach_instance(h, g.scene._h, _p(g.transform.m), None if g.transform_end is None else _p(g.transform_end.m), mask)
        if g.kind == "empty":
            return L.orc_attach_empty(h, mask)
        raise ValueError(g.kind)

    def count(self) -> int:
        return int(lib().orc_builder_count(self._h))

    def reserve(self, additional: int) -> None:
        pass

    def set_geometry(self, id: int, g: Geometry) -> None:
        L, h = lib(), self._h
        if g.kind == "mesh":
            n = g.normals
            rc = L.orc_set_triangle_mesh(h, id, _p(g.vertices), len(g.vertices), _p(g.indices), len(g.indices), _p(n), 0 if n is None else len(n))
        elif g.kind == "sphere":
            c = g.center
            rc = L.orc_set_sphere(h, id, float(c[0]), float(c[1]), float(c[2]), g.radius)
        elif g.kind == "instance":
            self._keep.append(g.scene)
            rc = L.orc_set_instance(h, id, g.scene._h, _p(g.transform.m), None if g.transform_end is None else _p(g.transform_end.m))
        else:
            raise ValueError(g.kind)
        if rc:
            raise IndexError("set_geometry: id was never attached")

    @staticmethod
    def empty_geometry() -> Geometry:
        return Geometry.Empty()

    def commit(self, flags: int = 0, host_threads: int = 0) -> "Scene":
        h, self._h = self._h, None
        s = Scene(C.c_void_p(lib().orc_commit(h, host_threads)))
        s._keep = self._keep
        return s


class Scene:
    def __init__(self, h):
        self._h = h
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scene_release(self._h)
            self._h = None

    def intersect_n(self, rays, time=None, mask=None, threads: int = 1, stats: bool = False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        time = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        hits = np.empty(n, dtype=HIT_DTYPE)
        ff = np.empty(n, dtype=np.uint8)
        st = np.zeros(16, dtype=np.uint64) if stats else None
        secs = lib().orc_intersect_n(self._h, n, _p(rays), _p(time), _p(mask), _p(hits), _p(ff), threads, _p(st))
        self.last_seconds = secs
        if stats:
            return hits, ff, dict(zip(STAT_NAMES, (int(x) for x in st)))
        return hits, ff

    def occluded_n(self, rays, time=None, mask=None, threads: int = 1, stats: bool = False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        time = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        occ = np.empty(n, dtype=np.uint8)
        st = np.zeros(16, dtype=np.uint64) if stats else None
        secs = lib().orc_occluded_n(self._h, n, _p(rays), _p(time), _p(mask), _p(occ), threads, _p(st))
        self.last_seconds = secs
        if stats:
            return occ, dict(zip(STAT_NAMES, (int(x) for x in st)))
        return occ

    def intersect(self, ray: Ray, t_min: float, t_max: float) -> Optional[RayHit]:
        r = make_rays([ray.origin], [ray.dir], t_min, t_max)
        hits, ff = self.intersect_n(r, np.array([ray.time], np.float32), np.array([ray.mask & 0xFFFFFFFF], np.uint32))
        h = hits[0]
        if h["geom_id"] == 0xFFFFFFFF:
            return None
        return RayHit(float(h["t"]), h["normal"].copy(), bool(ff[0]), float(h["u"]), float(h["v"]), int(h["geom_id"]), int(h["prim_id"]))

    def occluded(self, ray: Ray, t_min: float, t_max: float) -> bool:
        r = make_rays([ray.origin], [ray.dir], t_min, t_max)
        return bool(self.occluded_n(r, np.array([ray.time], np.float32), np.array([ray.mask & 0xFFFFFFFF], np.uint32))[0])

    def linear_scan(self, ray: Ray, t_min: float, t_max: float):
        """bvh.rs:1460-1470: closest hit by testing every primitive in input order (outward normal)."""
        r = make_rays([ray.origin], [ray.dir], t_min, t_max)
        hit = np.zeros(1, dtype=HIT_DTYPE)
        got = lib().orc_linear_scan(self._h, _p(r), np.float32(ray.time), ray.mask & 0xFFFFFFFF, _p(hit))
        return hit[0] if got else None

    def bounds(self):
        o = np.zeros(6, np.float32)
        if not lib().orc_bounds(self._h, _p(o)):
            return None
        from crust_render_b200.rt_types import _v3  # noqa

        class _B:
            minimum = o[:3].copy()
            maximum = o[3:].copy()

        return _B

    def geometry_count(self) -> int:
        return int(lib().orc_geometry_count(self._h))

    def primitive_count(self) -> int:
        return int(lib().orc_primitive_count(self._h))

    def has_motion(self) -> bool:
        return bool(lib().orc_has_motion(self._h))

    def primitive_breakdown(self) -> PrimitiveBreakdown:
        o = np.zeros(5, np.uint64)
        lib().orc_primitive_breakdown(self._h, _p(o))
        return PrimitiveBreakdown(*(int(x) for x in o))

    def unique_primitive_breakdown(self) -> PrimitiveBreakdown:
        o = np.zeros(5, np.uint64)
        lib().orc_unique_primitive_breakdown(self._h, _p(o))
        return PrimitiveBreakdown(*(int(x) for x in o))

    def tree(self):
        """Canonical tree export, same shape as crust_render_b200.Scene.debug_tree()."""
        sz = np.zeros(6, np.uint64)
        lib().orc_tree_sizes(self._h, _p(sz))
        n_wide, n_leaves, n_pk, n_idx, n_refs, n_bin = (int(x) for x in sz)
        nodes = np.zeros((n_wide, 32), np.uint32)
        lib().orc_tree_nodes(self._h, _p(nodes))
        tri = np.zeros(n_leaves, np.uint32)
        oth = np.zeros(n_leaves, np.uint32)
        refs = np.zeros(n_refs, np.uint32)
        lib().orc_tree_leaves(self._h, _p(tri), _p(oth), _p(refs))
        return {"nodes": nodes, "tri_counts": tri, "other_counts": oth, "refs": refs, "binary_nodes": n_bin,
                "packets": n_pk, "indices": n_idx, "packet_lanes": int(lib().orc_packet_lanes(self._h))}


# ---- unit hooks -----------------------------------------------------------------
def triangle_intersect(o, d, v0, v1, v2, t_min, t_max):
    """triangle.rs:97-106 -> (t, u, v) or None."""
    out = np.zeros(3, np.float32)
    v9 = _f3([v0, v1, v2])
    ok = lib().orc_test_triangle_intersect(_p(_f3(o)), _p(_f3(d)), _p(v9), t_min, t_max, _p(out))
    return tuple(out) if ok else None


def tri4_intersect(o, d, tris, masks, ray_mask, t_min, t_max):
    """triangle.rs:217-348: tris = list of (v0, v1, v2). Returns dict hits/fallback/active/t/u/v."""
    n = len(tris)
    t9 = _f3([[v for v in tri] for tri in tris])
    m = np.asarray(masks, np.uint32)
    bits = np.zeros(3, np.uint32)
    tuv = np.zeros(12, np.float32)
    lib().orc_test_tri4_intersect(_p(_f3(o)), _p(_f3(d)), _p(t9), _p(m), n, ray_mask & 0xFFFFFFFF, t_min, t_max, _p(bits), _p(tuv))
    return {"hits": int(bits[0]), "fallback": int(bits[1]), "active": int(bits[2]), "t": tuv[0:4], "u": tuv[4:8], "v": tuv[8:12]}


def rounded_cone_intersect(o, d, p0, p1, r0, r1, t_min, t_max):
    out = np.zeros(4, np.float32)
    ok = lib().orc_test_rounded_cone(_p(_f3(o)), _p(_f3(d)), _p(_f3(p0)), _p(_f3(p1)), r0, r1, t_min, t_max, _p(out))
    return (float(out[0]), out[1:4].copy()) if ok else None


def cubic_curve_intersect(o, d, cp, r0, r1, t_min, t_max):
    out = np.zeros(4, np.float32)
    ok = lib().orc_test_cubic_curve(_p(_f3(o)), _p(_f3(d)), _p(_f3(cp)), r0, r1, t_min, t_max, _p(out))
    return (float(out[0]), out[1:4].copy()) if ok else None


def flatness_depth(cp, max_width) -> int:
    return int(lib().orc_test_flatness_depth(_p(_f3(cp)), max_width))


def clip_triangle_aabb(v0, v1, v2, axis, mn, mx):
    out = np.zeros(6, np.float32)
    ok = lib().orc_test_clip_triangle_aabb(_p(_f3([v0, v1, v2])), axis, mn, mx, _p(out))
    return (out[:3].copy(), out[3:].copy()) if ok else None


# ---- engine side: World / Renderer over the oracle (crt_oracle_render.hpp) -----------------------
from crust_render_b200.render_types import (  # noqa: E402  value types only
    LIGHT_DTYPE,
    MATERIAL_DTYPE,
    STATS_DTYPE,
    Camera,
    LightList,
    RayStats,
    RenderSettings,
    WorldBuilderBase,
    volume_records,
)


class World:
    def __init__(self, scene: "Scene", materials: np.ndarray, geom_material: np.ndarray, lights: LightList):
        self.scene = scene
        self.materials, self.geom_material = materials, geom_material  # the geom_id-indexed table (rt_world.rs:142-160)
        self._lights = lights.records()
        self._h = C.c_void_p(lib().orc_world_new(scene._h, _p(materials), len(materials), _p(geom_material), len(geom_material),
                                                 _p(self._lights), len(self._lights)))
        for index, (px, cols) in lights.environments.items():
            if lib().orc_world_set_environment(self._h, index, px.shape[1], px.shape[0], _p(px), _p(cols)):
                raise ValueError("malformed environment map")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_world_free(self._h)
            self._h = None

    def with_volumes(self, regions) -> "World":
        self._volumes = list(regions)
        recs = volume_records(self._volumes)
        if lib().orc_world_set_volumes(self._h, _p(recs) if len(recs) else None, len(recs)):
            raise ValueError("bad volume regions")
        return self

    def count(self) -> int:
        return self.scene.geometry_count()

    def has_motion(self) -> bool:
        return self.scene.has_motion()

    def light_sample_li(self, index: int, from_, u: float, v: float):
        """Light::sample_li of light `index`, environment map included."""
        out = np.zeros(8, np.float32)
        if not lib().orc_world_light_sample(self._h, index, _p(_f3(from_)), u, v, _p(out)):
            return None
        return dict(direction=out[0:3].copy(), radiance=out[3:6].copy(), distance=float(out[6]), pdf=float(out[7]))

    def light_escaped(self, index: int, direction):
        r, pdf = np.zeros(3, np.float32), C.c_float()
        if not lib().orc_world_light_escaped(self._h, index, _p(_f3(direction)), _p(r), C.byref(pdf)):
            return None
        return r, pdf.value


class WorldBuilder(WorldBuilderBase):
    def _new_scene_builder(self):
        return SceneBuilder()

    def commit(self, lights: Optional[LightList] = None) -> World:
        mats, gm = self.material_table()
        return World(self._sb.commit(), mats, gm, lights or LightList())


class Renderer:
    """`Renderer::new(camera, world, lights, settings)` (tracer.rs:149-163); the LightList travels in the World here."""

    def __init__(self, camera: Camera, world: World, settings: RenderSettings):
        self.camera, self.world, self.settings = camera, world, settings
        self.seconds = 0.0

    def render_with_stats(self, threads: int = 0, var_map: bool = False):
        st = self.settings
        fb = np.zeros((st.height, st.width, 4), dtype=np.float32)
        vm = np.zeros((st.height, st.width), dtype=np.float64) if var_map else None
        stats = np.zeros((), dtype=STATS_DTYPE)
        cam, rec = self.camera.record(), st.record()
        self.seconds = lib().orc_render(self.world._h, _p(cam), _p(rec), _p(fb), _p(vm), _p(stats), threads)
        self.var_map = vm
        return fb, RayStats.from_record(stats)

    def render(self, threads: int = 0) -> np.ndarray:
        return self.render_with_stats(threads)[0]


# ---- leaf shading math hooks (tests/test_shade_kats.py, test_sampler.py) ---------------------------
def shade_fn(name: str, args, material: Optional[np.ndarray] = None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(args, dtype=np.float32).reshape(-1))
    out = np.zeros(16, dtype=np.float32)
    n = lib().orc_shade_fn(name.encode(), _p(material), _p(a), _p(out))
    if n < 0:
        raise KeyError(name)
    return out[:n].copy()


def material_eval(material, ray_dir, normal, front_face, wi, p=(0, 0, 0)):
    """Material::eval (material.rs:64-69): (value, pdf) or None."""
    v, pdf = np.zeros(3, np.float32), C.c_float()
    ok = lib().orc_material_eval(_p(material), _p(_f3(ray_dir)), _p(_f3(p)), _p(_f3(normal)), int(front_face), _p(_f3(wi)), _p(v), C.byref(pdf))
    return (v, pdf.value) if ok else None


def material_scatter(material, ray_dir, normal, front_face, s4, p=(0, 0, 0)):
    """Material::scatter_importance with an explicit 4-D block: dict(origin, dir, value, pdf, delta, enters_medium) or None."""
    out = np.zeros(11, np.float32)
    s = np.ascontiguousarray(np.asarray(s4, np.float32))
    if not lib().orc_material_scatter(_p(material), _p(_f3(ray_dir)), _p(_f3(p)), _p(_f3(normal)), int(front_face), _p(s), _p(out)):
        return None
    fl = int(out[10])
    return dict(origin=out[0:3].copy(), dir=out[3:6].copy(), value=out[6:9].copy(), pdf=float(out[9]), delta=bool(fl & 1), enters_medium=bool(fl & 2))


def material_medium(material):
    sa, ss, g = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
    has = lib().orc_material_medium(_p(material), _p(sa), _p(ss), C.byref(g))
    return (sa, ss, g.value) if has else None


def material_pmf(material) -> np.ndarray:
    out = np.zeros(5, np.float32)
    lib().orc_material_pmf(_p(material), _p(out))
    return out


def material_emitted_directional(material, cos_o: float) -> np.ndarray:
    out = np.zeros(3, np.float32)
    lib().orc_material_emitted_directional(_p(material), cos_o, _p(out))
    return out


def light_sample_li(light, from_, u, v):
    out = np.zeros(8, np.float32)
    if not lib().orc_light_sample_li(_p(light), _p(_f3(from_)), u, v, _p(out)):
        return None
    return dict(direction=out[0:3].copy(), radiance=out[3:6].copy(), distance=float(out[6]), pdf=float(out[7]))


def light_pdf_at_point(light, from_, point) -> float:
    return lib().orc_light_pdf_at_point(_p(light), _p(_f3(from_)), _p(_f3(point)))


def light_escaped(light, direction):
    r, pdf = np.zeros(3, np.float32), C.c_float()
    if not lib().orc_light_escaped(_p(light), _p(_f3(direction)), _p(r), C.byref(pdf)):
        return None
    return r, pdf.value


def light_pick(n: int, u: float) -> int:
    return lib().orc_light_pick(n, u)


def camera_ray(camera: Camera, s, t, lu=0.0, lv=0.0):
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    rec = camera.record()
    lib().orc_camera_ray(_p(rec), s, t, lu, lv, _p(o), _p(d))
    return o, d


def filter_sample(kind: int, radius: float, u: float):
    off, w = C.c_float(), C.c_float()
    lib().orc_filter_sample(kind, radius, u, C.byref(off), C.byref(w))
    return off.value, w.value


def filter_eval(kind: int, radius: float, x: float) -> float:
    return lib().orc_filter_eval(kind, radius, x)


def sampler_draw(x: int, y: int, frame: int, index: int, keys=()):
    """PathSampler::new(x, y, frame, index).new_domain(k)... -> (4-D block, rnd)."""
    k = np.ascontiguousarray(np.asarray(list(keys), dtype=np.int32))
    out, rnd = np.zeros(4, np.float32), C.c_float()
    lib().orc_sampler_draw(x, y, frame, index, _p(k) if len(k) else None, len(k), _p(out), C.byref(rnd))
    return out, rnd.value


def mis_weight(strategy: int, light_side: bool, a: float, b: float) -> float:
    return lib().orc_mis_weight(int(strategy), int(light_side), a, b)


def filter_sample_n(kind: int, radius: float, u: np.ndarray):
    u = np.ascontiguousarray(u, dtype=np.float32)
    off, w = np.zeros_like(u), np.zeros_like(u)
    lib().orc_filter_sample_n(kind, radius, len(u), _p(u), _p(off), _p(w))
    return off, w


class EnvironmentMap:
    """environment.rs EnvironmentMap over the oracle's tables (tests only)."""

    def __init__(self, width: int, height: int, pixels):
        px = np.ascontiguousarray(np.asarray(pixels, np.float32).reshape(-1))
        self.valid = width > 0 and height > 0 and len(px) == 3 * width * height  # EnvironmentMap::new -> None otherwise
        self._h = C.c_void_p(lib().orc_env_new(int(width), int(height), _p(px))) if self.valid else None

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_env_free(self._h)
            self._h = None

    def _call(self, op, args, n_out):
        a = np.ascontiguousarray(np.asarray(args, np.float32))
        out = np.zeros(8, np.float32)
        rc = lib().orc_env_call(self._h, op, _p(a), _p(out))
        return rc, out[:n_out].copy()

    def sample(self, u1, u2):
        rc, o = self._call(0, [u1, u2], 7)
        return None if rc <= 0 else (o[0:3], o[3:6], float(o[6]))

    def pdf(self, d) -> float:
        return float(self._call(1, d, 1)[1][0])

    def radiance(self, d):
        return self._call(2, d, 3)[1]

    @staticmethod
    def direction_to_uv(d):
        a = np.ascontiguousarray(np.asarray(d, np.float32))
        out = np.zeros(8, np.float32)
        lib().orc_env_call(None, 3, _p(a), _p(out))
        return float(out[0]), float(out[1])

    @staticmethod
    def uv_to_direction(u, v):
        a = np.asarray([u, v], np.float32)
        out = np.zeros(8, np.float32)
        lib().orc_env_call(None, 4, _p(a), _p(out))
        return out[:3].copy()


class Rng:
    """The walk stream (`openqmc::pcg::Rng` in the reference; shade/volume.h's PCG here)."""

    def __init__(self, seed: int):
        self.state = C.c_uint32(lib().orc_rng_new(seed & 0xFFFFFFFF))

    def next_f32(self) -> float:
        return float(lib().orc_rng_next(C.byref(self.state)))


class Volumes:
    """`Volumes::new(regions)` (volume.rs:364-384) over shade/volume.h, for the unit tests of volume.rs."""

    def __init__(self, regions):
        self._regions = list(regions)
        recs = volume_records(self._regions)
        self._h = C.c_void_p(lib().orc_volumes_new(_p(recs) if len(recs) else None, len(recs)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_volumes_free(self._h)
            self._h = None

    def __len__(self):
        return len(self._regions)

    def region(self, i: int) -> dict:
        o = np.zeros(15, np.float32)
        lib().orc_volumes_region_info(self._h, i, _p(o))
        return dict(sigma_s=o[0:3].copy(), sigma_a=o[3:6].copy(), g=float(o[6]), majorant=float(o[7]), aabb_min=o[8:11].copy(),
                    aabb_max=o[11:14].copy(), homogeneous=bool(o[14]))

    def transmittance(self, origin, direction, t_eps: float, t_max: float, rng: Rng, n: int = 1) -> np.ndarray:
        out = np.zeros((n, 3), np.float32)
        lib().orc_volumes_transmittance_n(self._h, _p(_f3(origin)), _p(_f3(direction)), t_eps, t_max, C.byref(rng.state), n, _p(out))
        return out if n > 1 else out[0]

    def sample_interaction(self, origin, direction, t_eps: float, t_max: float, rng: Rng, n: int = 1) -> np.ndarray:
        """Rows: scatter, t, p[3], weight-or-transmittance[3], emitted[3], n_lobes, (w, g) x 2."""
        out = np.zeros((n, 16), np.float32)
        lib().orc_volumes_sample_n(self._h, _p(_f3(origin)), _p(_f3(direction)), t_eps, t_max, C.byref(rng.state), n, _p(out))
        return out

    def field_density(self, i: int, u) -> float:
        return float(lib().orc_volumes_field_density(self._h, i, _p(_f3(u))))

    def density(self, i: int, p_world) -> float:
        return float(lib().orc_volumes_density(self._h, i, _p(_f3(p_world))))

    def intersect(self, i: int, origin, direction):
        out = np.zeros(2, np.float32)
        if not lib().orc_volumes_intersect(self._h, i, _p(_f3(origin)), _p(_f3(direction)), _p(out)):
            return None
        return float(out[0]), float(out[1])


def vol_hash3(x: int, y: int, z: int, seed: int) -> float:
    return float(lib().orc_vol_hash3(x, y, z, seed))


def hg_phase(cos_theta: float, g: float) -> float:
    return float(lib().orc_hg_phase(cos_theta, g))


def phase_single_pdf(g: float, cos_theta: float) -> float:
    return float(lib().orc_phase_single_pdf(g, cos_theta))


def mat4_inverse(m) -> np.ndarray:
    """glam Mat4::inverse on a 4x4 (column-vector convention)."""
    a = np.ascontiguousarray(np.asarray(m, np.float32).reshape(4, 4).T.reshape(16))
    out = np.zeros(16, np.float32)
    lib().orc_mat4_inverse(_p(a), _p(out))
    return out.reshape(4, 4).T.copy()
