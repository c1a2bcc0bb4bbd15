"""Host-side mirror of the `crust_rt` crate API (crates/crust-rt/src/lib.rs:45-53)
over the C ABI of ``libcrt_b200.so`` (include/crt.h): the same names, argument
meaning and error behaviour as the reference — ``SceneBuilder.attach`` →
``commit()`` → ``Scene.intersect`` / ``Scene.occluded`` — plus the batched
entry points a GPU needs.

There is no CPU execution path in this module: if the CUDA library or a device
is missing, commits and queries raise ``CrtError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .rt_types import (  # noqa: F401
    HIT_DTYPE,
    INVALID_ID,
    MASK_ALL,
    MASK_CAMERA,
    MASK_INDIRECT,
    MASK_SHADOW,
    RAY_DTYPE,
    Affine3A,
    CubicCurveSegment,
    CurveSegment,
    Geometry,
    PrimitiveBreakdown,
    Ray,
    RayHit,
    make_rays,
)

__all__ = [
    "AABB", "Affine3A", "CrtError", "CubicCurveSegment", "CurveSegment", "Geometry", "HIT_DTYPE", "INVALID_ID",
    "MASK_ALL", "MASK_CAMERA", "MASK_INDIRECT", "MASK_SHADOW", "PrimitiveBreakdown", "RAY_DTYPE", "Ray", "RayHit",
    "Scene", "SceneBuilder", "MemoryFootprint", "BuildInfo", "lib", "lib_path", "make_rays", "init", "device_info",
    "launch_count", "device_status", "device_status_reset", "trace_counters_enable", "trace_counters_read", "TRACE_COUNTER_NAMES", "l2_persist_info", "COMMIT_KEEP_HOST_TREE", "COMMIT_HOST_ONLY", "COMMIT_DEVICE_BUILD",
]

COMMIT_KEEP_HOST_TREE = 0x1
COMMIT_HOST_ONLY = 0x2
COMMIT_DEVICE_BUILD = 0x4  # build the BVH on the GPU (triangle-only scenes; see include/crt.h)

_STATUS = {1: "invalid argument", 2: "unknown geometry id", 3: "CUDA error", 4: "no CUDA device", 5: "instance nesting too deep", 6: "out of device memory"}


class CrtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"crt status {status} ({_STATUS.get(status, '?')}): {message}")
        self.status = status


def lib_path() -> str:
    """In-tree library; CRT_B200_LIB selects another build of the same sources (tuning harness only)."""
    return os.environ.get("CRT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcrt_b200.so")


class _PrimitiveBreakdown(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("triangles", "spheres", "curve_segments", "cubic_curve_spans", "instances")]


class MemoryFootprint(C.Structure):
    """scene.rs:34-57, restated for the device structures."""

    _fields_ = [(n, C.c_uint64) for n in ("bvh_nodes", "leaf_records", "prim_table", "descriptors")]

    def total(self) -> int:
        return self.bvh_nodes + self.leaf_records + self.prim_table + self.descriptors


class BuildInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primitives", "binary_nodes", "wide_nodes", "leaves", "leaf_refs", "node_bytes", "leaf_bytes")] + [
        ("build_seconds", C.c_double), ("upload_seconds", C.c_double), ("instance_depth", C.c_uint32), ("built_on_device", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    """Loads libcrt_b200.so (built by ``__graft_entry__.build()``); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise CrtError(4, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no Python or CPU fallback)")
    L = C.CDLL(path)
    vp, u32, u64, sz, f32p, u32p, u8p = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    L.crt_last_error.restype = C.c_char_p
    sig = {
        "crt_init": [C.c_int],
        "crt_device_info": [C.POINTER(C.c_int)] * 3 + [C.POINTER(u64)] * 2,
        "crt_l2_persist_info": [C.POINTER(u64)] * 3,
        "crt_builder_new": [C.POINTER(vp)],
        "crt_builder_reserve": [vp, sz],
        "crt_builder_count": [vp, C.POINTER(sz)],
        "crt_attach_triangle_mesh": [vp, vp, sz, vp, sz, vp, sz, u32, u32p],
        "crt_attach_sphere": [vp, f32p, C.c_float, u32, u32p],
        "crt_attach_round_curves": [vp, vp, sz, u32, u32p],
        "crt_attach_cubic_curves": [vp, vp, sz, u32, u32p],
        "crt_attach_instance": [vp, vp, f32p, f32p, u32, u32p],
        "crt_attach_empty": [vp, u32, u32p],
        "crt_set_triangle_mesh": [vp, u32, vp, sz, vp, sz, vp, sz],
        "crt_set_sphere": [vp, u32, f32p, C.c_float],
        "crt_set_round_curves": [vp, u32, vp, sz],
        "crt_set_cubic_curves": [vp, u32, vp, sz],
        "crt_set_instance": [vp, u32, vp, f32p, f32p],
        "crt_set_empty": [vp, u32],
        "crt_commit": [vp, C.POINTER(vp)],
        "crt_commit_ex": [vp, u32, C.c_int, C.POINTER(vp)],
        "crt_bounds": [vp, f32p, C.POINTER(C.c_int)],
        "crt_geometry_count": [vp, u32p],
        "crt_has_motion": [vp, C.POINTER(C.c_int)],
        "crt_primitive_count": [vp, C.POINTER(u64)],
        "crt_primitive_breakdown_of": [vp, C.POINTER(_PrimitiveBreakdown)],
        "crt_unique_primitive_breakdown": [vp, C.POINTER(_PrimitiveBreakdown)],
        "crt_primitive_extents": [vp, C.POINTER(u64), f32p, f32p, f32p],
        "crt_memory_footprint_of": [vp, C.POINTER(MemoryFootprint)],
        "crt_build_info_of": [vp, C.POINTER(BuildInfo)],
        "crt_intersect_n": [vp, sz, vp, vp, vp, vp, vp, vp],
        "crt_occluded_n": [vp, sz, vp, vp, vp, vp, vp],
        "crt_intersect_host_n": [vp, sz, vp, vp, vp, vp, vp],
        "crt_occluded_host_n": [vp, sz, vp, vp, vp, vp],
        "crt_intersect1": [vp, vp, C.c_float, u32, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "crt_occluded1": [vp, vp, C.c_float, u32, C.POINTER(C.c_int)],
        "crt_device_status": [u32p],
        "crt_device_status_reset": [],
        "crt_scene_export_size": [vp, C.POINTER(sz)],
        "crt_scene_export": [vp, vp, sz, C.c_int, vp],
        "crt_scene_import": [vp, sz, C.c_int, C.POINTER(vp), sz, C.POINTER(vp)],
        "crt_trace_counters_enable": [C.c_int],
        "crt_trace_counters_read": [C.POINTER(u64), C.c_int],
        "crt_debug_tree_sizes": [vp, C.POINTER(u64)],
        "crt_debug_tree_export": [vp, vp, u32p, u32p, u32p],
        "crt_debug_host_array": [vp, C.c_int, u32p, u64, C.POINTER(u64)],
        # include/crt_render.h
        "crt_world_new": [vp, vp, sz, vp, sz, vp, sz, C.POINTER(vp)],
        "crt_world_set_environment": [vp, u32, u32, u32, vp, vp],
        "crt_world_set_volumes": [vp, vp, sz],
        "crt_render": [vp, vp, vp, vp, vp, vp],
        "crt_render_host": [vp, vp, vp, vp, vp],
        "crt_render_last_info": [C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(u64)],
        "crt_steal_cursor_create": [C.POINTER(vp), vp],
        "crt_steal_cursor_open": [vp, C.POINTER(vp)],
        "crt_steal_cursor_reset": [vp],
        "crt_render_steal": [vp, vp, vp, vp, vp, vp, vp],
        "crt_shared_frame_create": [sz, C.POINTER(vp), vp],
        "crt_shared_frame_open": [vp, C.POINTER(vp)],
        "crt_shared_frame_close": [vp, C.c_int],
        "crt_debug_material_eval": [vp, sz, vp, vp, vp, vp, vp, vp],
        "crt_debug_material_scatter": [vp, sz, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name in ("crt_builder_free", "crt_scene_retain", "crt_scene_release", "crt_world_release", "crt_material_default", "crt_steal_cursor_close"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = None
    L.crt_launch_count.restype = u64
    _ = u8p
    _lib = L
    return L


u64_ = C.c_uint64


def l2_persist_info():
    """(bytes of L2 set aside for the trees' top levels, launches that got the access-policy window, launches refused)."""
    a, b, c = u64_(), u64_(), u64_()
    _ck(lib().crt_l2_persist_info(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def _ck(rc: int) -> None:
    if rc != 0:
        raise CrtError(rc, lib().crt_last_error().decode("utf-8", "replace"))


def init(device: int = 0) -> None:
    _ck(lib().crt_init(device))


def device_info() -> dict:
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    mem, l2 = C.c_uint64(), C.c_uint64()
    _ck(lib().crt_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem), C.byref(l2)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value, "l2_bytes": l2.value}


def launch_count() -> int:
    return int(lib().crt_launch_count())


def device_status_reset() -> None:
    _ck(lib().crt_device_status_reset())


TRACE_COUNTER_NAMES = ("rays", "nodes_fetched", "entries_culled_on_pop", "leaves_entered", "triangle_records", "other_records", "instance_descents")


def trace_counters_enable(on: bool = True) -> None:
    """crt_trace_counters_enable: traversal launches run the counting build of their kernel while enabled."""
    _ck(lib().crt_trace_counters_enable(1 if on else 0))


def trace_counters_read(reset: bool = True) -> dict:
    out = (C.c_uint64 * 7)()
    _ck(lib().crt_trace_counters_read(out, 1 if reset else 0))
    return dict(zip(TRACE_COUNTER_NAMES, (int(v) for v in out)))


def device_status() -> int:
    v = C.c_uint32()
    _ck(lib().crt_device_status(C.byref(v)))
    return v.value


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class AABB:
    """aabb.rs:4-8."""

    def __init__(self, minimum, maximum):
        self.minimum = np.asarray(minimum, dtype=np.float32)
        self.maximum = np.asarray(maximum, dtype=np.float32)


class SceneBuilder:
    """scene.rs:147-342. ``commit()`` consumes the builder, as in the reference."""

    def __init__(self):
        h = C.c_void_p()
        _ck(lib().crt_builder_new(C.byref(h)))
        self._h = h
        self._keep = []  # instanced Scene objects stay alive at least until commit
        self._inst = {}  # geom_id -> instanced Scene (what commit() will list as the scene's distinct children)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.crt_builder_free(self._h)
            self._h = None

    def _handle(self):
        if not self._h:
            raise CrtError(1, "SceneBuilder was consumed by commit()")
        return self._h

    def attach(self, geometry: Geometry) -> int:
        return self.attach_masked(geometry, MASK_ALL)

    def attach_masked(self, geometry: Geometry, mask: int) -> int:
        L, h, g = lib(), self._handle(), geometry
        gid = C.c_uint32()
        if g.kind == "mesh":
            n = g.normals
            _ck(L.crt_attach_triangle_mesh(h, _ptr(g.vertices), len(g.vertices), _ptr(g.indices), len(g.indices), _ptr(n),
                                           0 if n is None else len(n), mask, C.byref(gid)))
        elif g.kind == "sphere":
            _ck(L.crt_attach_sphere(h, _fptr(g.center), g.radius, mask, C.byref(gid)))
        elif g.kind == "curves":
            _ck(L.crt_attach_round_curves(h, _ptr(g.segments), len(g.segments), mask, C.byref(gid)))
        elif g.kind == "cubics":
            _ck(L.crt_attach_cubic_curves(h, _ptr(g.segments), len(g.segments), mask, C.byref(gid)))
        elif g.kind == "instance":
            end = None if g.transform_end is None else _fptr(g.transform_end.m)
            _ck(L.crt_attach_instance(h, g.scene._h, _fptr(g.transform.m), end, mask, C.byref(gid)))
            self._keep.append(g.scene)
            self._inst[gid.value] = g.scene
        elif g.kind == "empty":
            _ck(L.crt_attach_empty(h, mask, C.byref(gid)))
        else:
            raise CrtError(1, f"unknown geometry kind {g.kind!r}")
        return gid.value

    def reserve(self, additional: int) -> None:
        _ck(lib().crt_builder_reserve(self._handle(), additional))

    def count(self) -> int:
        n = C.c_size_t()
        _ck(lib().crt_builder_count(self._handle(), C.byref(n)))
        return n.value

    def set_geometry(self, id: int, geometry: Geometry) -> None:
        """scene.rs:199-201; an id that was never attached raises (the reference panics)."""
        L, h, g = lib(), self._handle(), geometry
        if g.kind == "mesh":
            n = g.normals
            _ck(L.crt_set_triangle_mesh(h, id, _ptr(g.vertices), len(g.vertices), _ptr(g.indices), len(g.indices), _ptr(n),
                                        0 if n is None else len(n)))
        elif g.kind == "sphere":
            _ck(L.crt_set_sphere(h, id, _fptr(g.center), g.radius))
        elif g.kind == "curves":
            _ck(L.crt_set_round_curves(h, id, _ptr(g.segments), len(g.segments)))
        elif g.kind == "cubics":
            _ck(L.crt_set_cubic_curves(h, id, _ptr(g.segments), len(g.segments)))
        elif g.kind == "instance":
            end = None if g.transform_end is None else _fptr(g.transform_end.m)
            _ck(L.crt_set_instance(h, id, g.scene._h, _fptr(g.transform.m), end))
            self._keep.append(g.scene)
        elif g.kind == "empty":
            _ck(L.crt_set_empty(h, id))
        else:
            raise CrtError(1, f"unknown geometry kind {g.kind!r}")
        if g.kind == "instance":
            self._inst[id] = g.scene
        else:
            self._inst.pop(id, None)

    @staticmethod
    def empty_geometry() -> Geometry:
        return Geometry.Empty()

    def distinct_children(self) -> list:
        """The instanced scenes in first-attachment order, each once: the order crt_commit lists a scene's children in and
        crt_scene_export / crt_scene_import name them by."""
        seen, out = set(), []
        for gid in sorted(self._inst):
            sc = self._inst[gid]
            if sc._h.value not in seen:
                seen.add(sc._h.value)
                out.append(sc)
        return out

    def discard(self) -> None:
        """Frees the builder without committing (a rank that receives the committed scene from another one)."""
        h = self._handle()
        self._h = None
        lib().crt_builder_free(h)

    def commit(self, flags: int = 0, host_threads: int = 0) -> "Scene":
        if _commit_override is not None:
            return _commit_override(self, flags, host_threads)
        return self.commit_local(flags, host_threads)

    def commit_local(self, flags: int = 0, host_threads: int = 0) -> "Scene":
        h = self._handle()
        self._h = None  # consumed, also on failure
        out = C.c_void_p()
        _ck(lib().crt_commit_ex(h, flags, host_threads, C.byref(out)))
        sc = Scene(out)
        sc._children = self.distinct_children()  # kept alive with the scene on the Python side too
        return sc


# crust_render_b200.distributed.shared_commits installs a function here: (builder, flags, host_threads) -> Scene
_commit_override = None


class Scene:
    """scene.rs:345-479. Immutable; queries run on the GPU."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # module globals are gone at interpreter shutdown
            _lib.crt_scene_release(self._h)
            self._h = None

    # -- single-ray calls in the reference's shape (scene.rs:354, 370) ---------
    def export_size(self) -> int:
        n = C.c_size_t()
        _ck(lib().crt_scene_export_size(self._h, C.byref(n)))
        return n.value

    def export_to(self, ptr: int, nbytes: int, on_device: bool, stream: int = 0) -> None:
        """crt_scene_export into host or device memory at `ptr`."""
        _ck(lib().crt_scene_export(self._h, C.c_void_p(ptr), nbytes, 1 if on_device else 0, C.c_void_p(stream)))

    @staticmethod
    def import_from(ptr: int, nbytes: int, on_device: bool, children=()) -> "Scene":
        """crt_scene_import: `children` are the scenes the exported one instances, in its distinct-children order."""
        arr = (C.c_void_p * max(1, len(children)))(*[c._h for c in children])
        out = C.c_void_p()
        _ck(lib().crt_scene_import(C.c_void_p(ptr), nbytes, 1 if on_device else 0, arr, len(children), C.byref(out)))
        sc = Scene(out)
        sc._children = list(children)
        return sc

    def intersect(self, ray: Ray, t_min: float, t_max: float) -> Optional[RayHit]:
        r = make_rays([ray.origin], [ray.dir], t_min, t_max)
        hit = np.zeros(1, dtype=HIT_DTYPE)
        ff, got = C.c_int(), C.c_int()
        _ck(lib().crt_intersect1(self._h, _ptr(r), np.float32(ray.time), ray.mask & 0xFFFFFFFF, _ptr(hit), C.byref(ff), C.byref(got)))
        if not got.value:
            return None
        h = hit[0]
        return RayHit(float(h["t"]), h["normal"].copy(), bool(ff.value), float(h["u"]), float(h["v"]), int(h["geom_id"]), int(h["prim_id"]))

    def occluded(self, ray: Ray, t_min: float, t_max: float) -> bool:
        r = make_rays([ray.origin], [ray.dir], t_min, t_max)
        occ = C.c_int()
        _ck(lib().crt_occluded1(self._h, _ptr(r), np.float32(ray.time), ray.mask & 0xFFFFFFFF, C.byref(occ)))
        return bool(occ.value)

    # -- batched, host buffers (numpy) ----------------------------------------
    def intersect_n(self, rays: np.ndarray, time: Optional[np.ndarray] = None, mask: Optional[np.ndarray] = None):
        """rays: RAY_DTYPE[n]. Returns (hits HIT_DTYPE[n], front_face uint8[n])."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        time = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        hits = np.empty(n, dtype=HIT_DTYPE)
        ff = np.empty(n, dtype=np.uint8)
        _ck(lib().crt_intersect_host_n(self._h, n, _ptr(rays), _ptr(time), _ptr(mask), _ptr(hits), _ptr(ff)))
        return hits, ff

    def occluded_n(self, rays: np.ndarray, time: Optional[np.ndarray] = None, mask: Optional[np.ndarray] = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        time = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        occ = np.empty(n, dtype=np.uint8)
        _ck(lib().crt_occluded_host_n(self._h, n, _ptr(rays), _ptr(time), _ptr(mask), _ptr(occ)))
        return occ

    # -- batched, raw pointers (device for *_n, host for *_host_n) ---------------
    def intersect_ptr(self, n: int, rays: int, time: int, mask: int, hits: int, front_face: int, stream: int = 0) -> None:
        _ck(lib().crt_intersect_n(self._h, n, rays or None, time or None, mask or None, hits or None, front_face or None, stream or None))

    def occluded_ptr(self, n: int, rays: int, time: int, mask: int, occluded: int, stream: int = 0) -> None:
        _ck(lib().crt_occluded_n(self._h, n, rays or None, time or None, mask or None, occluded or None, stream or None))

    def intersect_host_ptr(self, n: int, rays: int, time: int, mask: int, hits: int, front_face: int) -> None:
        _ck(lib().crt_intersect_host_n(self._h, n, rays or None, time or None, mask or None, hits or None, front_face or None))

    def occluded_host_ptr(self, n: int, rays: int, time: int, mask: int, occluded: int) -> None:
        _ck(lib().crt_occluded_host_n(self._h, n, rays or None, time or None, mask or None, occluded or None))

    # -- scene.rs:375-472 --------------------------------------------------------
    def bounds(self) -> Optional[AABB]:
        o = (C.c_float * 6)()
        ne = C.c_int()
        _ck(lib().crt_bounds(self._h, o, C.byref(ne)))
        return AABB(o[0:3], o[3:6]) if ne.value else None

    def geometry_count(self) -> int:
        v = C.c_uint32()
        _ck(lib().crt_geometry_count(self._h, C.byref(v)))
        return v.value

    def has_motion(self) -> bool:
        v = C.c_int()
        _ck(lib().crt_has_motion(self._h, C.byref(v)))
        return bool(v.value)

    def primitive_count(self) -> int:
        v = C.c_uint64()
        _ck(lib().crt_primitive_count(self._h, C.byref(v)))
        return v.value

    def _breakdown(self, fn) -> PrimitiveBreakdown:
        b = _PrimitiveBreakdown()
        _ck(fn(self._h, C.byref(b)))
        return PrimitiveBreakdown(b.triangles, b.spheres, b.curve_segments, b.cubic_curve_spans, b.instances)

    def primitive_breakdown(self) -> PrimitiveBreakdown:
        return self._breakdown(lib().crt_primitive_breakdown_of)

    def unique_primitive_breakdown(self) -> PrimitiveBreakdown:
        return self._breakdown(lib().crt_unique_primitive_breakdown)

    def primitive_extents(self):
        n = C.c_uint64()
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _ck(lib().crt_primitive_extents(self._h, C.byref(n), C.byref(a), C.byref(b), C.byref(c)))
        return n.value, a.value, b.value, c.value

    def memory_footprint(self) -> MemoryFootprint:
        m = MemoryFootprint()
        _ck(lib().crt_memory_footprint_of(self._h, C.byref(m)))
        return m

    def build_info(self) -> BuildInfo:
        b = BuildInfo()
        _ck(lib().crt_build_info_of(self._h, C.byref(b)))
        return b

    # -- test support ------------------------------------------------------------
    def debug_tree(self):
        """Canonical host-tree export (needs COMMIT_KEEP_HOST_TREE): dict with
        nodes [n,32] u32 words, tri_counts, other_counts, refs."""
        sizes = (C.c_uint64 * 4)()
        _ck(lib().crt_debug_tree_sizes(self._h, sizes))
        n_nodes, n_leaves, n_refs, n_bin = (int(x) for x in sizes)
        nodes = np.zeros((n_nodes, 32), dtype=np.uint32)
        tri = np.zeros(n_leaves, dtype=np.uint32)
        oth = np.zeros(n_leaves, dtype=np.uint32)
        refs = np.zeros(n_refs, dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        _ck(lib().crt_debug_tree_export(self._h, _ptr(nodes), tri.ctypes.data_as(u32p), oth.ctypes.data_as(u32p), refs.ctypes.data_as(u32p)))
        return {"nodes": nodes, "tri_counts": tri, "other_counts": oth, "refs": refs, "binary_nodes": n_bin}

    def debug_host_arrays(self):
        """The flattened host arrays word for word (needs COMMIT_KEEP_HOST_TREE): wide nodes, leaf records, motion pairs,
        instance record offsets."""
        out = {}
        u32p = C.POINTER(C.c_uint32)
        for which, name in enumerate(("nodes", "leafs", "motion", "instance_record_q")):
            n = C.c_uint64(0)
            _ck(lib().crt_debug_host_array(self._h, which, None, 0, C.byref(n)))
            a = np.zeros(int(n.value), dtype=np.uint32)
            if n.value:
                _ck(lib().crt_debug_host_array(self._h, which, a.ctypes.data_as(u32p), n.value, C.byref(n)))
            out[name] = a
        return out
