"""Engine side of the seam over libcrt_b200.so: `WorldBuilder` / `World` (crates/crust-core/src/
rt_world.rs:142-293) and `Renderer` (tracer.rs:136-212), bound to include/crt_render.h. The render
itself is the CUDA wavefront integrator (csrc/render.cu); there is no host fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import rt
from .render_types import (  # noqa: F401
    LIGHT_DTYPE,
    MATERIAL_DTYPE,
    STATS_DTYPE,
    Camera,
    DensityField,
    DistantLight,
    DomeLight,
    Emissive,
    LightList,
    OpenPBR,
    PixelFilter,
    PixelFilterKind,
    RayStats,
    RectLight,
    RenderSettings,
    SamplingStrategy,
    SphereLight,
    VolumeRegion,
    WorldBuilderBase,
    volume_records,
)

__all__ = ["World", "WorldBuilder", "Renderer", "StealCursor", "SharedFrame", "Camera", "RenderSettings", "SamplingStrategy", "PixelFilter", "PixelFilterKind",
           "RayStats", "LightList", "OpenPBR", "Emissive", "SphereLight", "RectLight", "DistantLight", "DomeLight", "VolumeRegion", "DensityField", "debug_material_eval", "debug_material_scatter"]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class World:
    """rt_world.rs:244-293 (+ the LightList, which the reference hands to the Renderer separately)."""

    def __init__(self, scene: "rt.Scene", materials: np.ndarray, geom_material: np.ndarray, lights: LightList):
        self.scene = scene
        self.materials, self.geom_material = materials, geom_material  # the geom_id-indexed table (rt_world.rs:142-160)
        self._lights = lights.records()
        h = C.c_void_p()
        rt._ck(rt.lib().crt_world_new(scene._h, _p(materials), len(materials), _p(geom_material), len(geom_material), _p(self._lights),
                                      len(self._lights), C.byref(h)))
        self._h = h
        for index, (px, cols) in lights.environments.items():
            rt._ck(rt.lib().crt_world_set_environment(h, index, px.shape[1], px.shape[0], _p(px), _p(cols)))

    def __del__(self):
        L = getattr(rt, "_lib", None) if rt is not None else None  # module globals are gone at interpreter shutdown
        if getattr(self, "_h", None) and L is not None:
            L.crt_world_release(self._h)
            self._h = None

    def with_volumes(self, regions) -> "World":
        """`Scene::with_volumes` (tracer.rs:167-170): free-standing participating media, outside the BVH."""
        self._volumes = list(regions)  # keeps grid voxels alive during the call
        recs = volume_records(self._volumes)
        rt._ck(rt.lib().crt_world_set_volumes(self._h, _p(recs) if len(recs) else None, len(recs)))
        return self

    def count(self) -> int:
        return self.scene.geometry_count()

    def has_motion(self) -> bool:
        return self.scene.has_motion()

    def bounds(self):
        return self.scene.bounds()

    def primitive_count(self) -> int:
        return self.scene.primitive_count()


class WorldBuilder(WorldBuilderBase):
    def _new_scene_builder(self):
        return rt.SceneBuilder()

    def commit(self, lights: Optional[LightList] = None) -> World:  # rt_world.rs:235-241
        mats, gm = self.material_table()
        return World(self._sb.commit(), mats, gm, lights or LightList())


class StealCursor:
    """The shared tile cursor of crt_render_steal (include/crt_render.h): created by one process, opened by the
    others from its 64-byte CUDA IPC handle."""

    def __init__(self, handle: Optional[bytes] = None):
        h = C.c_void_p()
        if handle is None:
            buf = (C.c_ubyte * 64)()
            rt._ck(rt.lib().crt_steal_cursor_create(C.byref(h), buf))
            self.handle, self.owner = bytes(buf), True
        else:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle)
            rt._ck(rt.lib().crt_steal_cursor_open(buf, C.byref(h)))
            self.handle, self.owner = bytes(handle), False
        self._h = h

    def reset(self) -> None:
        rt._ck(rt.lib().crt_steal_cursor_reset(self._h))

    def __del__(self):
        L = getattr(rt, "_lib", None) if rt is not None else None
        if getattr(self, "_h", None) and L is not None:
            L.crt_steal_cursor_close(self._h)
            self._h = None


class SharedFrame:
    """The final frame in the owning process's HBM, mapped by the peers through CUDA IPC (crt_shared_frame_*): every
    process's film kernel stores its tiles straight into it over NVLink / NVSwitch peer access."""

    def __init__(self, height: int, width: int, handle: Optional[bytes] = None):
        self.height, self.width = height, width
        self.nbytes = height * width * 16
        p = C.c_void_p()
        if handle is None:
            buf = (C.c_ubyte * 64)()
            rt._ck(rt.lib().crt_shared_frame_create(self.nbytes, C.byref(p), buf))
            self.handle, self.owner = bytes(buf), True
        else:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle)
            rt._ck(rt.lib().crt_shared_frame_open(buf, C.byref(p)))
            self.handle, self.owner = bytes(handle), False
        self.ptr = p.value
        # lets torch.as_tensor(frame, device="cuda") view the memory without copying
        self.__cuda_array_interface__ = {"shape": (height, width, 4), "typestr": "<f4", "data": (self.ptr, False), "version": 3, "strides": None}

    def close(self) -> None:
        L = getattr(rt, "_lib", None) if rt is not None else None
        if getattr(self, "ptr", None) and L is not None:
            L.crt_shared_frame_close(C.c_void_p(self.ptr), int(self.owner))
            self.ptr = None

    def __del__(self):
        self.close()


class Renderer:
    """`Renderer::new(camera, world, lights, settings)` (tracer.rs:149-163); lights live in the World here."""

    def __init__(self, camera: Camera, world: World, settings: RenderSettings):
        self.camera, self.world, self.settings = camera, world, settings
        self.device_ms = 0.0
        self.launches = 0
        self.rounds = 0

    def _info(self):
        n, ms, r = C.c_uint64(), C.c_double(), C.c_uint64()
        rt.lib().crt_render_last_info(C.byref(n), C.byref(ms), C.byref(r))
        self.launches, self.device_ms, self.rounds = n.value, ms.value, r.value

    def render_with_stats(self, out: Optional[np.ndarray] = None):
        """tracer.rs:194-200. Returns (framebuffer[h, w, 4] with row 0 = bottom row, RayStats)."""
        st = self.settings
        fb = out if out is not None else np.zeros((st.height, st.width, 4), dtype=np.float32)
        assert fb.dtype == np.float32 and fb.shape == (st.height, st.width, 4) and fb.flags.c_contiguous
        stats = np.zeros((), dtype=STATS_DTYPE)
        cam, rec = self.camera.record(), st.record()
        rt._ck(rt.lib().crt_render_host(self.world._h, _p(cam), _p(rec), _p(fb), _p(stats)))
        self._info()
        return fb, RayStats.from_record(stats)

    def render(self) -> np.ndarray:  # tracer.rs:172-174
        return self.render_with_stats()[0]

    def render_steal(self, fb_ptr: int, cursor: StealCursor, stream: int = 0) -> RayStats:
        """crt_render_steal: renders the chunks of tiles this process wins off the shared cursor."""
        stats = np.zeros((), dtype=STATS_DTYPE)
        cam, rec = self.camera.record(), self.settings.record()
        rt._ck(rt.lib().crt_render_steal(self.world._h, _p(cam), _p(rec), C.c_void_p(fb_ptr), _p(stats), C.c_void_p(stream), cursor._h))
        self._info()
        return RayStats.from_record(stats)

    def render_device(self, fb_ptr: int, stream: int = 0) -> RayStats:
        """Same with a device framebuffer (width*height float4) and a CUDA stream."""
        stats = np.zeros((), dtype=STATS_DTYPE)
        cam, rec = self.camera.record(), self.settings.record()
        rt._ck(rt.lib().crt_render(self.world._h, _p(cam), _p(rec), C.c_void_p(fb_ptr), _p(stats), C.c_void_p(stream)))
        self._info()
        return RayStats.from_record(stats)


def debug_material_eval(material: np.ndarray, ray_dir, normal, front_face, wi):
    """crt_debug_material_eval: `Material::eval` of one crt_material record on the device for n queries.
    Returns (value[n, 3] = brdf * |cos|, pdf[n], some[n] bool)."""
    rd, nn, w = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (ray_dir, normal, wi))
    ff = np.ascontiguousarray(front_face, dtype=np.uint8).reshape(-1)
    n = len(rd)
    out, some = np.zeros((n, 4), np.float32), np.zeros(n, np.uint8)
    mat = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
    rt._ck(rt.lib().crt_debug_material_eval(_p(mat), n, _p(rd), _p(nn), _p(ff), _p(w), _p(out), _p(some)))
    return out[:, :3], out[:, 3], some.astype(bool)


def debug_material_scatter(material: np.ndarray, ray_dir, normal, front_face, s4):
    """crt_debug_material_scatter: `Material::scatter_importance` given the 4-D BSDF block of each query.
    Returns (direction[n, 3], value[n, 3], pdf[n], delta[n] bool, some[n] bool)."""
    rd, nn = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (ray_dir, normal))
    s = np.ascontiguousarray(s4, dtype=np.float32).reshape(-1, 4)
    ff = np.ascontiguousarray(front_face, dtype=np.uint8).reshape(-1)
    n = len(rd)
    out, some = np.zeros((n, 8), np.float32), np.zeros(n, np.uint8)
    mat = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
    rt._ck(rt.lib().crt_debug_material_scatter(_p(mat), n, _p(rd), _p(nn), _p(ff), _p(s), _p(out), _p(some)))
    return out[:, :3], out[:, 3:6], out[:, 6], out[:, 7] != 0, some.astype(bool)
