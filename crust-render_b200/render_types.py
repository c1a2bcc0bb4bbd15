"""Plain value types of the engine side of the seam, mirroring crust-core's public types
(crates/crust-core/src/{material/openpbr.rs, material/emissive.rs, light.rs, camera.rs, filter.rs,
tracer.rs, stats.rs}) as the records of include/crt_render.h. Shared by the CUDA binding
(render.py) and, in tests only, by the oracle binding. No rendering happens here."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import IntEnum
from typing import List, Optional, Sequence

import numpy as np

_V = ("<f4", 3)
# crt_material (openpbr.rs:66-128 order), 236 B
MATERIAL_DTYPE = np.dtype([
    ("kind", "<u4"), ("base_weight", "<f4"), ("base_color", *_V), ("base_diffuse_roughness", "<f4"), ("base_metalness", "<f4"),
    ("specular_weight", "<f4"), ("specular_color", *_V), ("specular_roughness", "<f4"), ("specular_ior", "<f4"),
    ("specular_roughness_anisotropy", "<f4"), ("transmission_weight", "<f4"), ("transmission_color", *_V),
    ("transmission_depth", "<f4"), ("transmission_scatter", *_V), ("transmission_scatter_anisotropy", "<f4"),
    ("transmission_dispersion_scale", "<f4"), ("transmission_dispersion_abbe_number", "<f4"), ("subsurface_weight", "<f4"),
    ("subsurface_color", *_V), ("subsurface_radius", "<f4"), ("subsurface_radius_scale", *_V),
    ("subsurface_scatter_anisotropy", "<f4"), ("fuzz_weight", "<f4"), ("fuzz_color", *_V), ("fuzz_roughness", "<f4"),
    ("coat_weight", "<f4"), ("coat_color", *_V), ("coat_roughness", "<f4"), ("coat_roughness_anisotropy", "<f4"),
    ("coat_ior", "<f4"), ("coat_darkening", "<f4"), ("thin_film_weight", "<f4"), ("thin_film_thickness", "<f4"),
    ("thin_film_ior", "<f4"), ("emission_luminance", "<f4"), ("emission_color", *_V), ("geometry_opacity", "<f4"),
    ("geometry_thin_walled", "<u4"), ("emissive_color", *_V),
])
LIGHT_DTYPE = np.dtype([("kind", "<u4"), ("geom_id", "<u4"), ("radiance", *_V), ("a", *_V), ("b", *_V), ("c", *_V), ("n", *_V),
                        ("r", "<f4"), ("s", "<f4"), ("env", "<i4"), ("w2l", "<f4", 9)])  # crt_light, 116 B
CAMERA_DTYPE = np.dtype([("lookfrom", *_V), ("lookat", *_V), ("vup", *_V), ("vfov_deg", "<f4"), ("aspect", "<f4"),
                         ("aperture", "<f4"), ("focus_dist", "<f4")])  # crt_camera
SETTINGS_DTYPE = np.dtype([("width", "<u4"), ("height", "<u4"), ("samples_per_pixel", "<u4"), ("max_depth", "<u4"),
                           ("min_samples_per_pixel", "<u4"), ("variance_threshold", "<f4"), ("frame", "<i4"),
                           ("sampling_strategy", "<u4"), ("adaptive", "<u4"), ("pixel_filter", "<u4"), ("filter_radius", "<f4"),
                           ("tile_size", "<u4"), ("tile_rank", "<u4"), ("tile_world", "<u4"), ("max_paths", "<u4"),
                           ("steal_chunk_tiles", "<u4")])  # crt_render_settings
STATS_DTYPE = np.dtype([(n, "<u8") for n in ("camera_rays", "closest_hit", "shadow_rays", "vertices", "rr_tested", "rr_killed",
                                              "ended_escaped", "ended_depth")])  # crt_ray_stats
# crt_volume (volume.rs:194-204 VolumeRegion::new arguments), 168 B with the grid pointer 8-aligned at 160
VOLUME_DTYPE = np.dtype([("local_to_world", "<f4", 16), ("half_extent", *_V), ("sigma_s", *_V), ("sigma_a", *_V), ("g", "<f4"),
                         ("emission", *_V), ("density_scale", "<f4"), ("field", "<u4"), ("noise_scale", "<f4"), ("noise_octaves", "<u4"),
                         ("noise_gain", "<f4"), ("noise_lacunarity", "<f4"), ("noise_threshold", "<f4"), ("noise_seed", "<u4"),
                         ("grid_dims", "<u4", 3), ("grid_data", "<u8")])
assert VOLUME_DTYPE.itemsize == 168 and VOLUME_DTYPE.fields["grid_data"][1] == 160
assert MATERIAL_DTYPE.itemsize == 236 and LIGHT_DTYPE.itemsize == 116 and CAMERA_DTYPE.itemsize == 52
assert SETTINGS_DTYPE.itemsize == 64 and STATS_DTYPE.itemsize == 64

MAT_OPENPBR, MAT_EMISSIVE = 0, 1
LIGHT_SPHERE, LIGHT_RECT, LIGHT_DISTANT, LIGHT_DOME = 0, 1, 2, 3
NO_GEOM = 0xFFFFFFFF

# openpbr.rs:130-173 `impl Default for OpenPBR`
_OPENPBR_DEFAULTS = dict(
    base_weight=1.0, base_color=(0.8, 0.8, 0.8), base_diffuse_roughness=0.0, base_metalness=0.0, specular_weight=1.0,
    specular_color=(1.0, 1.0, 1.0), specular_roughness=0.3, specular_ior=1.5, specular_roughness_anisotropy=0.0,
    transmission_weight=0.0, transmission_color=(1.0, 1.0, 1.0), transmission_depth=0.0, transmission_scatter=(0.0, 0.0, 0.0),
    transmission_scatter_anisotropy=0.0, transmission_dispersion_scale=0.0, transmission_dispersion_abbe_number=20.0,
    subsurface_weight=0.0, subsurface_color=(0.8, 0.8, 0.8), subsurface_radius=1.0, subsurface_radius_scale=(1.0, 0.5, 0.25),
    subsurface_scatter_anisotropy=0.0, fuzz_weight=0.0, fuzz_color=(1.0, 1.0, 1.0), fuzz_roughness=0.5, coat_weight=0.0,
    coat_color=(1.0, 1.0, 1.0), coat_roughness=0.0, coat_roughness_anisotropy=0.0, coat_ior=1.6, coat_darkening=1.0,
    thin_film_weight=0.0, thin_film_thickness=0.5, thin_film_ior=1.4, emission_luminance=0.0, emission_color=(1.0, 1.0, 1.0),
    geometry_opacity=1.0, geometry_thin_walled=0,
)


def OpenPBR(**params) -> np.ndarray:
    """`OpenPBR { ..Default::default() }` (openpbr.rs:66-173) as one crt_material record."""
    m = np.zeros((), dtype=MATERIAL_DTYPE)
    m["kind"] = MAT_OPENPBR
    for k, v in {**_OPENPBR_DEFAULTS, **params}.items():
        m[k] = v
    return m


def Emissive(color) -> np.ndarray:
    """`Emissive::new(color)` (emissive.rs:17-20)."""
    m = OpenPBR()
    m["kind"] = MAT_EMISSIVE
    m["emissive_color"] = color
    return m


def _norm(v):
    v = np.asarray(v, dtype=np.float32)
    return v * (np.float32(1.0) / np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1]) + np.float32(v[2] * v[2])))


def SphereLight(center, radius: float, radiance, geom_id: int) -> np.ndarray:
    """AreaLight over a SphereShape (light.rs:21-44, 150-160)."""
    l = np.zeros((), dtype=LIGHT_DTYPE)
    l["kind"], l["geom_id"], l["radiance"], l["a"], l["r"], l["env"] = LIGHT_SPHERE, geom_id, radiance, center, radius, -1
    return l


def RectLight(origin, edge_u, edge_v, normal, radiance, geom_id: int) -> np.ndarray:
    """AreaLight over `RectShape::new` (light.rs:55-81); the normal is normalised like the constructor does."""
    l = np.zeros((), dtype=LIGHT_DTYPE)
    l["kind"], l["geom_id"], l["radiance"] = LIGHT_RECT, geom_id, radiance
    l["a"], l["b"], l["c"], l["n"], l["env"] = origin, edge_u, edge_v, _norm(normal), -1
    return l


def DistantLight(direction, irradiance, angle_deg: float) -> np.ndarray:
    """`DistantLight::new` (light.rs:253-265)."""
    l = np.zeros((), dtype=LIGHT_DTYPE)
    diameter = np.float32(min(max(angle_deg, 0.05), 179.0))
    half = np.float32(0.5) * np.float32(np.radians(diameter, dtype=np.float32))
    cos_half = np.float32(np.cos(half, dtype=np.float32))
    l["kind"], l["geom_id"], l["radiance"], l["n"] = LIGHT_DISTANT, NO_GEOM, irradiance, _norm(direction)
    l["r"] = cos_half
    l["s"] = np.float32(2.0) * np.float32(math.pi) * (np.float32(1.0) - cos_half)
    l["env"] = -1
    return l


def DomeLight(tint) -> np.ndarray:
    """`DomeLight::new(tint, None, _)` — uniform dome (light.rs:320-389). Give it an environment map with
    `LightList.add(DomeLight(tint), environment=(pixels[h, w, 3], light_to_world 3x3))`."""
    l = np.zeros((), dtype=LIGHT_DTYPE)
    l["kind"], l["geom_id"], l["radiance"], l["env"] = LIGHT_DOME, NO_GEOM, tint, -1
    return l


FIELD_HOMOGENEOUS, FIELD_NOISE, FIELD_GRID = 0, 1, 2
MAX_VOLUMES = 16


class DensityField:
    """volume.rs:22-48. `Homogeneous()`, `Noise(scale, octaves, gain, lacunarity, threshold, seed)`, `Grid(nx, ny, nz, data)`."""

    def __init__(self, kind: int, **kw):
        self.kind, self.params = kind, kw

    @staticmethod
    def Homogeneous() -> "DensityField":
        return DensityField(FIELD_HOMOGENEOUS)

    @staticmethod
    def Noise(scale=4.0, octaves=4, gain=0.5, lacunarity=2.0, threshold=0.3, seed=0) -> "DensityField":  # defaults: usd_import.rs:440-447
        return DensityField(FIELD_NOISE, scale=scale, octaves=max(int(octaves), 1), gain=gain, lacunarity=lacunarity, threshold=threshold,
                            seed=int(seed) & 0xFFFFFFFF)

    @staticmethod
    def Grid(nx: int, ny: int, nz: int, data) -> "DensityField":
        data = np.ascontiguousarray(np.asarray(data, np.float32).reshape(-1))
        assert nx >= 1 and ny >= 1 and nz >= 1 and data.size == nx * ny * nz
        return DensityField(FIELD_GRID, dims=(nx, ny, nz), data=data)


class VolumeRegion:
    """`VolumeRegion::new(local_to_world, half_extent, sigma_s, sigma_a, g, emission, density_scale, field)`
    (volume.rs:194-231). `local_to_world` is a 4x4 matrix in the column-vector convention (world = M @ local)."""

    def __init__(self, local_to_world, half_extent, sigma_s, sigma_a, g: float, emission, density_scale: float, field: DensityField):
        self.field = field
        r = np.zeros((), dtype=VOLUME_DTYPE)
        r["local_to_world"] = np.asarray(local_to_world, np.float32).reshape(4, 4).T.reshape(16)  # glam column-major
        r["half_extent"], r["sigma_s"], r["sigma_a"], r["g"] = half_extent, sigma_s, sigma_a, g
        r["emission"], r["density_scale"], r["field"] = emission, density_scale, field.kind
        if field.kind == FIELD_NOISE:
            q = field.params
            r["noise_scale"], r["noise_octaves"], r["noise_gain"] = q["scale"], q["octaves"], q["gain"]
            r["noise_lacunarity"], r["noise_threshold"], r["noise_seed"] = q["lacunarity"], q["threshold"], q["seed"]
        elif field.kind == FIELD_GRID:
            r["grid_dims"] = field.params["dims"]
            r["grid_data"] = field.params["data"].ctypes.data  # kept alive by self.field
        self.rec = r

    def is_homogeneous(self) -> bool:  # volume.rs:274-276
        return self.field.kind == FIELD_HOMOGENEOUS


def volume_records(regions: Sequence["VolumeRegion"]) -> np.ndarray:
    out = np.zeros(len(regions), dtype=VOLUME_DTYPE)
    for i, v in enumerate(regions):
        out[i] = v.rec
    return out


class SamplingStrategy(IntEnum):  # tracer.rs:62-75
    PowerMis = 0
    BalanceMis = 1
    LightOnly = 2
    BsdfOnly = 3


class PixelFilterKind(IntEnum):  # filter.rs:28-45
    Box = 0
    Triangle = 1
    Gaussian = 2
    BlackmanHarris = 3
    Mitchell = 4


@dataclass
class PixelFilter:
    """filter.rs:28-60; the default is the importer's (triangle, radius 1; usd_import.rs:46-52)."""

    kind: PixelFilterKind = PixelFilterKind.Triangle
    radius: float = 1.0


@dataclass
class Camera:
    """`Camera::new` arguments (camera.rs:27-35)."""

    lookfrom: Sequence[float]
    lookat: Sequence[float]
    vup: Sequence[float]
    vfov: float
    aspect: float
    aperture: float = 0.0
    focus_dist: float = 1.0

    def record(self) -> np.ndarray:
        c = np.zeros((), dtype=CAMERA_DTYPE)
        c["lookfrom"], c["lookat"], c["vup"] = self.lookfrom, self.lookat, self.vup
        c["vfov_deg"], c["aspect"], c["aperture"], c["focus_dist"] = self.vfov, self.aspect, self.aperture, self.focus_dist
        return c


@dataclass
class RenderSettings:
    """`RenderSettings::new` + builders (tracer.rs:639-735)."""

    samples_per_pixel: int
    max_depth: int
    width: int
    height: int
    min_samples_per_pixel: int = 32
    variance_threshold: float = 0.05
    frame: int = 0
    sampling_strategy: SamplingStrategy = SamplingStrategy.PowerMis
    pixel_filter: PixelFilter = field(default_factory=PixelFilter)
    adaptive: bool = True
    # multi-GPU tile split and wavefront pool (crt_render.h)
    tile_size: int = 32
    tile_rank: int = 0
    tile_world: int = 1
    max_paths: int = 0
    steal_chunk_tiles: int = 0

    def with_samples_per_pixel(self, spp: int) -> "RenderSettings":
        self.samples_per_pixel = max(int(spp), 1)
        return self

    def with_sampling_strategy(self, s: SamplingStrategy) -> "RenderSettings":
        self.sampling_strategy = s
        return self

    def with_pixel_filter(self, f: PixelFilter) -> "RenderSettings":
        self.pixel_filter = f
        return self

    def get_dimensions(self):
        return (self.width, self.height)

    def record(self) -> np.ndarray:
        s = np.zeros((), dtype=SETTINGS_DTYPE)
        s["width"], s["height"] = self.width, self.height
        s["samples_per_pixel"], s["max_depth"], s["min_samples_per_pixel"] = self.samples_per_pixel, self.max_depth, self.min_samples_per_pixel
        s["variance_threshold"], s["frame"], s["sampling_strategy"] = self.variance_threshold, self.frame, int(self.sampling_strategy)
        s["adaptive"], s["pixel_filter"], s["filter_radius"] = int(self.adaptive), int(self.pixel_filter.kind), self.pixel_filter.radius
        s["tile_size"], s["tile_rank"], s["tile_world"], s["max_paths"] = self.tile_size, self.tile_rank, self.tile_world, self.max_paths
        s["steal_chunk_tiles"] = self.steal_chunk_tiles
        return s


@dataclass
class RayStats:
    """stats.rs:128-152."""

    camera_rays: int = 0
    closest_hit: int = 0
    shadow_rays: int = 0
    vertices: int = 0
    rr_tested: int = 0
    rr_killed: int = 0
    ended_escaped: int = 0
    ended_depth: int = 0

    def total_rays(self) -> int:
        return self.closest_hit + self.shadow_rays

    @staticmethod
    def from_record(r) -> "RayStats":
        return RayStats(*[int(r[n]) for n in STATS_DTYPE.names])


class LightList:
    """light.rs:404-446."""

    def __init__(self):
        self.lights: List[np.ndarray] = []
        self.environments = {}  # light index -> (pixels f32 [h, w, 3], light_to_world f32 [9], glam column order)

    def add(self, light: np.ndarray, environment=None) -> None:
        """`environment` = (lat-long pixels [h, w, 3], light_to_world 3x3 with columns = images of the axes) for a
        DomeLight (light.rs:327-335)."""
        if environment is not None:
            px, m = environment
            px = np.ascontiguousarray(np.asarray(px, np.float32))
            assert int(light["kind"]) == LIGHT_DOME and px.ndim == 3 and px.shape[2] == 3
            cols = np.ascontiguousarray(np.asarray(m, np.float32).reshape(3, 3).T.reshape(9))  # column-major
            self.environments[len(self.lights)] = (px, cols)
        self.lights.append(light)

    def count(self) -> int:
        return len(self.lights)

    def records(self) -> np.ndarray:
        out = np.zeros(len(self.lights), dtype=LIGHT_DTYPE)
        for i, l in enumerate(self.lights):
            out[i] = l
        return out


class WorldBuilderBase:
    """`WorldBuilder` (rt_world.rs:142-242): geometry plus a geom_id-indexed material table. The
    backend supplies `_new_scene_builder()`."""

    def __init__(self):
        self._sb = self._new_scene_builder()
        self._materials: List[np.ndarray] = []
        self._geom_material: List[int] = []

    def _intern(self, material: np.ndarray) -> int:
        key = material.tobytes()
        for i, m in enumerate(self._materials):
            if m.tobytes() == key:
                return i
        self._materials.append(material.copy())
        return len(self._materials) - 1

    def attach(self, geometry, material: np.ndarray) -> int:  # rt_world.rs:165-167
        return self.attach_masked(geometry, material, 0xFFFFFFFF)

    def attach_masked(self, geometry, material: np.ndarray, mask: int) -> int:  # :172-178
        gid = self._sb.attach_masked(geometry, mask)
        assert gid == len(self._geom_material)
        self._geom_material.append(self._intern(material))
        return gid

    def reserve_slot(self, material: np.ndarray, mask: int) -> int:  # :186-221
        return self.attach_masked(self._sb.empty_geometry(), material, mask)

    def set_geometry(self, geom_id: int, geometry) -> None:  # :226-230
        self._sb.set_geometry(geom_id, geometry)

    def count(self) -> int:
        return len(self._geom_material)

    def material_table(self):
        mats = np.zeros(max(len(self._materials), 1), dtype=MATERIAL_DTYPE)
        for i, m in enumerate(self._materials):
            mats[i] = m
        return mats[: len(self._materials)], np.asarray(self._geom_material, dtype=np.uint32)
