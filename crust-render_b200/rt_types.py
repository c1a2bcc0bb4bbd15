"""Plain-Python value types mirroring the `crust_rt` crate's public types
(crates/crust-rt/src/{ray.rs, scene.rs}); shared by the CUDA binding (rt.py)
and, in tests only, by the oracle binding. No computation happens here."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

# ray.rs:8-11, lib.rs:53
MASK_CAMERA = 1 << 0
MASK_SHADOW = 1 << 1
MASK_INDIRECT = 1 << 2
MASK_ALL = 0xFFFFFFFF
INVALID_ID = 0xFFFFFFFF

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])  # crt_ray, 32 B
HIT_DTYPE = np.dtype(
    [("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("geom_id", "<u4"), ("prim_id", "<u4"), ("normal", "<f4", 3)]
)  # crt_hit, 32 B
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 32


def _f32(x):
    return np.float32(x)


def _v3(v) -> np.ndarray:
    a = np.asarray(v, dtype=np.float32).reshape(3)
    return a


@dataclass
class Ray:
    """ray.rs:17-55."""

    origin: Sequence[float]
    dir: Sequence[float]
    time: float = 0.0
    mask: int = MASK_ALL

    def with_time(self, time: float) -> "Ray":
        return Ray(self.origin, self.dir, time, self.mask)

    def with_mask(self, mask: int) -> "Ray":
        return Ray(self.origin, self.dir, self.time, mask)

    def at(self, t: float) -> np.ndarray:
        return _v3(self.origin) + np.float32(t) * _v3(self.dir)


@dataclass
class RayHit:
    """scene.rs:134-142."""

    t: float
    normal: np.ndarray
    front_face: bool
    u: float
    v: float
    geom_id: int
    prim_id: int


@dataclass
class PrimitiveBreakdown:
    """scene.rs:62-68."""

    triangles: int = 0
    spheres: int = 0
    curve_segments: int = 0
    cubic_curve_spans: int = 0
    instances: int = 0


class Affine3A:
    """glam `Affine3A`: 3 matrix columns + translation, f32. Only the
    constructors the reference's tests and importer use."""

    def __init__(self, cols12: Sequence[float]):
        self.m = np.asarray(cols12, dtype=np.float32).reshape(12).copy()

    IDENTITY: "Affine3A"

    @staticmethod
    def from_translation(t) -> "Affine3A":
        return Affine3A([1, 0, 0, 0, 1, 0, 0, 0, 1, *map(float, t)])

    @staticmethod
    def from_scale(s) -> "Affine3A":
        return Affine3A([s[0], 0, 0, 0, s[1], 0, 0, 0, s[2], 0, 0, 0])

    @staticmethod
    def from_rotation_z(a: float) -> "Affine3A":
        s, c = np.float32(math.sin(np.float32(a))), np.float32(math.cos(np.float32(a)))
        return Affine3A([c, s, 0, -s, c, 0, 0, 0, 1, 0, 0, 0])

    @staticmethod
    def from_rotation_y(a: float) -> "Affine3A":
        s, c = np.float32(math.sin(np.float32(a))), np.float32(math.cos(np.float32(a)))
        return Affine3A([c, 0, -s, 0, 1, 0, s, 0, c, 0, 0, 0])

    @staticmethod
    def from_rotation_x(a: float) -> "Affine3A":
        s, c = np.float32(math.sin(np.float32(a))), np.float32(math.cos(np.float32(a)))
        return Affine3A([1, 0, 0, 0, c, s, 0, -s, c, 0, 0, 0])

    def matrix3(self) -> np.ndarray:
        """3x3 with columns = x_axis, y_axis, z_axis."""
        return self.m[:9].reshape(3, 3).T.copy()

    def __mul__(self, other: "Affine3A") -> "Affine3A":
        """self ∘ other (apply `other` first), f32."""
        a, b = self.matrix3(), other.matrix3()
        m = (a @ b).astype(np.float32)
        t = (a @ other.m[9:12] + self.m[9:12]).astype(np.float32)
        return Affine3A([*m.T.reshape(9), *t])


Affine3A.IDENTITY = Affine3A([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0])


@dataclass
class CurveSegment:
    """scene.rs:18-23."""

    p0: Sequence[float]
    p1: Sequence[float]
    r0: float
    r1: float


@dataclass
class CubicCurveSegment:
    """scene.rs:76-80."""

    cp: Sequence[Sequence[float]]
    r0: float
    r1: float


@dataclass
class Geometry:
    """scene.rs:87-122 — tagged union; build with the static constructors."""

    kind: str
    vertices: Optional[np.ndarray] = None
    indices: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None
    center: Optional[np.ndarray] = None
    radius: float = 0.0
    segments: Optional[np.ndarray] = None  # curves: [n, 8]; cubics: [n, 14]
    scene: object = None
    transform: Optional[Affine3A] = None
    transform_end: Optional[Affine3A] = None

    @staticmethod
    def TriangleMesh(vertices, indices, normals=None) -> "Geometry":
        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
        i = np.ascontiguousarray(np.asarray(indices, dtype=np.uint32).reshape(-1, 3))
        n = None if normals is None else np.ascontiguousarray(np.asarray(normals, dtype=np.float32).reshape(-1, 3))
        return Geometry("mesh", vertices=v, indices=i, normals=n)

    @staticmethod
    def Sphere(center, radius) -> "Geometry":
        return Geometry("sphere", center=_v3(center), radius=float(radius))

    @staticmethod
    def RoundCurves(segments) -> "Geometry":
        rows = [[*s.p0, *s.p1, s.r0, s.r1] if isinstance(s, CurveSegment) else list(s) for s in segments]
        return Geometry("curves", segments=np.asarray(rows, dtype=np.float32).reshape(-1, 8))

    @staticmethod
    def CubicCurves(segments) -> "Geometry":
        rows = [
            [*np.asarray(s.cp, dtype=np.float32).reshape(12), s.r0, s.r1] if isinstance(s, CubicCurveSegment) else list(s)
            for s in segments
        ]
        return Geometry("cubics", segments=np.asarray(rows, dtype=np.float32).reshape(-1, 14))

    @staticmethod
    def Instance(scene, transform: Affine3A, transform_end: Optional[Affine3A] = None) -> "Geometry":
        return Geometry("instance", scene=scene, transform=transform, transform_end=transform_end)

    @staticmethod
    def Empty() -> "Geometry":
        """SceneBuilder::empty_geometry (scene.rs:205-211)."""
        return Geometry("empty")


def make_rays(origins, dirs, tmin=0.001, tmax=np.inf) -> np.ndarray:
    """Pack [n,3] origins/dirs (+ scalar or [n] ranges) into crt_ray records."""
    o = np.asarray(origins, dtype=np.float32).reshape(-1, 3)
    d = np.asarray(dirs, dtype=np.float32).reshape(-1, 3)
    r = np.empty(o.shape[0], dtype=RAY_DTYPE)
    r["o"], r["d"] = o, d
    r["tmin"] = np.asarray(tmin, dtype=np.float32)
    r["tmax"] = np.asarray(tmax, dtype=np.float32)
    return r
