"""Synthetic scenes and ray batches for parity tests and bench.py.

Everything here is input generation (numpy, host): the reference's own probe
scenes (crates/crust-rt/examples/ray_throughput.rs:18-116), its unit-test
fixtures (crates/crust-rt/src/bvh.rs:1423-1458, 1686-1725) and the stress mesh
of scripts/gen_stress_scene.py scaled as SURVEY.md §8(d) "K1" describes.
Scene builders take an ``api`` namespace (anything exposing ``SceneBuilder``
and ``Geometry`` with the crust_rt shape) so the same scene can be committed
on the CUDA backend and, in tests, on the oracle.
"""
from __future__ import annotations

import math
import random
from typing import Tuple

import numpy as np

from .rt_types import RAY_DTYPE, Affine3A, Geometry


# ---------------------------------------------------------------------------
# LCG ray batches (ray_throughput.rs:103-116)
# ---------------------------------------------------------------------------
def lcg_stream(count: int, seed: int) -> np.ndarray:
    """state_k for k = 1..count of  s <- s * 1664525 + 1013904223 (mod 2^32), vectorised by doubling."""
    a = np.zeros(count, dtype=np.uint32)
    c = np.zeros(count, dtype=np.uint32)
    a[0], c[0] = 1664525, 1013904223
    m = 1
    with np.errstate(over="ignore"):
        while m < count:
            k = min(m, count - m)
            am, cm = a[m - 1], c[m - 1]
            a[m:m + k] = a[:k] * am
            c[m:m + k] = a[:k] * cm + c[:k]
            m += k
        return a * np.uint32(seed) + c


def ray_batch(count: int, extent: float, seed: int = 0x2545F491, t_min: float = 0.001, t_max: float = np.inf) -> np.ndarray:
    """The reference's probe batch: origins in a cube of side 4*extent, targets in
    a cube of side extent, normalised directions. Returns RAY_DTYPE[count]."""
    s = lcg_stream(6 * count, seed).reshape(count, 6)
    u = (s >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
    half = np.float32(0.5)
    origin = (u[:, 0:3] - half) * np.float32(4.0 * extent)
    target = (u[:, 3:6] - half) * np.float32(extent)
    d = target - origin
    # glam normalize: v * (1 / sqrt((x*x + y*y) + z*z)) in f32
    l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    inv = np.float32(1.0) / np.sqrt(l2, dtype=np.float32)
    d = d * inv[:, None]
    r = np.empty(count, dtype=RAY_DTYPE)
    r["o"], r["d"] = origin, d
    r["tmin"], r["tmax"] = np.float32(t_min), np.float32(t_max)
    return r


# ---------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------
def uv_sphere(center, radius: float, segs: int, rings: int) -> Tuple[np.ndarray, np.ndarray]:
    """ray_throughput.rs:18-48 (f32 trig like the Rust code)."""
    r = np.arange(rings + 1, dtype=np.float32)
    s = np.arange(segs + 1, dtype=np.float32)
    phi = (r / np.float32(rings)) * np.float32(math.pi)
    theta = (s / np.float32(segs)) * np.float32(2.0 * math.pi)
    sp, cp = np.sin(phi, dtype=np.float32), np.cos(phi, dtype=np.float32)
    st, ct = np.sin(theta, dtype=np.float32), np.cos(theta, dtype=np.float32)
    unit = np.stack([np.outer(sp, ct), np.outer(cp, np.ones_like(ct)), np.outer(sp, st)], axis=-1).astype(np.float32)
    verts = (np.asarray(center, np.float32) + np.float32(radius) * unit).reshape(-1, 3).astype(np.float32)
    row = segs + 1
    rr, ss = np.meshgrid(np.arange(rings), np.arange(segs), indexing="ij")
    a = (rr * row + ss).ravel()
    b = a + 1
    c = ((rr + 1) * row + ss + 1).ravel()
    d = ((rr + 1) * row + ss).ravel()
    idx = np.empty((2 * len(a), 3), dtype=np.uint32)
    idx[0::2] = np.stack([a, b, c], axis=1)
    idx[1::2] = np.stack([a, c, d], axis=1)
    return verts, idx


def triangle_scene(api):
    """ray_throughput.rs:50-66: 27 UV spheres (40x20) = 43 200 triangles, flat."""
    b = api.SceneBuilder()
    for x in (-1, 0, 1):
        for y in (-1, 0, 1):
            for z in (-1, 0, 1):
                v, i = uv_sphere(np.array([x, y, z], np.float32) * np.float32(2.5), 1.0, 40, 20)
                b.attach(api.Geometry.TriangleMesh(v, i))
    return b


def sphere_grid_scene(api):
    """ray_throughput.rs:68-82: 12^3 analytic spheres."""
    b = api.SceneBuilder()
    for x in range(12):
        for y in range(12):
            for z in range(12):
                c = np.array([x, y, z], np.float32) * np.float32(2.0) - np.float32(12.0)
                b.attach(api.Geometry.Sphere(c, 0.6))
    return b


def instance_scene(api, commit_kwargs=None):
    """ray_throughput.rs:84-101: 125 instances of one 576-triangle sphere."""
    inner = api.SceneBuilder()
    v, i = uv_sphere(np.zeros(3, np.float32), 1.0, 24, 12)
    inner.attach(api.Geometry.TriangleMesh(v, i))
    inner = inner.commit(**(commit_kwargs or {}))
    b = api.SceneBuilder()
    for x in range(-2, 3):
        for y in range(-2, 3):
            for z in range(-2, 3):
                t = np.array([x, y, z], np.float32) * np.float32(2.5)
                b.attach(api.Geometry.Instance(inner, Affine3A.from_translation(t)))
    return b


def sphere_grid_prims(api, n: int):
    """bvh.rs:1423-1438 test fixture: n^3 spheres, radius 0.5, spacing 3; geom_id = x*n*n + y*n + z."""
    b = api.SceneBuilder()
    for x in range(n):
        for y in range(n):
            for z in range(n):
                b.attach(api.Geometry.Sphere(np.array([x, y, z], np.float32) * np.float32(3.0), 0.5))
    return b


def diagonal_shards(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """bvh.rs:1443-1458: long thin diagonal triangles (spatial-split bait), one mesh."""
    o = np.arange(n, dtype=np.float32) * np.float32(0.35)
    v = np.empty((n, 3, 3), np.float32)
    v[:, 0] = np.stack([o, o, o], 1)
    v[:, 1] = np.stack([o + np.float32(10.0), o + np.float32(10.0), o + np.float32(10.2)], 1)
    v[:, 2] = np.stack([o + np.float32(10.0), o + np.float32(10.3), o + np.float32(10.0)], 1)
    return v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


# ---------------------------------------------------------------------------
# stress mesh (scripts/gen_stress_scene.py, scaled per SURVEY §8(d) K1)
# ---------------------------------------------------------------------------
RINGS, SEGS, SPACING, FIELD = 50, 100, 2.5, 16.0


def _round5(x: np.ndarray) -> np.ndarray:
    """The script prints coordinates with `%.5g`; reproduce the 5-significant-digit grid."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    nz = x != 0
    e = np.floor(np.log10(np.abs(x[nz])))
    scale = np.power(10.0, 4 - e)
    out[nz] = np.rint(x[nz] * scale) / scale
    return out


def _script_sphere(radius: float) -> Tuple[np.ndarray, np.ndarray]:
    """gen_stress_scene.py:31-56 (pole + rings + pole; 9 800 triangles at 50x100)."""
    r = np.arange(1, RINGS, dtype=np.float64)
    k = np.arange(SEGS, dtype=np.float64)
    phi = np.pi * r / RINGS
    y = radius * np.cos(phi)
    s = radius * np.sin(phi)
    th = 2.0 * np.pi * k / SEGS
    ring_pts = np.stack([np.outer(s, np.cos(th)), np.repeat(y[:, None], SEGS, 1), np.outer(s, np.sin(th))], -1).reshape(-1, 3)
    pts = np.concatenate([[[0.0, radius, 0.0]], ring_pts, [[0.0, -radius, 0.0]]], 0)
    last = len(pts) - 1

    def ring(rr, kk):
        return 1 + (rr - 1) * SEGS + (kk % SEGS)

    kk = np.arange(SEGS)
    top = np.stack([np.zeros(SEGS, np.int64), ring(1, kk + 1), ring(1, kk)], 1)
    rr, k2 = np.meshgrid(np.arange(1, RINGS - 1), kk, indexing="ij")
    a, b, c, d = ring(rr, k2), ring(rr, k2 + 1), ring(rr + 1, k2), ring(rr + 1, k2 + 1)
    quads = np.stack([a, b, d, a, d, c], -1).reshape(-1, 3)
    bot = np.stack([np.full(SEGS, last), ring(RINGS - 1, kk), ring(RINGS - 1, kk + 1)], 1)
    return pts, np.concatenate([top, quads, bot], 0).astype(np.uint32)


def stress_meshes(grid: int = 6, shards: int = 30000, jitter: bool = False, seed: int = 42):
    """Returns a list of (vertices f32 [n,3], indices u32 [m,3]) in the script's prim order:
    grid*grid spheres, the shard field, the floor, the occluder slab.

    ``grid=6, jitter=False`` is the scene as shipped (383 k triangles); ``grid=32,
    jitter=True`` is K1: every sphere gets its own radius 0.8*(1+0.01*u) so the importer's
    content hash would bake each one flat (10 035 200 sphere triangles + shards + 4)."""
    rng = random.Random(seed)
    shard_pts = []
    for _ in range(shards):  # gen_stress_scene.py:71-82, same draw order
        x = rng.uniform(-FIELD, FIELD)
        z = rng.uniform(-FIELD, FIELD)
        ln = rng.uniform(1.0, 3.0)
        dx, dz = rng.uniform(-1, 1), rng.uniform(-1, 1)
        n = math.hypot(dx, dz) or 1.0
        dx, dz = dx / n * ln, dz / n * ln
        h = rng.uniform(0.3, 1.2)
        shard_pts += [(x, 0.0, z), (x + dx, h, z + dz), (x + dx * 0.9 + 0.03, h, z + dz * 0.9)]
    meshes = []
    half = (grid - 1) * SPACING * 0.5
    base_pts, base_idx = _script_sphere(0.8)
    base32 = _round5(base_pts).astype(np.float32)
    for i in range(grid):
        for j in range(grid):
            if jitter:
                radius = 0.8 * (1.0 + 0.01 * rng.uniform(-1, 1))
                pts, _ = _script_sphere(radius)
                p32 = _round5(pts).astype(np.float32)
            else:
                p32 = base32
            t = np.array([i * SPACING - half, 0.8, j * SPACING - half], np.float32)
            meshes.append(((p32 + t).astype(np.float32), base_idx))
    sp = _round5(np.asarray(shard_pts, np.float64).reshape(-1, 3)).astype(np.float32)
    meshes.append((sp, np.arange(3 * shards, dtype=np.uint32).reshape(-1, 3)))
    quad = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)  # importer fan triangulation (usd_import.rs:1164-1214)
    meshes.append((np.array([(-40, 0, -40), (40, 0, -40), (40, 0, 40), (-40, 0, 40)], np.float32), quad))
    meshes.append((np.array([(-8, 6, -8), (8, 6, -8), (8, 6, 8), (-8, 6, 8)], np.float32), quad))
    return meshes


def stress_scene(api, grid: int = 6, shards: int = 30000, jitter: bool = False):
    """Flat (world-baked) stress scene: one geometry per USD mesh prim, ids in prim order."""
    b = api.SceneBuilder()
    n_tris = 0
    for v, i in stress_meshes(grid, shards, jitter):
        b.attach(api.Geometry.TriangleMesh(v, i))
        n_tris += len(i)
    return b, n_tris


def stress_scene_instanced(api, grid: int = 32, shards: int = 30000):
    """K2: the stress scene as the importer hands it over when every sphere authors the same points
    (no jitter) — content-hash dedup turns the grid*grid spheres into instances of ONE 9 800-triangle BLAS, each placed
    by its translate (usd_import.rs:769-798, 1025-1089); shards, floor and slab are placed once and stay flat."""
    from .rt_types import Affine3A

    meshes = stress_meshes(1, shards, False)  # [sphere, shards, floor, slab]
    base_pts, base_idx = _script_sphere(0.8)
    blas = api.SceneBuilder()
    blas.attach(api.Geometry.TriangleMesh(_round5(base_pts).astype(np.float32), base_idx))
    blas = blas.commit()
    b = api.SceneBuilder()
    half = (grid - 1) * SPACING * 0.5
    for i in range(grid):
        for j in range(grid):
            b.attach(api.Geometry.Instance(blas, Affine3A.from_translation((i * SPACING - half, 0.8, j * SPACING - half))))
    for v, idx in meshes[1:]:
        b.attach(api.Geometry.TriangleMesh(v, idx))
    return b


def scene_extent(bounds) -> float:
    """Half-diagonal of the scene bounds (the `extent` SURVEY §8(d) feeds ray_batch)."""
    d = np.asarray(bounds.maximum, np.float64) - np.asarray(bounds.minimum, np.float64)
    return float(0.5 * np.sqrt((d * d).sum()))


# ---------------------------------------------------------------------------
# ray sets (ii)-(iv) of SURVEY §8(d): camera, bounce and shadow rays for the stress scene
# ---------------------------------------------------------------------------
def _hash_u01(idx: np.ndarray, salt: int) -> np.ndarray:
    """Stateless 32-bit integer hash -> [0,1) f32 (input generation only; the
    integrator's Sobol sampler is a separate component)."""
    x = (idx.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(salt & 0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return (x >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)


class PinholeCamera:
    """crust-core/src/camera.rs:25-84 with the importer's USD mapping
    (scene/usd_import.rs:2174-2220), pinhole case (fStop 0). f32 throughout."""

    def __init__(self, lookfrom, forward, up, focal_length, horiz_aperture, width, height, focus_dist=10.0):
        f = np.float32
        lookfrom, forward, up = (np.asarray(v, f) for v in (lookfrom, forward, up))
        vert_aperture = f(horiz_aperture) * f(height) / f(width)
        vfov = f(2.0) * np.degrees(np.arctan(vert_aperture / (f(2.0) * f(focal_length))), dtype=f)
        theta = np.radians(vfov, dtype=f)
        h = np.tan(theta / f(2.0), dtype=f)
        vh = f(2.0) * h
        vw = f(width) / f(height) * vh

        def nrm(v):
            return (v / np.sqrt((v * v).sum(dtype=f), dtype=f)).astype(f)

        lookat = lookfrom + nrm(forward) * f(focus_dist)
        w = nrm(lookfrom - lookat)
        u = nrm(np.cross(nrm(up), w).astype(f))
        v = np.cross(w, u).astype(f)
        self.origin = lookfrom
        self.horizontal = (f(focus_dist) * vw * u).astype(f)
        self.vertical = (f(focus_dist) * vh * v).astype(f)
        self.lower_left = (lookfrom - self.horizontal / f(2.0) - self.vertical / f(2.0) - f(focus_dist) * w).astype(f)
        self.width, self.height = width, height

    def rays(self, spp: int, seed: int = 0, rows=None) -> np.ndarray:
        """Scanline-ordered camera rays (pixel-major, sample-minor); u = (i + fx) / w, v = (j + fy) / h
        (tracer.rs:574-575). `rows` = (j0, j1) restricts to a band of image rows."""
        j0, j1 = rows if rows is not None else (0, self.height)
        jj, ii, ss = np.meshgrid(np.arange(j0, j1), np.arange(self.width), np.arange(spp), indexing="ij")
        idx = ((jj * self.width + ii) * spp + ss).ravel()
        fx, fy = _hash_u01(idx, seed * 2 + 1), _hash_u01(idx, seed * 2 + 2)
        s = ((ii.ravel().astype(np.float32) + fx) / np.float32(self.width))[:, None]
        t = ((jj.ravel().astype(np.float32) + fy) / np.float32(self.height))[:, None]
        d = self.lower_left + s * self.horizontal + t * self.vertical - self.origin
        r = np.empty(len(idx), dtype=RAY_DTYPE)
        r["o"], r["d"] = self.origin, d.astype(np.float32)
        r["tmin"], r["tmax"] = np.float32(0.001), np.float32(np.inf)
        return r


def stress_camera(width: int, height: int) -> PinholeCamera:
    """gen_stress_scene.py:88-93: translate (0, 9, 20), rotateX -22 deg, focalLength 24, aperture 20.955."""
    a = math.radians(-22.0)
    forward = (0.0, math.sin(a), -math.cos(a))  # Rx(a) * (0, 0, -1)
    up = (0.0, math.cos(a), math.sin(a))        # Rx(a) * (0, 1, 0)
    return PinholeCamera((0.0, 9.0, 20.0), forward, up, 24.0, 20.955, width, height)


def bounce_rays(rays: np.ndarray, hits: np.ndarray, seed: int = 1) -> np.ndarray:
    """Set (iii): cosine-hemisphere directions (utils/src/common.rs:128-135) about the ray-facing
    normal at each hit point of `rays`; misses are dropped. t_min = 0.001 (tracer.rs:1152)."""
    f = np.float32
    ok = hits["geom_id"] != 0xFFFFFFFF
    o = (rays["o"][ok] + hits["t"][ok, None] * rays["d"][ok]).astype(f)
    n = hits["normal"][ok]
    idx = np.nonzero(ok)[0]
    u, v = _hash_u01(idx, seed * 2 + 101), _hash_u01(idx, seed * 2 + 102)
    z = np.sqrt(f(1.0) - v, dtype=f)
    phi = f(2.0 * math.pi) * u
    x, y = np.cos(phi, dtype=f) * np.sqrt(v, dtype=f), np.sin(phi, dtype=f) * np.sqrt(v, dtype=f)
    # any orthonormal frame around n (input generation; not the reference's align_to_normal)
    a = np.where(np.abs(n[:, :1]) > 0.9, np.array([[0, 1, 0]], f), np.array([[1, 0, 0]], f))
    t = np.cross(n, a)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    d = (x[:, None] * t + y[:, None] * b + z[:, None] * n).astype(f)
    r = np.empty(len(o), dtype=RAY_DTYPE)
    r["o"], r["d"] = o, d
    r["tmin"], r["tmax"] = f(0.001), f(np.inf)
    return r


def shadow_rays(rays: np.ndarray, hits: np.ndarray, seed: int = 2, half: float = 6.0, y: float = 12.0):
    """Set (iv): from each hit point of `rays` to a uniform point on the stress scene's 12x12
    RectLight; range [0.001, dist - 0.001] (tracer.rs:1016-1028). Returns (rays, mask=SHADOW)."""
    f = np.float32
    ok = hits["geom_id"] != 0xFFFFFFFF
    o = (rays["o"][ok] + hits["t"][ok, None] * rays["d"][ok]).astype(f)
    idx = np.nonzero(ok)[0]
    u, v = _hash_u01(idx, seed * 2 + 201), _hash_u01(idx, seed * 2 + 202)
    target = np.stack([(u * f(2.0) - f(1.0)) * f(half), np.full(len(o), y, f), (v * f(2.0) - f(1.0)) * f(half)], 1).astype(f)
    d = target - o
    dist = np.sqrt((d * d).sum(1, dtype=f), dtype=f)
    d = (d / dist[:, None]).astype(f)
    r = np.empty(len(o), dtype=RAY_DTYPE)
    r["o"], r["d"] = o, d
    r["tmin"], r["tmax"] = f(0.001), dist - f(0.001)
    return r, np.full(len(o), 2, np.uint32)
