"""USD stage -> (World, Camera, RenderSettings): the decisions of crust-core's importer
(crates/crust-core/src/scene/usd_import.rs) restated over the text reader in usda.py.

What is reproduced, with the reference lines it follows:
  * stage walk order — a LIFO stack, so siblings are visited in REVERSE authored order (:140-269);
    stages whose root (or single root's children) has >= 4 subtrees are imported subtree by subtree
    in authored order (`stream_roots`, :67-115) — this fixes the geom_id and light order;
  * xformOp composition in f32, ops multiplied in xformOpOrder order (:563-650);
  * meshes: fan triangulation (:1164-1214), content + material dedup (:769-798), "placed once and
    static => baked to world space, else instanced over a shared BLAS" (:1025-1089), mirrored
    placements swap indices 1 and 2 (:1109-1116), `crust:motion:translate` (:755-759), `crust:rayMask`;
  * spheres (:1235-1272), SphereLight / RectLight geometry + AreaLight sharing a geom_id with the
    SHADOW|INDIRECT[|CAMERA] mask (:747-753, 2259-2349), DistantLight, DomeLight (uniform, or a lat-long
    environment map decoded by the host and importance-sampled, :2389-2460);
  * camera from focal length / apertures / fStop / focusDistance (:2174-2220);
  * materials: `crust:openpbr` decoded 1:1 (:2902-2975), `UsdPreviewSurface` mapped (:2661-2705),
    unbound => default grey OpenPBR; RenderSettings and the crust:* attributes (:3015-3100).
  * PointInstancer (:1674-1875): T * R * S placements, one Instance per prototype part, `invisibleIds`;
    binary crate files through usdc.py; a PointInstancer inside a prototype nests as sub-scenes instead of flattening
    (`nested_instancer_parts`, :1549-1614);
  * native instancing (:166-188, 1616-1635): an `instanceable` prim with an internal reference / inherit / specialize
    arc places the shared prototype's parts and its own subtree is never walked; a native instance nested inside a
    prototype is skipped exactly like the reference has to (:1418-1440);
  * BasisCurves (:1936-2172): linear -> round cone segments, cubic bezier / bspline / catmullRom -> one Bezier-form span
    per authored span, widths by vertex / curve / constant, as one Instance over its own committed scene;
  * volumes: a prim carrying `crust:volume:type` (homogeneous | smoke | grid) imports as a VolumeRegion only, never
    as geometry (:190-209, 427-503).
Composition across files (subLayers, variant selections, references and payloads into other layers) happens before the
import, in usd/compose.py (the reference gets it from openusd, :271-300). Not reproduced: Ptex (internal arcs resolve to prims
of the same layer, which covers every bundled sample).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from ..render_types import (
    Camera,
    DensityField,
    DistantLight,
    DomeLight,
    Emissive,
    LightList,
    OpenPBR,
    PixelFilter,
    PixelFilterKind,
    RectLight,
    RenderSettings,
    SamplingStrategy,
    SphereLight,
    VolumeRegion,
)
from ..rt_types import MASK_ALL, MASK_CAMERA, MASK_INDIRECT, MASK_SHADOW, Affine3A, CubicCurveSegment, CurveSegment, Geometry
from . import usda

F = np.float32
# usd_import.rs:46-52
DEFAULTS = dict(width=640, height=360, spp=128, max_depth=32, min_spp=32, variance=0.05, frame=0)
MIN_STREAM_CHUNKS = 4
MAX_INSTANCE_NESTING = 8  # usd_import.rs:1277

# cubic basis matrices, row-major [t^3 t^2 t 1] . M . [P0 P1 P2 P3]^T (:1912-1931)
_BASES = {
    "bezier": (np.asarray([[-1, 3, -3, 1], [3, -6, 3, 0], [-3, 3, 0, 0], [1, 0, 0, 0]], np.float32), 3),
    "bspline": (np.asarray([[-1 / 6, 3 / 6, -3 / 6, 1 / 6], [3 / 6, -6 / 6, 3 / 6, 0], [-3 / 6, 0, 3 / 6, 0], [1 / 6, 4 / 6, 1 / 6, 0]], np.float32), 1),
    "catmullRom": (np.asarray([[-0.5, 1.5, -1.5, 0.5], [1.0, -2.5, 2.0, -0.5], [-0.5, 0.0, 0.5, 0.0], [0.0, 1.0, 0.0, 0.0]], np.float32), 1),
}


def basis_to_bezier(m, cp):
    """:1962-1977 — the Bernstein control points tracing the same cubic (f32, the reference's summation order)."""
    cp = np.asarray(cp, np.float32)

    def coeff(row):
        acc = np.zeros(3, np.float32)
        for c in range(4):
            acc = acc + cp[c] * m[row][c]
        return acc
    a, b, c, d = coeff(0), coeff(1), coeff(2), coeff(3)
    third = np.float32(3.0)
    return np.stack([d, d + c / third, d + c * np.float32(2.0 / 3.0) + b / third, a + b + c + d]).astype(np.float32)


def curve_segments(prim):
    """:1993-2131 — (round segments, cubic spans) of a BasisCurves prim, or None."""
    pts, counts = prim.attr("points"), prim.attr("curveVertexCounts")
    if pts is None or counts is None:
        return None
    pts = np.asarray(pts, np.float64).reshape(-1, 3).astype(np.float32)
    counts = [int(c) for c in np.asarray(counts).ravel()]
    widths = prim.attr("widths")
    widths = [1.0] if widths is None or len(widths) == 0 else [float(np.float32(w)) for w in np.asarray(widths, np.float64).ravel()]
    ty = prim.attr("type") if isinstance(prim.attr("type"), str) else "cubic"
    basis_name = prim.attr("basis") if isinstance(prim.attr("basis"), str) else "bezier"
    if basis_name not in _BASES:
        return None
    basis, vstep = _BASES[basis_name]
    n_points, n_curves = len(pts), len(counts)

    def width_of(global_idx, curve_idx):
        if len(widths) == n_points:
            return widths[global_idx]
        if len(widths) == n_curves:
            return widths[curve_idx]
        return widths[0]
    segments, cubics = [], []
    offset = 0
    for curve_idx, cnt in enumerate(counts):
        if cnt < 0 or offset + cnt > n_points:
            break
        cp = pts[offset:offset + cnt]
        radius = lambda k: float(np.float32(0.5) * np.float32(max(width_of(offset + k, curve_idx), 1e-6)))  # noqa: E731
        if ty == "linear":
            for k in range(max(cnt - 1, 0)):
                segments.append(CurveSegment(cp[k], cp[k + 1], radius(k), radius(k + 1)))
        elif cnt >= 4:
            for sp in range((cnt - 4) // vstep + 1):
                base = sp * vstep
                cubics.append(CubicCurveSegment(basis_to_bezier(basis, cp[base:base + 4]), radius(base), radius(base + 3)))
        offset += cnt
    if not segments and not cubics:
        return None
    return segments, cubics


# ---- f32 4x4 matrices in glam's column-vector convention ---------------------------------------
def _ident():
    return np.eye(4, dtype=F)


def _translate(v):
    m = _ident()
    m[:3, 3] = np.asarray(v, F)
    return m


def _scale(v):
    m = _ident()
    m[0, 0], m[1, 1], m[2, 2] = (F(x) for x in v)
    return m


def _rot(axis: int, deg: float):
    a = F(np.radians(F(deg), dtype=F))
    s, c = F(np.sin(a, dtype=F)), F(np.cos(a, dtype=F))
    m = _ident()
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    m[i, i], m[j, j], m[j, i], m[i, j] = c, c, s, -s
    return m


def _quat(w, x, y, z):
    q = np.asarray([x, y, z, w], F)
    q = q / F(np.sqrt((q * q).sum(dtype=F)))
    x, y, z, w = q
    m = _ident()
    m[:3, :3] = np.asarray([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], F)
    return m


def _mul(a, b):
    return (a @ b).astype(F)


def _point(m, p):
    return (m[:3, :3] @ np.asarray(p, F) + m[:3, 3]).astype(F)


def _vector(m, v):
    return (m[:3, :3] @ np.asarray(v, F)).astype(F)


def _affine(m) -> Affine3A:
    return Affine3A([*m[:3, 0], *m[:3, 1], *m[:3, 2], *m[:3, 3]])


def _normalize(v):
    v = np.asarray(v, F)
    return v * (F(1.0) / F(np.sqrt((v * v).sum(dtype=F))))


def local_matrix(prim: usda.Prim):
    """compose_xform_ops (:563-588); returns (matrix, resets_xform_stack)."""
    order = prim.attr("xformOpOrder")
    if order is None:
        return _ident(), False
    local = _ident()
    resets = False
    for token in order:
        if token == "!resetXformStack!":
            resets = True
            continue
        inverted = token.startswith("!invert!")
        name = token[len("!invert!"):] if inverted else token
        kind = name[len("xformOp:"):].split(":")[0]
        v = prim.attr(name)
        if v is None:
            continue
        if kind == "translate":
            m = _translate(v)
        elif kind == "scale":
            m = _scale(v)
        elif kind == "transform":
            m = np.asarray(v, F).reshape(4, 4).T.copy()  # USD row-vector matrix -> column-vector
        elif kind == "orient":
            m = _quat(*v)
        elif kind in ("rotateX", "rotateY", "rotateZ"):
            m = _rot("XYZ".index(kind[-1]), v)
        elif kind.startswith("rotate") and len(kind) == 9:
            rx, ry, rz = _rot(0, v[0]), _rot(1, v[1]), _rot(2, v[2])
            r = {"X": rx, "Y": ry, "Z": rz}
            a, b, c = kind[6:]
            m = _mul(_mul(r[c], r[b]), r[a])  # rotateXYZ = Rz * Ry * Rx (:618-627)
        else:
            continue
        if inverted:
            m = np.linalg.inv(m.astype(np.float64)).astype(F)
        local = _mul(local, m)
    return local, resets


# ---- materials ------------------------------------------------------------------------------------
_CAMEL = {  # decode_crust_openpbr (:2902-2975): input name -> record field
    "baseWeight": "base_weight", "baseColor": "base_color", "baseDiffuseRoughness": "base_diffuse_roughness", "baseMetalness": "base_metalness",
    "specularWeight": "specular_weight", "specularColor": "specular_color", "specularRoughness": "specular_roughness", "specularIor": "specular_ior",
    "specularRoughnessAnisotropy": "specular_roughness_anisotropy", "transmissionWeight": "transmission_weight", "transmissionColor": "transmission_color",
    "transmissionDepth": "transmission_depth", "transmissionScatter": "transmission_scatter", "transmissionScatterAnisotropy": "transmission_scatter_anisotropy",
    "transmissionDispersionScale": "transmission_dispersion_scale", "transmissionDispersionAbbeNumber": "transmission_dispersion_abbe_number",
    "subsurfaceWeight": "subsurface_weight", "subsurfaceColor": "subsurface_color", "subsurfaceRadius": "subsurface_radius",
    "subsurfaceRadiusScale": "subsurface_radius_scale", "subsurfaceScatterAnisotropy": "subsurface_scatter_anisotropy", "fuzzWeight": "fuzz_weight",
    "fuzzColor": "fuzz_color", "fuzzRoughness": "fuzz_roughness", "coatWeight": "coat_weight", "coatColor": "coat_color", "coatRoughness": "coat_roughness",
    "coatRoughnessAnisotropy": "coat_roughness_anisotropy", "coatIor": "coat_ior", "coatDarkening": "coat_darkening", "thinFilmWeight": "thin_film_weight",
    "thinFilmThickness": "thin_film_thickness", "thinFilmIor": "thin_film_ior", "emissionLuminance": "emission_luminance", "emissionColor": "emission_color",
    "geometryOpacity": "geometry_opacity", "geometryThinWalled": "geometry_thin_walled",
}


class _Materials:
    def __init__(self, stage: usda.Stage):
        self.stage = stage
        self.cache: Dict[str, Tuple[int, np.ndarray]] = {}
        self.default = (0, OpenPBR())
        self.next_id = 1

    def resolve(self, prim: usda.Prim) -> Tuple[int, np.ndarray]:
        """resolve_material (:2547-2564): (identity for the dedup key, record)."""
        targets = prim.relationships.get("material:binding")
        if not targets:
            return self.default
        path = targets[0]
        if path in self.cache:
            return self.cache[path]
        mat = self.stage.prim_at(path)
        rec = OpenPBR()
        if mat is not None:
            shader = None
            out = mat.attributes.get("outputs:surface")
            if out is not None and "connect" in out.metadata:
                shader = self.stage.prim_at(str(out.metadata["connect"]).rsplit(".", 1)[0])
            if shader is None:
                shader = next((c for c in mat.children if c.type == "Shader"), None)
            if shader is not None:
                sid = shader.attr("info:id")
                if sid == "crust:openpbr":
                    rec = OpenPBR(**{_CAMEL[k[7:]]: (int(bool(v)) if k[7:] == "geometryThinWalled" else v) for k, a in shader.attributes.items()
                                     if k.startswith("inputs:") and k[7:] in _CAMEL and (v := a.get()) is not None})
                elif sid == "UsdPreviewSurface":  # preview_surface_openpbr (:2661-2705)
                    kw = {}
                    g = lambda n: shader.attr("inputs:" + n)  # noqa: E731
                    for src, dst in (("diffuseColor", "base_color"), ("metallic", "base_metalness"), ("roughness", "specular_roughness"),
                                     ("opacity", "geometry_opacity"), ("ior", "specular_ior"), ("clearcoat", "coat_weight"), ("clearcoatRoughness", "coat_roughness")):
                        if g(src) is not None:
                            kw[dst] = g(src)
                    if g("emissiveColor") is not None:
                        kw["emission_color"] = g("emissiveColor")
                        if max(g("emissiveColor")) > 0.0:
                            kw["emission_luminance"] = 1.0
                    rec = OpenPBR(**kw)
        self.cache[path] = (self.next_id, rec)
        self.next_id += 1
        return self.cache[path]


def triangulate(counts, indices, n_verts: int) -> np.ndarray:
    """:1164-1214 — fans, skipping degenerate faces and out-of-range corners."""
    tris: List[Tuple[int, int, int]] = []
    off = 0
    for fc in counts:
        fc = int(fc)
        if fc >= 3 and off + fc <= len(indices):
            i0 = int(indices[off])
            for k in range(1, fc - 1):
                i1, i2 = int(indices[off + k]), int(indices[off + k + 1])
                if min(i0, i1, i2) >= 0 and max(i0, i1, i2) < n_verts:
                    tris.append((i0, i1, i2))
        off += fc
    return np.asarray(tris, np.uint32).reshape(-1, 3)


def _ray_mask(prim) -> int:  # :735-739
    m = prim.attr("crust:rayMask")
    return MASK_ALL if m is None else int(m) & 0xFFFFFFFF


def _light_mask(prim) -> int:  # :747-753
    m = prim.attr("crust:rayMask")
    if m is not None:
        return int(m) & 0xFFFFFFFF
    return MASK_SHADOW | MASK_INDIRECT | (MASK_CAMERA if prim.attr("crust:light:cameraVisible") else 0)


def _lux_emission(prim):  # :2248-2255
    intensity = prim.attr("inputs:intensity")
    exposure = prim.attr("inputs:exposure")
    color = prim.attr("inputs:color")
    gain = F(1.0 if intensity is None else intensity) * F(2.0) ** F(0.0 if exposure is None else exposure)
    return (np.asarray((1.0, 1.0, 1.0) if color is None else color, F) * gain).astype(F)


def import_render_settings(stage: usda.Stage) -> RenderSettings:  # :3015-3100
    p = None
    rsp = stage.metadata.get("renderSettingsPrimPath")
    if rsp:
        p = stage.prim_at(rsp)
    if p is None:
        p = stage.prim_at("/Render/settings")
    d = DEFAULTS
    if p is None or p.type != "RenderSettings":
        return RenderSettings(d["spp"], d["max_depth"], d["width"], d["height"], d["min_spp"], d["variance"], d["frame"])
    g = p.attr
    res = g("resolution")
    w, h = (int(res[0]), int(res[1])) if res is not None else (d["width"], d["height"])
    strategy = {None: SamplingStrategy.PowerMis, "power": SamplingStrategy.PowerMis, "mis": SamplingStrategy.PowerMis, "balance": SamplingStrategy.BalanceMis,
                "light": SamplingStrategy.LightOnly, "bsdf": SamplingStrategy.BsdfOnly}.get(g("crust:samplingStrategy"), SamplingStrategy.PowerMis)
    names = {"box": (PixelFilterKind.Box, 0.5), "triangle": (PixelFilterKind.Triangle, 1.0), "gaussian": (PixelFilterKind.Gaussian, 1.5),
             "blackman": (PixelFilterKind.BlackmanHarris, 1.5), "mitchell": (PixelFilterKind.Mitchell, 2.0)}
    kind, radius = names.get(g("crust:pixelFilter"), (PixelFilterKind.Triangle, 1.0))
    if g("crust:pixelFilterRadius") is not None:
        radius = max(float(g("crust:pixelFilterRadius")), 0.01)
    pick = lambda n, dv: dv if g(n) is None else g(n)  # noqa: E731
    return RenderSettings(int(pick("crust:samplesPerPixel", d["spp"])), int(pick("crust:maxDepth", d["max_depth"])), w, h,
                          int(pick("crust:minSamplesPerPixel", d["min_spp"])), float(pick("crust:varianceThreshold", d["variance"])),
                          int(pick("crust:frame", d["frame"])), strategy, PixelFilter(kind, radius))


def build_camera(prim: usda.Prim, world, settings: RenderSettings) -> Camera:  # :2174-2220
    lookfrom = _point(world, (0, 0, 0))
    forward = _normalize(_vector(world, (0, 0, -1)))
    up = _normalize(_vector(world, (0, 1, 0)))
    g = lambda n, dv: dv if prim.attr(n) is None else prim.attr(n)  # noqa: E731
    focal, hap = F(g("focalLength", 50.0)), F(g("horizontalAperture", 20.955))
    w, h = F(settings.width), F(settings.height)
    vap = F(prim.attr("verticalAperture")) if prim.attr("verticalAperture") is not None else hap * h / w
    f_stop, focus = F(g("fStop", 0.0)), F(g("focusDistance", 10.0))
    vfov = F(2.0) * F(np.degrees(np.arctan(vap / (F(2.0) * focal), dtype=F), dtype=F))
    aperture = focal / f_stop if f_stop > 0 else F(0.0)
    return Camera(lookfrom, (lookfrom + forward * focus).astype(F), up, float(vfov), float(w / h), float(aperture), float(focus))


class _Import:
    def __init__(self, api, stage: usda.Stage, stage_path: str = "."):
        self.api, self.stage, self.stage_path = api, stage, stage_path
        self.wb = api.WorldBuilder()
        self.lights = LightList()
        self.camera: Optional[Camera] = None
        self.settings = import_render_settings(stage)
        self.materials = _Materials(stage)
        self.slots: Dict[tuple, dict] = {}      # MeshArena (:800-910)
        self.pending: List[dict] = []
        self.volumes: List[VolumeRegion] = []

    # -- stage walk (:140-269) ------------------------------------------------------------------
    def traverse(self, roots: List[Tuple[usda.Prim, np.ndarray]]):
        stack = list(roots)
        while stack:
            prim, parent_world = stack.pop()
            if prim.is_abstract() or not prim.is_active():
                continue
            local, resets = self.local_matrix_at(prim)
            world = local if resets else _mul(parent_world, local)
            t = self.type_of(prim)
            if self.is_instance(prim):  # native instancing (:166-188): place the prototype, never walk the proxy subtree
                proto = self.prototype_of(prim)
                if proto is not None:
                    self.attach_proto_parts(self.prototype_parts(proto, 0), world)
                    continue
            if t == "PointInstancer":
                self.emit_point_instancer(prim, world)
                continue  # prototypes are drawn through the instancer, never on their own (:197-208)
            if isinstance(prim.attr("crust:volume:type"), str):  # checked before every geometry schema (:190-209)
                self.emit_volume(prim, world)
            elif t == "Mesh":
                self.emit_mesh(prim, world)
            elif t == "Sphere":
                self.emit_sphere(prim, world)
            elif t == "BasisCurves":
                self.emit_curves(prim, world)
            elif t == "Camera":
                if self.camera is None:
                    self.camera = build_camera(prim, world, self.settings)
            elif t == "SphereLight":
                self.emit_sphere_light(prim, world)
            elif t == "RectLight":
                self.emit_rect_light(prim, world)
            elif t == "DistantLight":  # emit_distant_light (:2361-2380)
                angle = prim.attr("inputs:angle")
                self.lights.add(DistantLight(_vector(world, (0, 0, -1)), _lux_emission(prim), 0.53 if angle is None else float(angle)))
            elif t == "DomeLight":
                self.emit_dome_light(prim, world)
            for c in prim.children:
                stack.append((c, world))

    def emit_mesh(self, prim, world):  # :912-994
        pts, counts, idx = prim.attr("points"), prim.attr("faceVertexCounts"), prim.attr("faceVertexIndices")
        if pts is None or counts is None or idx is None:
            return
        pts = np.asarray(pts, np.float64).reshape(-1, 3).astype(F)
        counts = np.asarray(counts).astype(np.int64).ravel()
        idx = np.asarray(idx).astype(np.int64).ravel()
        mat_id, mat = self.materials.resolve(prim)
        mask = _ray_mask(prim)
        motion = prim.attr("crust:motion:translate")
        if abs(float(np.linalg.det(world.astype(np.float64)))) < 1e-12:
            tris = triangulate(counts, idx, len(pts))
            if len(tris):
                verts = (pts @ world[:3, :3].T + world[:3, 3]).astype(F)
                self.wb.attach_masked(Geometry.TriangleMesh(verts, tris), mat, mask)
            return
        key = (pts.tobytes(), counts.tobytes(), idx.tobytes(), mat_id)  # MeshKey (:769-798): content + material identity
        slot = self.slots.get(key)
        if slot is None:
            tris = triangulate(counts, idx, len(pts))
            if not len(tris):
                return
            slot = self.slots[key] = dict(verts=pts, tris=tris, n_place=0, scene=None)
        slot["n_place"] += 1
        gid = self.wb.reserve_slot(mat, mask)
        self.pending.append(dict(geom_id=gid, slot=slot, l2w=world, motion=motion))

    def flush_meshes(self):  # :1025-1089
        for p in self.pending:
            slot, l2w = p["slot"], p["l2w"]
            mirrored = float(np.linalg.det(l2w[:3, :3].astype(np.float64))) < 0.0
            if slot["n_place"] == 1 and slot["scene"] is None and p["motion"] is None:
                verts = ((slot["verts"][:, 0:1] * l2w[:3, 0] + slot["verts"][:, 1:2] * l2w[:3, 1]) + slot["verts"][:, 2:3] * l2w[:3, 2] + l2w[:3, 3]).astype(F)
                tris = slot["tris"][:, [0, 2, 1]] if mirrored else slot["tris"]
                self.wb.set_geometry(p["geom_id"], Geometry.TriangleMesh(verts, np.ascontiguousarray(tris)))
            else:
                if slot["scene"] is None:
                    sb = self.api.SceneBuilder()
                    sb.attach(Geometry.TriangleMesh(slot["verts"], slot["tris"]))
                    slot["scene"] = sb.commit()
                end = None
                if p["motion"] is not None:
                    end = _affine(_mul(_translate(p["motion"]), l2w))
                self.wb.set_geometry(p["geom_id"], Geometry.Instance(slot["scene"], _affine(l2w), end))

    # -- composition arcs, as far as one layer goes -----------------------------------------------------
    def arc_target(self, prim):
        """The prim an internal reference / inherit / specialize arc of `prim` names, or None."""
        for key in ("references", "inherits", "specializes"):
            v = prim.metadata.get(key)
            if v is None:
                continue
            for item in (v if isinstance(v, (list, tuple)) else [v]):
                if isinstance(item, str) and item.strip().startswith("</"):
                    return self.stage.prim_at(item.strip().strip("<>"))
        return None

    def is_instance(self, prim) -> bool:  # Prim::is_instance: `instanceable = true` plus a composition arc
        return str(prim.metadata.get("instanceable", "false")).lower() == "true" and self.arc_target(prim) is not None

    def prototype_of(self, prim):
        return self.arc_target(prim)

    def type_of(self, prim) -> str:
        """A typeless `def`/`over` with an arc takes the referenced prim's type (composition)."""
        if prim.type:
            return prim.type
        target = self.arc_target(prim)
        return target.type if target is not None else prim.type

    def local_matrix_at(self, prim):
        """The instance's own xformOps are the stronger opinion; without any, the referenced prim's apply."""
        if prim.attr("xformOpOrder") is None:
            target = self.arc_target(prim)
            if target is not None:
                return local_matrix(target)
        return local_matrix(prim)

    # -- prototypes (:1379-1700) -----------------------------------------------------------------------
    def prototype_parts(self, root: usda.Prim, depth: int):  # :1808-1830, memoized by path
        cache = self.__dict__.setdefault("_protos", {})
        if root.path not in cache:
            cache[root.path] = self.collect_proto_parts(root, depth)
        return cache[root.path]

    def curve_scene(self, prim):
        segs = curve_segments(prim)
        if segs is None:
            return None
        sb = self.api.SceneBuilder()
        if segs[0]:
            sb.attach(Geometry.RoundCurves(segs[0]))
        if segs[1]:
            sb.attach(Geometry.CubicCurves(segs[1]))
        return sb.commit()

    def collect_proto_parts(self, root: usda.Prim, depth: int):
        """:1379-1540 — every leaf geometry under a prototype root as (BLAS, root-relative matrix, material, mask).
        Abstract prims are NOT skipped here: naming a class as a prototype is how instancing-only geometry is authored."""
        parts = []
        if depth > MAX_INSTANCE_NESTING:
            return parts
        stack = [(root, _ident())]
        while stack:
            prim, parent_local = stack.pop()
            if not prim.is_active():
                continue
            if prim is root:
                local = _ident()  # the root's own transform is excluded: the instance supplies the placement
            else:
                m, resets = self.local_matrix_at(prim)
                local = m if resets else _mul(parent_local, m)
            mask = _ray_mask(prim)
            if prim is not root and self.is_instance(prim):
                continue  # nested native instance: unreadable with openusd 0.5, skipped with a warning (:1418-1440)
            t = self.type_of(prim)
            if t == "Mesh":
                pts, counts, idx = prim.attr("points"), prim.attr("faceVertexCounts"), prim.attr("faceVertexIndices")
                if pts is not None and counts is not None and idx is not None:
                    pts = np.asarray(pts, np.float64).reshape(-1, 3).astype(F)
                    counts = np.asarray(counts).astype(np.int64).ravel()
                    idx = np.asarray(idx).astype(np.int64).ravel()
                    mat_id, mat = self.materials.resolve(prim)
                    skey = (pts.tobytes(), counts.tobytes(), idx.tobytes(), mat_id)
                    slot = self.slots.get(skey)
                    if slot is None:
                        tris = triangulate(counts, idx, len(pts))
                        slot = self.slots[skey] = dict(verts=pts, tris=tris, n_place=0, scene=None) if len(tris) else None
                    if slot is not None:
                        if slot["scene"] is None:  # committing also marks the slot ineligible for baking (:1446-1449)
                            sb = self.api.SceneBuilder()
                            sb.attach(Geometry.TriangleMesh(slot["verts"], slot["tris"]))
                            slot["scene"] = sb.commit()
                        parts.append((slot["scene"], local, mat, mask))
            elif t == "Sphere":
                r = prim.attr("radius")
                sb = self.api.SceneBuilder()
                sb.attach(Geometry.Sphere((0.0, 0.0, 0.0), 1.0 if r is None else float(F(r))))
                parts.append((sb.commit(), local, self.materials.resolve(prim)[1], mask))
            elif t == "BasisCurves":
                scene = self.curve_scene(prim)
                if scene is not None:
                    parts.append((scene, local, self.materials.resolve(prim)[1], mask))
            elif isinstance(prim.attr("crust:volume:type"), str):
                pass  # volumes cannot be instanced (:1497-1506)
            elif t == "PointInstancer":
                parts.extend(self.nested_instancer_parts(prim, local, mask, depth + 1))
                continue  # its prototypes are reached through it, never drawn directly
            for c in prim.children:
                stack.append((c, local))
        return parts

    def nested_instancer_parts(self, prim, local, mask, depth):
        """:1549-1614 — one part per (prototype, part) of the nested instancer: a sub-scene holding that part once per
        nested placement, so nesting nests instead of multiplying the outer instance count."""
        layout = self.read_instancer(prim)
        if layout is None:
            return []
        targets, placements = layout
        out = []
        for k, target in enumerate(targets):
            for scene, part_local, mat, part_mask in (self.prototype_parts(target, depth) if target is not None else []):
                sub = self.api.SceneBuilder()
                placed = 0
                for kk, xf in placements:
                    if kk != k:
                        continue
                    placement = _mul(xf, part_local)
                    if abs(float(np.linalg.det(placement.astype(np.float64)))) < 1e-12:
                        continue
                    sub.attach_masked(Geometry.Instance(scene, _affine(placement)), part_mask)
                    placed += 1
                if placed:
                    out.append((sub.commit(), local, mat, mask))
        return out

    def attach_proto_parts(self, parts, placement) -> int:  # :1637-1671
        attached = 0
        for scene, local, mat, mask in parts:
            xf = _mul(placement, local)
            if abs(float(np.linalg.det(xf.astype(np.float64)))) < 1e-12:
                continue
            self.wb.attach_masked(Geometry.Instance(scene, _affine(xf)), mat, mask)
            attached += 1
        return attached

    # -- PointInstancer (:1674-1875) ---------------------------------------------------------------
    def proto_parts(self, root: usda.Prim):
        return self.prototype_parts(root, 0)

    def read_instancer(self, prim):
        """:1702-1792 — (prototype root prims, [(target index, T * R * S)]) with `invisibleIds` already dropped;
        `orientationsf` wins over `orientations`."""
        targets = prim.relationships.get("prototypes") or []
        proto_idx = prim.attr("protoIndices")
        if not targets or proto_idx is None:
            return None
        proto_idx = np.asarray(proto_idx).astype(np.int64).ravel()
        positions = prim.attr("positions")
        positions = np.zeros((0, 3), F) if positions is None else np.asarray(positions, np.float64).reshape(-1, 3).astype(F)
        scales = prim.attr("scales")
        scales = None if scales is None else np.asarray(scales, np.float64).reshape(-1, 3).astype(F)
        orient = prim.attr("orientationsf")
        if orient is None or len(orient) == 0:
            orient = prim.attr("orientations")
        orient = None if orient is None or len(orient) == 0 else np.asarray(orient, np.float32).reshape(-1, 4)  # (w, x, y, z)
        ids = prim.attr("ids")
        invisible = set(int(v) for v in (prim.attr("invisibleIds") if prim.attr("invisibleIds") is not None else []))
        placements = []
        for i, k in enumerate(proto_idx):
            if i >= len(positions):
                break
            inst_id = int(ids[i]) if ids is not None and i < len(ids) else i
            if inst_id in invisible:
                continue
            if k < 0 or k >= len(targets):
                continue
            m = _translate(positions[i])
            if orient is not None and i < len(orient):
                m = _mul(m, _quat(*orient[i]))
            if scales is not None and i < len(scales):
                m = _mul(m, _scale(scales[i]))
            placements.append((int(k), m))
        return [self.stage.prim_at(t) for t in targets], placements

    def emit_point_instancer(self, prim, world):  # :1832-1875
        layout = self.read_instancer(prim)
        if layout is None:
            return
        targets, placements = layout
        protos = [self.prototype_parts(root, 0) if root is not None else [] for root in targets]
        for k, xf in placements:
            self.attach_proto_parts(protos[k], _mul(world, xf))

    def emit_curves(self, prim, world):  # :2133-2168
        scene = self.curve_scene(prim)
        if scene is None or abs(float(np.linalg.det(world.astype(np.float64)))) < 1e-12:
            return
        self.wb.attach_masked(Geometry.Instance(scene, _affine(world)), self.materials.resolve(prim)[1], _ray_mask(prim))

    def emit_volume(self, prim, world):  # :435-503
        ty = prim.attr("crust:volume:type")
        g = lambda n, dv: dv if prim.attr(n) is None else prim.attr(n)  # noqa: E731
        if ty == "homogeneous":
            field = DensityField.Homogeneous()
        elif ty == "smoke":
            field = DensityField.Noise(float(g("crust:volume:noiseScale", 4.0)), max(int(g("crust:volume:noiseOctaves", 4)), 1), float(g("crust:volume:noiseGain", 0.5)),
                                       float(g("crust:volume:noiseLacunarity", 2.0)), float(g("crust:volume:noiseThreshold", 0.3)), int(g("crust:volume:noiseSeed", 0)))
        elif ty == "grid":
            dims, data = prim.attr("crust:volume:gridDims"), prim.attr("crust:volume:gridData")
            if dims is None or data is None or len(dims) != 3:
                return
            nx, ny, nz = (max(int(d), 1) for d in dims)
            data = np.asarray(data, np.float64).astype(F).ravel()
            if nx * ny * nz != data.size:
                return
            field = DensityField.Grid(nx, ny, nz, data)
        else:
            return  # unknown type: skipped with a warning in the reference
        size = prim.attr("size")
        half = F(0.5) if size is None else F(size) * F(0.5)
        self.volumes.append(VolumeRegion(world, (half, half, half), g("crust:volume:sigmaS", (0.5, 0.5, 0.5)), g("crust:volume:sigmaA", (0.0, 0.0, 0.0)),
                                         float(g("crust:volume:anisotropy", 0.0)), g("crust:volume:emission", (0.0, 0.0, 0.0)),
                                         float(g("crust:volume:densityScale", 1.0)), field))

    def emit_dome_light(self, prim, world):  # :2389-2460
        tint = _lux_emission(prim)
        tex, fmt = prim.attr("inputs:texture:file"), prim.attr("inputs:texture:format")
        env = None
        if tex and fmt in (None, "latlong", "automatic"):
            pixels = self.load_environment(tex)
            if pixels is not None:
                # the light's orientation only: columns of the world matrix, normalised (:2443-2448)
                cols = [world[:3, k] for k in range(3)]
                cols = [c / F(np.sqrt((c * c).sum(dtype=F))) if float((c * c).sum()) > 0 else np.eye(3, dtype=F)[k] for k, c in enumerate(cols)]
                env = (pixels, np.stack(cols, axis=1))
        self.lights.add(DomeLight(tint), environment=env)

    def load_environment(self, asset: str):
        """AssetLoader::load_environment (scene.rs:94-135): the host decodes the image; here OpenCV's EXR/HDR reader,
        relative to the stage. Returns float32 [h, w, 3] RGB or None."""
        import os
        path = asset if os.path.isabs(asset) else os.path.join(os.path.dirname(os.path.abspath(self.stage_path)), asset)
        if not os.path.exists(path):
            return None
        if path.endswith(".npy"):  # pre-decoded float32 [h, w, 3] (tests/golden fixtures)
            return np.ascontiguousarray(np.load(path).astype(np.float32))
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        try:
            import cv2
            im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        except Exception:
            return None
        if im is None or im.ndim != 3:
            return None
        return np.ascontiguousarray(im[:, :, 2::-1].astype(np.float32))

    def emit_sphere(self, prim, world):  # :1235-1272
        r = prim.attr("radius")
        radius = 1.0 if r is None else float(F(r))
        center = _point(world, (0, 0, 0))
        _, mat = self.materials.resolve(prim)
        mask = _ray_mask(prim)
        motion = prim.attr("crust:motion:translate")
        if motion is not None:
            sb = self.api.SceneBuilder()
            sb.attach(Geometry.Sphere(center, radius))
            self.wb.attach_masked(Geometry.Instance(sb.commit(), Affine3A.IDENTITY, Affine3A.from_translation(motion)), mat, mask)
        else:
            self.wb.attach_masked(Geometry.Sphere(center, radius), mat, mask)

    def emit_sphere_light(self, prim, world):  # :2259-2297
        r = prim.attr("inputs:radius")
        radius = 0.5 if r is None else float(F(r))
        radiance = _lux_emission(prim)
        pos = _point(world, (0, 0, 0))
        gid = self.wb.attach_masked(Geometry.Sphere(pos, radius), Emissive(radiance), _light_mask(prim))
        self.lights.add(SphereLight(pos, radius, radiance, gid))

    def emit_rect_light(self, prim, world):  # :2299-2349
        w = F(1.0 if prim.attr("inputs:width") is None else prim.attr("inputs:width"))
        h = F(1.0 if prim.attr("inputs:height") is None else prim.attr("inputs:height"))
        radiance = _lux_emission(prim)
        origin = _point(world, (F(-0.5) * w, F(-0.5) * h, 0))
        eu, ev, nz = _vector(world, (w, 0, 0)), _vector(world, (0, h, 0)), _vector(world, (0, 0, -1))
        verts = np.stack([origin, origin + eu, origin + eu + ev, origin + ev]).astype(F)
        gid = self.wb.attach_masked(Geometry.TriangleMesh(verts, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)), Emissive(radiance), _light_mask(prim))
        self.lights.add(RectLight(origin, eu, ev, nz, radiance, gid))


def world_from_meshes(api, meshes, settings: RenderSettings):
    """Meshes given as (points f32[n,3], triangles u32[m,3], local-to-world 4x4) in import order, all on the
    default material: applies the importer's dedup + bake-vs-instance rule (:769-798, 1025-1089). Used to
    rebuild a stage from a binary fixture (tests/golden/cornellbox_scene.npz)."""
    imp = _Import.__new__(_Import)
    imp.api, imp.wb, imp.lights, imp.slots, imp.pending, imp.stage_path = api, api.WorldBuilder(), LightList(), {}, [], "."
    imp.settings = settings
    mat = OpenPBR()
    for pts, tris, l2w in meshes:
        pts = np.ascontiguousarray(pts, F)
        tris = np.ascontiguousarray(tris, np.uint32)
        key = (pts.tobytes(), tris.tobytes(), 0)
        slot = imp.slots.get(key)
        if slot is None:
            slot = imp.slots[key] = dict(verts=pts, tris=tris, n_place=0, scene=None)
        slot["n_place"] += 1
        imp.pending.append(dict(geom_id=imp.wb.reserve_slot(mat, MASK_ALL), slot=slot, l2w=np.asarray(l2w, F), motion=None))
    imp.flush_meshes()
    return imp.wb.commit(imp.lights)


def load_scene(api, path: str, **overrides):
    """`Scene::from_usd` (scene.rs:67-85 -> usd_import.rs:287-424). Returns (world, camera, settings);
    keyword overrides replace RenderSettings fields (the CLI's -s / --strategy / --filter flags)."""
    from .compose import load_composed
    stage = load_composed(path)  # the layer stack with variant selections, external references and payloads flattened in
    return import_stage(api, stage, path, **overrides)


def import_stage(api, stage: usda.Stage, path: str = "<stage>", **overrides):
    """The import proper, from an already parsed (or synthesised) Stage."""
    imp = _Import(api, stage, path)
    for k, v in overrides.items():
        setattr(imp.settings, k, v)
    # stream_roots (:67-115): import subtree by subtree when there are at least 4 of them
    tops = stage.roots
    chunks = tops[0].children if len(tops) == 1 else tops
    ident = _ident()
    if len(chunks) >= MIN_STREAM_CHUNKS:
        if len(tops) == 1:
            top = tops[0]
            local, resets = local_matrix(top)
            for c in chunks:
                if top.is_abstract() or not top.is_active():
                    break
                imp.traverse([(c, local)])
        else:
            for c in chunks:
                imp.traverse([(c, ident)])
    else:
        imp.traverse([(t, ident) for t in tops])
    imp.flush_meshes()
    world = imp.wb.commit(imp.lights)
    world.volumes = imp.volumes
    if imp.volumes:
        world.with_volumes(imp.volumes)  # Scene::with_volumes (:421)
    if imp.camera is None:
        raise ValueError(f"{path}: no UsdGeomCamera on the stage")
    return world, imp.camera, imp.settings
