"""Render scenes for BASELINE.json's configs, built through the WorldBuilder API.

The bundled samples are transcribed the way crust-core's USD importer would hand them to the
engine (crates/crust-core/src/scene/usd_import.rs): geometry in the importer's visiting order so `geom_id`s and the light order match
(the stage walk is a LIFO stack: siblings in reverse authored order, :140-269), quads fan-triangulated
(:1164-1214), lights as geometry + AreaLight sharing a geom_id with mask SHADOW|INDIRECT[|CAMERA]
(:747-753, 2259-2349), cameras from focal length / aperture (:2174-2220), unbound prims on the
default grey OpenPBR. The text-USD importer (usd/importer.py) reads the real files where /root/reference exists;
tests/test_usd_import.py proves these transcriptions equal what it produces, and the GPU box — which
has no reference tree — uses them plus the binary fixture behind `cornellbox()`.

Every builder takes an ``api`` namespace exposing ``WorldBuilder`` and ``Geometry`` (the CUDA
binding, or the oracle binding in tests) and returns (world, camera, settings).
"""
from __future__ import annotations

import math
import os
from typing import Tuple

import numpy as np

from .render_types import (
    Camera,
    DomeLight,
    Emissive,
    LightList,
    OpenPBR,
    PixelFilter,
    RectLight,
    RenderSettings,
    SamplingStrategy,
    SphereLight,
)
from .rt_types import MASK_CAMERA, MASK_INDIRECT, MASK_SHADOW, Affine3A, Geometry
from . import workloads as W

QUAD = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)  # fan triangulation of one quad


def _f32(x):
    return np.float32(x)


def usd_camera(translate, rotate_x_deg: float, focal_length: float, horizontal_aperture: float, width: int, height: int,
               vertical_aperture=None, f_stop: float = 0.0, focus_distance: float = 10.0, matrix=None) -> Camera:
    """`build_camera` (usd_import.rs:2174-2220) for a camera authored as translate + rotateX (or a full
    row-vector `xformOp:transform` matrix), through the importer's own f32 arithmetic."""
    from .usd.importer import _mul, _rot, _translate, build_camera
    from .usd.usda import Attribute, Prim

    world = np.asarray(matrix, np.float32).reshape(4, 4).T.copy() if matrix is not None else _mul(_translate(translate), _rot(0, rotate_x_deg))
    prim = Prim("def", "Camera", "Cam", "/World/Cam")
    vals = {"focalLength": focal_length, "horizontalAperture": horizontal_aperture, "fStop": f_stop, "focusDistance": focus_distance}
    if vertical_aperture is not None:
        vals["verticalAperture"] = vertical_aperture
    for k, v in vals.items():
        prim.attributes[k] = Attribute(k, "float", v)
    return build_camera(prim, world, RenderSettings(1, 1, width, height))


def _light_mask(camera_visible: bool) -> int:  # usd_import.rs:747-753
    return MASK_SHADOW | MASK_INDIRECT | (MASK_CAMERA if camera_visible else 0)


def _sphere_light(wb, lights: LightList, center, radius: float, color, intensity: float, camera_visible: bool):
    """emit_sphere_light (usd_import.rs:2259-2297)."""
    radiance = (np.asarray(color, np.float32) * _f32(intensity)).astype(np.float32)
    gid = wb.attach_masked(Geometry.Sphere(center, radius), Emissive(radiance), _light_mask(camera_visible))
    lights.add(SphereLight(center, radius, radiance, gid))


def _rect_light(wb, lights: LightList, origin, edge_u, edge_v, normal, radiance, camera_visible: bool):
    """emit_rect_light (usd_import.rs:2299-2349): two triangles spanning the rectangle."""
    o, eu, ev = (np.asarray(a, np.float32) for a in (origin, edge_u, edge_v))
    verts = np.stack([o, o + eu, o + eu + ev, o + ev]).astype(np.float32)
    gid = wb.attach_masked(Geometry.TriangleMesh(verts, QUAD), Emissive(radiance), _light_mask(camera_visible))
    lights.add(RectLight(o, eu, ev, normal, radiance, gid))


# ---------------------------------------------------------------------------------------------
# a small scene touching every integrator branch (tests / smoke)
# ---------------------------------------------------------------------------------------------
def furnace_room(api, width: int = 96, height: int = 64, spp: int = 8, max_depth: int = 8, **settings_kw):
    """Floor + back wall, a diffuse, a metal, a coated, a glass and a thin-walled sphere, a triangle-mesh
    sphere inside an instance, one sphere light and one rect light. Not a reference sample."""
    wb = api.WorldBuilder()
    lights = LightList()
    floor = np.array([(-8, 0, -8), (8, 0, -8), (8, 0, 8), (-8, 0, 8)], np.float32)
    wall = np.array([(-8, 0, -4), (8, 0, -4), (8, 8, -4), (-8, 8, -4)], np.float32)
    wb.attach(Geometry.TriangleMesh(floor, np.array([[0, 2, 1], [0, 3, 2]], np.uint32)), OpenPBR(base_color=(0.6, 0.6, 0.6)))
    wb.attach(Geometry.TriangleMesh(wall, QUAD), OpenPBR(base_color=(0.7, 0.5, 0.3), specular_weight=0.0))
    wb.attach(Geometry.Sphere((-3.0, 1.0, 0.0), 1.0), OpenPBR(base_color=(0.8, 0.1, 0.1)))
    wb.attach(Geometry.Sphere((-1.0, 1.0, 0.5), 1.0), OpenPBR(base_color=(0.9, 0.9, 0.9), base_metalness=1.0, specular_roughness=0.1))
    wb.attach(Geometry.Sphere((1.0, 1.0, 0.0), 1.0), OpenPBR(base_color=(0.1, 0.3, 0.8), coat_weight=1.0, coat_roughness=0.05))
    wb.attach(Geometry.Sphere((3.0, 1.0, 0.5), 1.0),
              OpenPBR(specular_roughness=0.05, specular_ior=1.5, transmission_weight=1.0, transmission_color=(0.8, 0.95, 0.8),
                      transmission_depth=1.0))
    wb.attach(Geometry.Sphere((0.0, 0.6, 2.5), 0.6),
              OpenPBR(specular_roughness=0.02, specular_ior=1.33, transmission_weight=0.5, geometry_thin_walled=1, thin_film_weight=1.0,
                      thin_film_thickness=0.55))
    sv, si = W.uv_sphere((0.0, 0.0, 0.0), 0.5, 16, 8)
    blas = api.SceneBuilder()
    blas.attach(Geometry.TriangleMesh(sv, si))
    blas_scene = blas.commit()
    for k, x in enumerate((-2.0, 2.0)):
        xf = Affine3A.from_translation((x, 0.5, 3.0)) * Affine3A.from_scale((1.0, 1.0 + 0.5 * k, 1.0))
        wb.attach(Geometry.Instance(blas_scene, xf), OpenPBR(base_color=(0.2, 0.7, 0.3), fuzz_weight=0.5))
    _sphere_light(wb, lights, (2.0, 5.0, 2.0), 0.5, (1.0, 0.9, 0.8), 30.0, True)
    _rect_light(wb, lights, (-4.0, 6.0, -1.0), (2.0, 0.0, 0.0), (0.0, 0.0, 2.0), (0.0, -1.0, 0.0), (10.0, 10.0, 12.0), False)
    world = wb.commit(lights)
    cam = Camera((0.0, 3.0, 9.0), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0), 40.0, width / height, 0.05, 9.0)
    st = RenderSettings(spp, max_depth, width, height, **settings_kw)
    return world, cam, st


# ---------------------------------------------------------------------------------------------
# config 2: samples/veach_mis.usda
# ---------------------------------------------------------------------------------------------
def veach_mis(api, width: int = 960, height: int = 540, spp: int = 128, **settings_kw):
    wb = api.WorldBuilder()
    lights = LightList()
    # The importer's stage walk pops siblings off a stack, i.e. visits them in REVERSE authored order
    # (usd_import.rs:140-269): BackWall, Floor, Plate3..Plate0, then LightBig..LightTiny.
    plates = [
        ([(-4.5, 1.845, -0.896), (4.5, 1.845, -0.896), (4.5, 2.555, -2.104), (-4.5, 2.555, -2.104)], 0.01),
        ([(-4.5, 1.161, 1.158), (4.5, 1.161, 1.158), (4.5, 1.639, -0.158), (-4.5, 1.639, -0.158)], 0.05),
        ([(-4.5, 0.450, 3.184), (4.5, 0.450, 3.184), (4.5, 0.750, 1.816), (-4.5, 0.750, 1.816)], 0.12),
        ([(-4.5, -0.263, 5.197), (4.5, -0.263, 5.197), (4.5, -0.137, 3.803), (-4.5, -0.137, 3.803)], 0.25),
    ]
    backdrop = OpenPBR(base_color=(0.05, 0.05, 0.05), specular_weight=0.0)
    wb.attach(Geometry.TriangleMesh(np.asarray([(-12, -2.5, -10), (12, -2.5, -10), (12, 14, -10), (-12, 14, -10)], np.float32), QUAD), backdrop)
    wb.attach(Geometry.TriangleMesh(np.asarray([(-12, -2.5, 14), (12, -2.5, 14), (12, -2.5, -10), (-12, -2.5, -10)], np.float32), QUAD), backdrop)
    for pts, rough in reversed(plates):
        wb.attach(Geometry.TriangleMesh(np.asarray(pts, np.float32), QUAD),
                  OpenPBR(base_color=(0.9, 0.9, 0.9), base_metalness=1.0, specular_roughness=rough))
    for pos, radius, color, intensity in (((3.375, 6, -3), 1.35, (0.4, 0.6, 1.0), 12.0), ((1.125, 6, -3), 0.45, (0.45, 1.0, 0.55), 108.0),
                                           ((-1.125, 6, -3), 0.15, (1.0, 0.85, 0.45), 972.0), ((-3.375, 6, -3), 0.05, (1.0, 0.45, 0.35), 8748.0)):
        _sphere_light(wb, lights, pos, radius, color, intensity, True)
    world = wb.commit(lights)
    cam = usd_camera((0, 4, 12), -10.0, 16.0, 20.955, width, height)
    kw = dict(min_samples_per_pixel=32, variance_threshold=0.0, sampling_strategy=SamplingStrategy.BalanceMis)
    kw.update(settings_kw)
    return world, cam, RenderSettings(spp, 8, width, height, **kw)


# ---------------------------------------------------------------------------------------------
# config 3: samples/openpbr_showcase.usda
# ---------------------------------------------------------------------------------------------
def openpbr_showcase(api, width: int = 640, height: int = 360, spp: int = 128, **settings_kw):
    wb = api.WorldBuilder()
    lights = LightList()
    # reverse sibling order (see veach_mis): Light2, Light1, then Scene's spheres soap bubble .. chrome, ground
    _sphere_light(wb, lights, (4, 8, -2), 1.5, (18, 12, 9), 1.0, True)
    _sphere_light(wb, lights, (-2, 8, 3), 1.5, (15, 15, 15), 1.0, True)
    mats = [
        OpenPBR(base_color=(0.95, 0.95, 0.95), base_metalness=1.0, specular_roughness=0.02),
        OpenPBR(base_color=(0.8, 0.1, 0.1)),
        OpenPBR(base_color=(0.8, 0.05, 0.05), base_metalness=0.9, specular_roughness=0.25, coat_weight=1.0, coat_roughness=0.02, coat_ior=1.5),
        OpenPBR(base_color=(0.35, 0.15, 0.55), base_diffuse_roughness=0.9, specular_roughness=0.9, fuzz_weight=1.0, fuzz_roughness=0.3),
        OpenPBR(specular_roughness=0.02, specular_ior=1.55, transmission_weight=1.0, transmission_color=(0.7, 0.9, 0.75), transmission_depth=1.0,
                transmission_dispersion_scale=0.6),
        OpenPBR(base_color=(0.9, 0.75, 0.55), specular_roughness=0.5, subsurface_weight=1.0, subsurface_color=(0.95, 0.55, 0.35), subsurface_radius=1.5),
        OpenPBR(base_color=(0.8, 0.8, 0.8), specular_roughness=0.02, specular_ior=1.33, transmission_weight=0.4, transmission_color=(0.95, 0.95, 1.0),
                thin_film_weight=1.0, thin_film_thickness=0.55, geometry_thin_walled=1),
    ]
    for k in reversed(range(len(mats))):
        wb.attach(Geometry.Sphere((-7.5 + 2.5 * k, 1, 0), 1.0), mats[k])
    wb.attach(Geometry.Sphere((0, -1000, 0), 1000.0), OpenPBR())  # ground, unbound
    world = wb.commit(lights)
    cam = usd_camera(None, 0.0, 50.0, 73.63848, width, height, vertical_aperture=41.42165, f_stop=1000.0, focus_distance=22.0,
                     matrix=((1, 0, 0, 0), (0, 0.9958932, -0.09053575, 0), (0, 0.09053575, 0.9958932, 0), (0, 3, 22, 1)))
    kw = dict(min_samples_per_pixel=32, variance_threshold=0.05)
    kw.update(settings_kw)
    return world, cam, RenderSettings(spp, 32, width, height, **kw)


# ---------------------------------------------------------------------------------------------
# config 5: scripts/gen_stress_scene.py (grid=6 as shipped; grid=32 + jitter = the 10 M-triangle K1)
# ---------------------------------------------------------------------------------------------
def stress(api, grid: int = 6, shards: int = 30000, jitter: bool = False, width: int = 640, height: int = 360, spp: int = 16, **settings_kw):
    wb = api.WorldBuilder()
    lights = LightList()
    grey = OpenPBR(base_color=(0.55, 0.55, 0.6), specular_roughness=0.4)  # UsdPreviewSurface mapping, README.md:102-106
    meshes = W.stress_meshes(grid, shards, jitter)
    # reverse sibling order (see veach_mis): the RectLight "Key" first, then OccluderSlab, Floor, Shards, Ball_{G-1}_{G-1} .. Ball_0_0.
    # Key: 12 x 12 at (0, 12, 0), rotateX -90 => emits along -Y
    _rect_light(wb, lights, (-6.0, 12.0, 6.0), (12.0, 0.0, 0.0), (0.0, 0.0, -12.0), (0.0, -1.0, 0.0), (6.0, 6.0, 6.0), False)
    n_tris = 0
    for k in reversed(range(len(meshes))):
        v, i = meshes[k]
        wb.attach(Geometry.TriangleMesh(v, i), grey if k < grid * grid else OpenPBR())
        n_tris += len(i)
    world = wb.commit(lights)
    cam = usd_camera((0, 9, 20), -22.0, 24.0, 20.955, width, height)
    kw = dict(min_samples_per_pixel=16, variance_threshold=0.0)
    kw.update(settings_kw)
    return world, cam, RenderSettings(spp, 6, width, height, **kw), n_tris + 2


# ---------------------------------------------------------------------------------------------
# config 1: samples/cornellbox.usda (from the binary fixture the importer wrote)
# ---------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cornellbox(api, width: int = 512, height: int = 512, spp: int = 64, **settings_kw):
    """5 meshes / 2 342 triangles, no lights (sky gradient only), importer defaults (depth 32, min 32,
    threshold 0.05, triangle filter) with BASELINE.json's 512x512 / 64 spp override."""
    from .usd.importer import DEFAULTS, build_camera, world_from_meshes
    from .usd.usda import Attribute, Prim

    z = np.load(os.path.join(GOLDEN, "cornellbox_scene.npz"))
    kw = dict(min_samples_per_pixel=DEFAULTS["min_spp"], variance_threshold=DEFAULTS["variance"])
    kw.update(settings_kw)
    st = RenderSettings(spp, DEFAULTS["max_depth"], width, height, **kw)
    meshes = [(z[f"mesh{k}_points"], z[f"mesh{k}_tris"].astype(np.uint32), z[f"mesh{k}_l2w"]) for k in range(int(z["n_meshes"]))]
    world = world_from_meshes(api, meshes, st)
    fl, hap, vap, fstop, focus = (float(v) for v in z["camera_params"])
    cam_prim = Prim("def", "Camera", "camera1", "/scene/camera1")
    for name, val in (("focalLength", fl), ("horizontalAperture", hap), ("verticalAperture", vap), ("fStop", fstop), ("focusDistance", focus)):
        cam_prim.attributes[name] = Attribute(name, "float", val)
    return world, build_camera(cam_prim, z["camera_l2w"], st), st


# ---------------------------------------------------------------------------------------------
# config 4: samples/PointInstancedMedCity.usd (from the binary fixture the crate reader wrote)
# ---------------------------------------------------------------------------------------------
def medcity(api, width: int = 640, height: int = 360, spp: int = 128, n_instances=None, **settings_kw):
    """One PointInstancer, 40 000 instances (T * R * S) over 8 prototype meshes (9 832 unique triangles), Z-up,
    default grey OpenPBR, no lights (sky only), importer default settings. The fixture is replayed through the
    importer itself as a synthesised stage, so placement arithmetic and ids are the importer's."""
    from .usd.importer import import_stage
    from .usd.usda import Attribute, Prim, Stage

    z = np.load(os.path.join(GOLDEN, "medcity_scene.npz"))

    def prim(ptype, name, parent, **attrs):
        p = Prim("def", ptype, name, (parent.path if parent else "") + "/" + name, parent=parent)
        for k, v in attrs.items():
            p.attributes[k] = Attribute(k, "", v)
        if parent is not None:
            parent.children.append(p)
        return p

    def xf(m):  # column-vector f32 matrix -> a row-vector `xformOp:transform`
        return {"xformOp:transform": np.asarray(m, np.float64).T.copy(), "xformOpOrder": ["xformOp:transform"]}

    cams = prim("Xform", "Cameras", None)
    fl, hap, vap, fstop, focus = (float(v) for v in z["camera_params"])
    cam_attrs = dict(focalLength=fl, horizontalAperture=hap, fStop=fstop, focusDistance=focus, **xf(z["camera_l2w"]))
    if vap > 0:
        cam_attrs["verticalAperture"] = vap
    prim("Camera", "main_cam", cams, **cam_attrs)
    hills = prim("Xform", "MediterraneanHills", None)
    n = len(z["positions"]) if n_instances is None else min(int(n_instances), len(z["positions"]))
    inst = prim("PointInstancer", "Buildings", hills, positions=z["positions"][:n], orientations=z["orientations_wxyz"][:n], scales=z["scales"][:n],
                protoIndices=z["proto_indices"][:n].astype(np.int32), **xf(z["instancer_l2w"]))
    protos = prim("Xform", "Prototypes", inst)
    targets = []
    for k in range(int(z["n_protos"])):
        root = prim("Xform", f"prototype_{k}", protos)
        tris = z[f"proto{k}_tris"].astype(np.int64)
        prim("Mesh", "mesh_0", root, points=z[f"proto{k}_points"], faceVertexCounts=np.full(len(tris), 3, np.int64), faceVertexIndices=tris.reshape(-1),
             **xf(z[f"proto{k}_local"]))
        targets.append(root.path)
    inst.relationships["prototypes"] = targets
    kw = dict(width=width, height=height, samples_per_pixel=spp)
    kw.update(settings_kw)
    return import_stage(api, Stage({"upAxis": "Z"}, [cams, hills]), "tests/golden/medcity_scene.npz", **kw)


# ---------------------------------------------------------------------------------------------
# samples/domelight.usda: lights at infinity
# ---------------------------------------------------------------------------------------------
def domelight(api, width: int = 480, height: int = 270, spp: int = 96, **settings_kw):
    """A chrome and a clay sphere on a ground plane under a DomeLight carrying the lat-long HDRI sky_env.exr (rotated
    20 degrees about Y, with a small very bright sun disc that importance sampling has to find) and a 1.5-degree
    DistantLight; depth 8, min 32 spp, threshold 0.02. Replayed through the importer as a synthesised stage whose
    texture asset is the decoded fixture tests/golden/sky_env_128x64.npy."""
    from .usd.importer import import_stage
    from .usd.usda import Attribute, Prim, Stage

    def prim(ptype, name, parent, **attrs):
        p = Prim("def", ptype, name, (parent.path if parent else "") + "/" + name, parent=parent)
        for k, v in attrs.items():
            p.attributes[k.replace("__", ":")] = Attribute(k.replace("__", ":"), "", v)
        if parent is not None:
            parent.children.append(p)
        return p

    world = prim("Xform", "World", None)
    prim("Camera", "Cam", world, focalLength=28.0, horizontalAperture=20.955, xformOp__translate=(0.0, 2.3, 8.5), xformOp__rotateX=-8.0,
         xformOpOrder=["xformOp:translate", "xformOp:rotateX"])
    prim("DomeLight", "Sky", world, inputs__intensity=1.0, inputs__color=(1.0, 1.0, 1.0), inputs__texture__file="sky_env_128x64.npy",
         inputs__texture__format="latlong", xformOp__rotateY=20.0, xformOpOrder=["xformOp:rotateY"])
    prim("DistantLight", "Sun", world, inputs__intensity=2.2, inputs__angle=1.5, inputs__color=(1.0, 0.94, 0.82), xformOp__rotateY=40.0, xformOp__rotateX=-48.0,
         xformOpOrder=["xformOp:rotateY", "xformOp:rotateX"])
    looks = Prim("def", "Scope", "Looks", "/World/Looks", parent=world)

    def material(name, **inputs):
        m = prim("Material", name, looks)
        sh = prim("Shader", "Shader", m, **{"info__id": "crust:openpbr", **{"inputs__" + k: v for k, v in inputs.items()}})
        m.attributes["outputs:surface"] = Attribute("outputs:surface", "token", None, metadata={"connect": sh.path + ".outputs:surface"})
        return m.path

    metal = material("Metal", baseColor=(0.92, 0.9, 0.86), baseMetalness=1.0, specularRoughness=0.08)
    clay = material("Clay", baseColor=(0.72, 0.36, 0.28), specularRoughness=0.65)
    ground = material("Ground", baseColor=(0.48, 0.47, 0.45), specularRoughness=0.8)
    for name, x, mat in (("Chrome", -1.6, metal), ("Matte", 1.6, clay)):
        s = prim("Sphere", name, world, radius=1.0, xformOp__translate=(x, 1.0, 0.0), xformOpOrder=["xformOp:translate"])
        s.relationships["material:binding"] = [mat]
    floor = prim("Mesh", "Floor", world, faceVertexCounts=[4], faceVertexIndices=[0, 1, 2, 3], points=[(-40, 0, -40), (40, 0, -40), (40, 0, 40), (-40, 0, 40)])
    floor.relationships["material:binding"] = [ground]
    world.children.append(looks)
    render = prim("Scope", "Render", None)
    prim("RenderSettings", "settings", render, resolution=(480, 270), **{"crust__samplesPerPixel": 96, "crust__maxDepth": 8, "crust__minSamplesPerPixel": 32,
                                                                          "crust__varianceThreshold": 0.02, "crust__frame": 0})
    kw = dict(width=width, height=height, samples_per_pixel=spp)
    kw.update(settings_kw)
    return import_stage(api, Stage({"upAxis": "Y"}, [world, render]), os.path.join(GOLDEN, "domelight.usda"), **kw)


# ---------------------------------------------------------------------------------------------
# volume regions: the reference's samples/fog.usda and samples/smoke.usda, replayed through the
# importer as synthesised stages (tests/test_usd_import.py proves them equal to the importer's output on the files)
# ---------------------------------------------------------------------------------------------
def _stage_prim(ptype, name, parent, **attrs):
    from .usd.usda import Attribute, Prim

    p = Prim("def", ptype, name, (parent.path if parent else "") + "/" + name, parent=parent)
    for k, v in attrs.items():
        p.attributes[k.replace("__", ":")] = Attribute(k.replace("__", ":"), "", v)
    if parent is not None:
        parent.children.append(p)
    return p


def _grey_room_stage(volume_prims, extra=None):
    """The five-sided grey room under a 1.6 x 1.6 ceiling RectLight both samples share (fog.usda:11-27,49-60)."""
    from .usd.usda import Stage

    world = _stage_prim("Xform", "World", None)
    _stage_prim("Camera", "Cam", world, focalLength=24.0, horizontalAperture=20.955, xformOp__translate=(0.0, 2.0, 8.5), xformOpOrder=["xformOp:translate"])
    _stage_prim("Mesh", "Room", world, faceVertexCounts=[4] * 5, faceVertexIndices=list(range(20)),
                points=[(-2, 0, -2), (2, 0, -2), (2, 0, 2), (-2, 0, 2), (-2, 4, -2), (2, 4, -2), (2, 4, 2), (-2, 4, 2), (-2, 0, -2), (2, 0, -2), (2, 4, -2),
                        (-2, 4, -2), (-2, 0, -2), (-2, 0, 2), (-2, 4, 2), (-2, 4, -2), (2, 0, -2), (2, 0, 2), (2, 4, 2), (2, 4, -2)])
    if extra:
        extra(world)
    for name, attrs in volume_prims:
        _stage_prim("Cube", name, world, **attrs)
    _stage_prim("RectLight", "CeilingLight", world, inputs__width=1.6, inputs__height=1.6, inputs__color=(10.0, 10.0, 10.0), inputs__intensity=1.0,
                xformOp__rotateX=-90.0, xformOp__translate=(0.0, 3.98, 0.0), xformOpOrder=["xformOp:translate", "xformOp:rotateX"])
    render = _stage_prim("Scope", "Render", None)
    _stage_prim("RenderSettings", "settings", render, resolution=(320, 180), **{"crust__samplesPerPixel": 64, "crust__maxDepth": 16, "crust__minSamplesPerPixel": 16,
                                                                                "crust__varianceThreshold": 0.05, "crust__frame": 0})
    return Stage({"upAxis": "Y"}, [world, render])


def fog(api, **settings_kw):
    """samples/fog.usda: homogeneous fog filling a grey room, god rays around an occluding ball (volume NEE + the
    analytic transmittance fast path). 320x180, 64 spp, depth 16, adaptive (min 16, threshold 0.05)."""
    from .usd.importer import import_stage

    ball = lambda w: _stage_prim("Sphere", "Ball", w, radius=0.6, xformOp__translate=(-0.7, 0.6, 0.3), xformOpOrder=["xformOp:translate"])  # noqa: E731
    vols = [("Fog", dict(size=4.0, crust__volume__type="homogeneous", crust__volume__sigmaS=(0.15, 0.15, 0.15), crust__volume__sigmaA=(0.01, 0.015, 0.025),
                         crust__volume__anisotropy=0.3, xformOp__translate=(0.0, 2.0, 0.0), xformOpOrder=["xformOp:translate"]))]
    return import_stage(api, _grey_room_stage(vols, ball), "fog.usda", **settings_kw)


# smoke.usda:66 — a 4x4x4 spherical blob
_PUFF = [0.000, 0.092, 0.092, 0.000, 0.092, 0.309, 0.309, 0.092, 0.092, 0.309, 0.309, 0.092, 0.000, 0.092, 0.092, 0.000,
         0.092, 0.309, 0.309, 0.092, 0.309, 0.639, 0.639, 0.309, 0.309, 0.639, 0.639, 0.309, 0.092, 0.309, 0.309, 0.092,
         0.092, 0.309, 0.309, 0.092, 0.309, 0.639, 0.639, 0.309, 0.309, 0.639, 0.639, 0.309, 0.092, 0.309, 0.309, 0.092,
         0.000, 0.092, 0.092, 0.000, 0.092, 0.309, 0.309, 0.092, 0.092, 0.309, 0.309, 0.092, 0.000, 0.092, 0.092, 0.000]


def smoke(api, **settings_kw):
    """samples/smoke.usda: a procedural fBm smoke plume, an emissive absorbing ember and a 4x4x4 density-grid puff in
    the grey room (delta tracking, volume emission, ratio-tracked shadow rays, overlapping majorants)."""
    from .usd.importer import import_stage

    vols = [
        ("Smoke", dict(size=2.0, crust__volume__type="smoke", crust__volume__densityScale=12.0, crust__volume__sigmaS=(0.8, 0.8, 0.8),
                       crust__volume__sigmaA=(0.08, 0.08, 0.08), crust__volume__anisotropy=0.2, crust__volume__noiseScale=4.0, crust__volume__noiseOctaves=4,
                       crust__volume__noiseGain=0.5, crust__volume__noiseLacunarity=2.0, crust__volume__noiseThreshold=0.25, crust__volume__noiseSeed=42,
                       xformOp__scale=(0.9, 1.6, 0.9), xformOp__translate=(-0.4, 1.7, 0.0), xformOpOrder=["xformOp:translate", "xformOp:scale"])),
        ("Ember", dict(size=0.5, crust__volume__type="homogeneous", crust__volume__densityScale=4.0, crust__volume__sigmaS=(0.05, 0.05, 0.05),
                       crust__volume__sigmaA=(0.6, 0.35, 0.15), crust__volume__emission=(6.0, 2.2, 0.5), xformOp__translate=(0.9, 0.3, 0.4),
                       xformOpOrder=["xformOp:translate"])),
        ("GridPuff", dict(size=1.0, crust__volume__type="grid", crust__volume__densityScale=6.0, crust__volume__sigmaS=(0.7, 0.7, 0.7),
                          crust__volume__sigmaA=(0.05, 0.05, 0.05), crust__volume__gridDims=[4, 4, 4], crust__volume__gridData=list(_PUFF),
                          xformOp__translate=(1.1, 2.6, -0.8), xformOpOrder=["xformOp:translate"])),
    ]
    return import_stage(api, _grey_room_stage(vols), "smoke.usda", **settings_kw)


def sample_stage(api, name: str, **settings_kw):
    """samples/{instancing,nested_instancing,curves}.usda from their parsed-stage fixtures (tests/golden/*.stage.json,
    made by tests/golden/make_stage_fixtures.py), through the importer: native instancing of class prototypes, a
    PointInstancer nested inside a prototype, BasisCurves (linear + cubic bezier)."""
    import json

    from .usd.importer import import_stage
    from .usd.usda import stage_from_json

    with open(os.path.join(GOLDEN, name + ".stage.json")) as f:
        stage = stage_from_json(json.load(f))
    return import_stage(api, stage, os.path.join(GOLDEN, name + ".usda"), **settings_kw)


def volume_slab(api, kind: str = "absorb", width: int = 32, height: int = 32, spp: int = 64, max_depth: int = 64, **settings_kw):
    """Analytic probes of the volume integrator: a wide 1-unit-thick slab facing a 1-degree camera under a uniform
    dome of radiance 1 (no geometry). `absorb`: gray sigma_a = 0.7 => pixel = e^-0.7. `emit`: the same slab emitting
    Le => Le (1 - e^-0.7) + e^-0.7. `furnace` / `furnace_grid` / `furnace_noise`: albedo-1 scattering (g = 0.5)
    => exactly 1 whatever the path does (volume NEE + phase-sampled escapes must MIS-sum to one). Gray on purpose: with
    chromatic sigma_s the reference's tracker folds the weaker channels' null collisions into real scatters
    (p_scatter = max-channel sigma_s / majorant, volume.rs:466-477), which loses energy in those channels — reproduced
    here as is, so a chromatic furnace does not return 1."""
    from .render_types import DensityField, VolumeRegion

    wb = api.WorldBuilder()
    lights = LightList()
    lights.add(DomeLight((1.0, 1.0, 1.0)))
    m = np.eye(4, dtype=np.float32)
    half = (50.0, 50.0, 0.5)
    le = (4.0, 1.5, 0.3)
    if kind == "absorb":
        regions = [VolumeRegion(m, half, (0,) * 3, (0.7,) * 3, 0.0, (0,) * 3, 1.0, DensityField.Homogeneous())]
    elif kind == "emit":
        regions = [VolumeRegion(m, half, (0,) * 3, (0.7,) * 3, 0.0, le, 1.0, DensityField.Homogeneous())]
    elif kind == "furnace":
        regions = [VolumeRegion(m, half, (1.2,) * 3, (0,) * 3, 0.5, (0,) * 3, 1.0, DensityField.Homogeneous())]
    elif kind == "furnace_grid":
        rng = np.random.default_rng(5)
        regions = [VolumeRegion(m, half, (1.0,) * 3, (0,) * 3, 0.5, (0,) * 3, 2.0, DensityField.Grid(4, 4, 4, rng.random(64).astype(np.float32))),
                   VolumeRegion(m, (50.0, 50.0, 0.25), (0.5,) * 3, (0,) * 3, -0.3, (0,) * 3, 1.0, DensityField.Homogeneous())]  # overlapping: mixed phase lobes
    elif kind == "furnace_noise":
        regions = [VolumeRegion(m, (2.0, 2.0, 0.5), (3.0,) * 3, (0,) * 3, 0.2, (0,) * 3, 1.0, DensityField.Noise(4.0, 3, 0.5, 2.0, 0.2, 7))]
    else:
        raise ValueError(kind)
    world = wb.commit(lights).with_volumes(regions)
    world.volumes = regions
    cam = Camera((0.0, 0.0, 20.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1.0, width / height)
    kw = dict(variance_threshold=0.0, min_samples_per_pixel=spp)
    kw.update(settings_kw)
    return world, cam, RenderSettings(spp, max_depth, width, height, **kw)


# ---------------------------------------------------------------------------------------------
# config 4 stand-in: a point-instanced city (two-level BVH)
# ---------------------------------------------------------------------------------------------
def _building(rng, floors: int, sides: int):
    """A closed prism tower with setbacks: ~1.1-1.3 k triangles like the MedCity prototypes."""
    verts, tris = [], []
    rings = floors * 6
    r0 = 1.0
    for k in range(rings + 1):
        y = k / rings * floors * 0.6
        r = r0 * (1.0 - 0.35 * (k // (rings // 3 + 1)) / 3.0) * (1.0 + 0.04 * math.sin(k * 1.7))
        for s in range(sides):
            a = 2.0 * math.pi * s / sides
            verts.append((r * math.cos(a), y, r * math.sin(a)))
    for k in range(rings):
        for s in range(sides):
            a, b = k * sides + s, k * sides + (s + 1) % sides
            c, d = a + sides, b + sides
            tris += [(a, c, b), (b, c, d)]
    top = len(verts)
    verts.append((0.0, floors * 0.6, 0.0))
    for s in range(sides):
        tris.append((rings * sides + s, top, rings * sides + (s + 1) % sides))
    return np.asarray(verts, np.float32), np.asarray(tris, np.uint32)


def point_instanced_city(api, n_instances: int = 40000, width: int = 640, height: int = 360, spp: int = 128, seed: int = 7, **settings_kw):
    """A synthetic, scalable relative of config 4 (the real file is `medcity()`): one top-level scene of
    `n_instances` instances (T * R * S per instance, usd_import.rs:1775-1782) over 4 distinct ~1.1-1.3 k-triangle
    prototype BLASes, Z-up, default grey OpenPBR, no lights (sky only, no shadow rays), importer default settings."""
    rng = np.random.default_rng(seed)
    protos = []
    for floors, sides in ((9, 10), (8, 11), (10, 10), (9, 11)):
        v, t = _building(rng, floors, sides)
        v = v[:, [0, 2, 1]].copy()  # Z-up
        t = t[:, [0, 2, 1]].copy()
        sb = api.SceneBuilder()
        sb.attach(Geometry.TriangleMesh(v, t))
        protos.append((sb.commit(), len(t)))
    wb = api.WorldBuilder()
    side = int(math.ceil(math.sqrt(n_instances)))
    grey = OpenPBR()
    k = 0
    for iy in range(side):
        for ix in range(side):
            if k >= n_instances:
                break
            yaw = float(rng.uniform(0.0, 2.0 * math.pi))
            s = float(rng.uniform(0.7, 1.3))
            pos = ((ix - side / 2) * 3.0 + float(rng.uniform(-0.5, 0.5)), (iy - side / 2) * 3.0 + float(rng.uniform(-0.5, 0.5)), 0.0)
            xf = Affine3A.from_translation(pos) * Affine3A.from_rotation_z(yaw) * Affine3A.from_scale((s, s, s * float(rng.uniform(0.6, 1.8))))
            wb.attach(Geometry.Instance(protos[int(rng.integers(0, 4))][0], xf), grey)
            k += 1
    wb.attach(Geometry.TriangleMesh(np.asarray([(-side * 2, -side * 2, 0), (side * 2, -side * 2, 0), (side * 2, side * 2, 0), (-side * 2, side * 2, 0)], np.float32), QUAD), grey)
    world = wb.commit(LightList())
    ext = side * 1.5
    cam = Camera((-ext * 0.9, -ext * 1.1, ext * 0.55), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 40.0, width / height, 0.0, 10.0)
    kw = dict(min_samples_per_pixel=32, variance_threshold=0.05)
    kw.update(settings_kw)
    return world, cam, RenderSettings(spp, 32, width, height, **kw), sum(n for _, n in protos)
