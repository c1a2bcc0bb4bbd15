"""usd/compose.py: a stage spread over several layer files — a sublayer carrying the render settings and the camera, an
asset referenced twice with local overrides, a payload holding the mesh data, variant sets that pick geometry and a material
binding — imports exactly like the same stage written out flat in one file (the reference reads composed stages through
openusd with payloads loaded, crates/crust-core/src/scene/usd_import.rs:271-300)."""
import os

import numpy as np
import pytest

from crust_render_b200.usd import compose, usda
from crust_render_b200.usd.importer import load_scene

CUBE_POINTS = "[(-1,-1,-1),(1,-1,-1),(1,1,-1),(-1,1,-1),(-1,-1,1),(1,-1,1),(1,1,1),(-1,1,1)]"
CUBE_COUNTS = "[4,4,4,4,4,4]"
CUBE_INDICES = "[0,3,2,1, 4,5,6,7, 0,1,5,4, 2,3,7,6, 1,2,6,5, 3,0,4,7]"

SETTINGS = '''#usda 1.0
(
    defaultPrim = "World"
)
def Scope "Render"
{
    def RenderSettings "settings"
    {
        uniform int2 resolution = (48, 32)
        custom int crust:samplesPerPixel = 4
        custom int crust:maxDepth = 3
    }
}
def Xform "World"
{
    def Camera "Cam"
    {
        double3 xformOp:translate = (0, 1, 12)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        float focalLength = 35
        float horizontalAperture = 36
    }
}
'''

GEO = '''#usda 1.0
(
    defaultPrim = "Geo"
)
def Mesh "Geo"
{
    point3f[] points = %s
    int[] faceVertexCounts = %s
    int[] faceVertexIndices = %s
}
''' % (CUBE_POINTS, CUBE_COUNTS, CUBE_INDICES)

ASSET = '''#usda 1.0
(
    defaultPrim = "Asset"
)
def Xform "Asset" (
    variants = { string shape = "box" }
    prepend variantSets = "shape"
)
{
    def Scope "Looks"
    {
        def Material "Red"
        {
            token outputs:surface.connect = </Asset/Looks/Red/Surface.outputs:surface>
            def Shader "Surface"
            {
                uniform token info:id = "crust:openpbr"
                color3f inputs:baseColor = (0.8, 0.1, 0.1)
                token outputs:surface
            }
        }
        def Material "Blue"
        {
            token outputs:surface.connect = </Asset/Looks/Blue/Surface.outputs:surface>
            def Shader "Surface"
            {
                uniform token info:id = "crust:openpbr"
                color3f inputs:baseColor = (0.1, 0.1, 0.8)
                token outputs:surface
            }
        }
    }
    variantSet "shape" = {
        "box" {
            def Mesh "Body" (
                payload = @./geo.usda@
            )
            {
                rel material:binding = </Asset/Looks/Red>
            }
        }
        "ball" {
            def Sphere "Body"
            {
                double radius = 1.5
                rel material:binding = </Asset/Looks/Blue>
            }
        }
    }
}
'''

ROOT = '''#usda 1.0
(
    subLayers = [@./settings.usda@]
    defaultPrim = "World"
)
def Xform "World"
{
    def Xform "A" (
        prepend references = @./assets/asset.usda@</Asset>
    )
    {
        double3 xformOp:translate = (-2.5, 0, 0)
        uniform token[] xformOpOrder = ["xformOp:translate"]
    }
    def Xform "B" (
        references = @./assets/asset.usda@
        variants = { string shape = "ball" }
    )
    {
        double3 xformOp:translate = (2.5, 0, 0)
        uniform token[] xformOpOrder = ["xformOp:translate"]
    }
    def SphereLight "Key"
    {
        double3 xformOp:translate = (0, 6, 4)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        float inputs:radius = 0.5
        float inputs:intensity = 40
    }
}
'''

FLAT = '''#usda 1.0
(
    defaultPrim = "World"
)
def Xform "World"
{
    def Xform "A"
    {
        double3 xformOp:translate = (-2.5, 0, 0)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        def Scope "Looks"
        {
            def Material "Red"
            {
                token outputs:surface.connect = </World/A/Looks/Red/Surface.outputs:surface>
                def Shader "Surface"
                {
                    uniform token info:id = "crust:openpbr"
                    color3f inputs:baseColor = (0.8, 0.1, 0.1)
                    token outputs:surface
                }
            }
            def Material "Blue"
            {
                token outputs:surface.connect = </World/A/Looks/Blue/Surface.outputs:surface>
                def Shader "Surface"
                {
                    uniform token info:id = "crust:openpbr"
                    color3f inputs:baseColor = (0.1, 0.1, 0.8)
                    token outputs:surface
                }
            }
        }
        def Mesh "Body"
        {
            rel material:binding = </World/A/Looks/Red>
            point3f[] points = %s
            int[] faceVertexCounts = %s
            int[] faceVertexIndices = %s
        }
    }
    def Xform "B"
    {
        double3 xformOp:translate = (2.5, 0, 0)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        def Scope "Looks"
        {
            def Material "Red"
            {
                token outputs:surface.connect = </World/B/Looks/Red/Surface.outputs:surface>
                def Shader "Surface"
                {
                    uniform token info:id = "crust:openpbr"
                    color3f inputs:baseColor = (0.8, 0.1, 0.1)
                    token outputs:surface
                }
            }
            def Material "Blue"
            {
                token outputs:surface.connect = </World/B/Looks/Blue/Surface.outputs:surface>
                def Shader "Surface"
                {
                    uniform token info:id = "crust:openpbr"
                    color3f inputs:baseColor = (0.1, 0.1, 0.8)
                    token outputs:surface
                }
            }
        }
        def Sphere "Body"
        {
            double radius = 1.5
            rel material:binding = </World/B/Looks/Blue>
        }
    }
    def SphereLight "Key"
    {
        double3 xformOp:translate = (0, 6, 4)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        float inputs:radius = 0.5
        float inputs:intensity = 40
    }
    def Camera "Cam"
    {
        double3 xformOp:translate = (0, 1, 12)
        uniform token[] xformOpOrder = ["xformOp:translate"]
        float focalLength = 35
        float horizontalAperture = 36
    }
}
def Scope "Render"
{
    def RenderSettings "settings"
    {
        uniform int2 resolution = (48, 32)
        custom int crust:samplesPerPixel = 4
        custom int crust:maxDepth = 3
    }
}
''' % (CUBE_POINTS, CUBE_COUNTS, CUBE_INDICES)


@pytest.fixture()
def layered(tmp_path):
    (tmp_path / "assets").mkdir()
    (tmp_path / "root.usda").write_text(ROOT)
    (tmp_path / "settings.usda").write_text(SETTINGS)
    (tmp_path / "assets" / "asset.usda").write_text(ASSET)
    (tmp_path / "assets" / "geo.usda").write_text(GEO)
    (tmp_path / "flat.usda").write_text(FLAT)
    return tmp_path


def _outline(prim, depth=0):
    rows = [(depth, prim.type, prim.name, sorted(prim.attributes), {k: v for k, v in prim.relationships.items()})]
    for c in prim.children:
        rows += _outline(c, depth + 1)
    return rows


def test_composed_stage_has_the_flat_stage_s_prims(layered):
    st = compose.load_composed(str(layered / "root.usda"))
    world = st.prim_at("/World")
    a_body, b_body = st.prim_at("/World/A/Body"), st.prim_at("/World/B/Body")
    assert a_body.type == "Mesh" and len(a_body.attr("points")) == 8                    # variant "box" (the asset's own selection) + payload
    assert b_body.type == "Sphere" and b_body.attr("radius") == 1.5                     # the referencing prim's selection wins
    assert a_body.relationships["material:binding"] == ["/World/A/Looks/Red"]           # paths re-rooted into the destination namespace
    assert b_body.relationships["material:binding"] == ["/World/B/Looks/Blue"]
    assert st.prim_at("/World/A").attr("xformOp:translate") == (-2.5, 0, 0)             # local opinion kept
    assert st.prim_at("/World/Cam") is not None and st.prim_at("/Render/settings").attr("crust:samplesPerPixel") == 4  # sublayer
    assert "references" not in st.prim_at("/World/A").metadata and "payload" not in a_body.metadata
    flat = usda.load(str(layered / "flat.usda"))
    norm = lambda rows: sorted((d, t, n, tuple(a), tuple(sorted((k, tuple(v)) for k, v in r.items()))) for d, t, n, a, r in rows)  # noqa: E731
    assert norm(_outline(world)) == norm(_outline(flat.prim_at("/World")))


def test_layered_stage_imports_like_the_flat_one(layered, orc):
    wl, cl, sl = load_scene(orc, str(layered / "root.usda"))
    wf, cf, sf = load_scene(orc, str(layered / "flat.usda"))
    assert (sl.width, sl.height, sl.samples_per_pixel, sl.max_depth) == (48, 32, 4, 3) == (sf.width, sf.height, sf.samples_per_pixel, sf.max_depth)
    assert wl.count() == wf.count() == 3  # the cube mesh, the sphere, the light's sphere — in the same geom_id order
    assert np.array_equal(wl.materials, wf.materials) and np.array_equal(wl.geom_material, wf.geom_material)
    fl, stl = orc.Renderer(cl, wl, sl).render_with_stats()
    ff, stf = orc.Renderer(cf, wf, sf).render_with_stats()
    assert fl.tobytes() == ff.tobytes() and stl == stf
    reds = wl.materials["base_color"]
    assert any(np.allclose(c, (0.8, 0.1, 0.1)) for c in reds) and any(np.allclose(c, (0.1, 0.1, 0.8)) for c in reds)


def test_reference_cycles_and_missing_targets_are_errors(layered):
    (layered / "loop_a.usda").write_text('#usda 1.0\ndef Xform "P" ( references = @./loop_b.usda@</Q> )\n{\n}\n')
    (layered / "loop_b.usda").write_text('#usda 1.0\ndef Xform "Q" ( references = @./loop_a.usda@</P> )\n{\n}\n')
    with pytest.raises(ValueError, match="cycle"):
        compose.load_composed(str(layered / "loop_a.usda"))
    (layered / "dangling.usda").write_text('#usda 1.0\ndef Xform "P" ( references = @./settings.usda@</Nope> )\n{\n}\n')
    with pytest.raises(ValueError, match="not found"):
        compose.load_composed(str(layered / "dangling.usda"))
