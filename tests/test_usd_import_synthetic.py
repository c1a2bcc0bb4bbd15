"""Importer decisions on small synthetic stages (text USD written inline), so they also run where the reference's sample
files are absent: BasisCurves width interpolation and bases (usd_import.rs:1993-2131), nested PointInstancers with
invisibleIds (:1549-1614, 1702-1792), native instancing through inherits (:166-188), volume prims that must not become
geometry and malformed ones that are skipped (:435-503), inactive and abstract prims (:146-160)."""
import numpy as np
import pytest

from crust_render_b200.rt_types import Ray
from crust_render_b200.usd import usda
from crust_render_b200.usd.importer import curve_segments, import_stage

CAM = 'def Camera "cam" { double3 xformOp:translate = (0, 0, 10)\n uniform token[] xformOpOrder = ["xformOp:translate"] }\n'


def load(orc, body):
    return import_stage(orc, usda.parse('#usda 1.0\n' + CAM + body), "<inline>")


def hit(world, origin, direction=(0, 0, -1), t_max=100.0):
    return world.scene.intersect(Ray(origin, direction), 0.001, t_max)


def test_curve_widths_per_vertex_per_curve_and_constant():
    def prim(widths):
        st = usda.parse(f'''#usda 1.0
def BasisCurves "c" {{
    uniform token type = "linear"
    int[] curveVertexCounts = [2, 3]
    point3f[] points = [(0,0,0), (1,0,0), (0,1,0), (1,1,0), (2,1,0)]
    float[] widths = {widths}
}}''')
        return curve_segments(st.roots[0])[0]
    segs = prim("[0.2, 0.4, 0.6, 0.8, 1.0]")  # one per point
    assert [round(s.r0, 3) for s in segs] == [0.1, 0.3, 0.4] and [round(s.r1, 3) for s in segs] == [0.2, 0.4, 0.5]
    segs = prim("[0.2, 0.6]")  # one per curve
    assert [round(s.r0, 3) for s in segs] == [0.1, 0.3, 0.3]
    segs = prim("[0.5]")  # constant
    assert all(abs(s.r0 - 0.25) < 1e-7 and abs(s.r1 - 0.25) < 1e-7 for s in segs)
    segs = prim("[0.0]")  # floored at 1e-6 before halving (:2080)
    assert all(abs(s.r0 - 5e-7) < 1e-12 for s in segs)


def test_cubic_spans_per_basis_and_unknown_basis():
    def spans(basis, n):
        pts = ", ".join(f"({i},{i % 2},0)" for i in range(n))
        st = usda.parse(f'''#usda 1.0
def BasisCurves "c" {{
    uniform token basis = "{basis}"
    int[] curveVertexCounts = [{n}]
    point3f[] points = [{pts}]
}}''')
        r = curve_segments(st.roots[0])
        return None if r is None else len(r[1])
    assert spans("bezier", 7) == 2 and spans("bezier", 4) == 1 and spans("bezier", 6) == 1  # vstep 3: (cnt - 4) / 3 + 1
    assert spans("bspline", 7) == 4 and spans("catmullRom", 5) == 2                        # vstep 1
    assert spans("bezier", 3) is None                                                       # fewer than 4 points: nothing
    assert spans("hermite", 8) is None                                                      # unsupported basis: skipped


def test_native_instancing_through_inherits_and_inactive_prims(orc):
    world, _, _ = load(orc, '''
class Xform "_Proto" { def Sphere "s" { double radius = 0.5 } }
def Xform "A" (instanceable = true; inherits = </_Proto>) { double3 xformOp:translate = (-2, 0, 0)
 uniform token[] xformOpOrder = ["xformOp:translate"] }
def Xform "B" (instanceable = true; inherits = </_Proto>) { double3 xformOp:translate = (2, 0, 0)
 float3 xformOp:scale = (1, 3, 1)
 uniform token[] xformOpOrder = ["xformOp:translate", "xformOp:scale"] }
def Sphere "hidden" (active = false) { double radius = 5 }
''')
    assert world.count() == 2 and world.scene.primitive_breakdown().instances == 2
    assert hit(world, (0, 0, 10)) is None                      # neither the class nor the inactive sphere is drawn
    assert abs(hit(world, (-2, 0, 10)).t - 9.5) < 1e-4
    assert hit(world, (2, 1.2, 10)) is not None and hit(world, (-2, 1.2, 10)) is None  # B is an ellipsoid, 1.5 high


def test_nested_point_instancer_with_invisible_ids(orc):
    world, _, _ = load(orc, '''
def PointInstancer "outer" {
    rel prototypes = [</outer/Protos/Cluster>]
    int[] protoIndices = [0, 0]
    point3f[] positions = [(-3, 0, 0), (3, 0, 0)]
    def Scope "Protos" {
        def PointInstancer "Cluster" {
            rel prototypes = [</outer/Protos/Cluster/P/ball>]
            int[] protoIndices = [0, 0, 0]
            int64[] ids = [7, 8, 9]
            int64[] invisibleIds = [8]
            point3f[] positions = [(0, -1, 0), (0, 0, 0), (0, 1, 0)]
            def Scope "P" { def Sphere "ball" { double radius = 0.25 } }
        }
    }
}
''')
    assert world.count() == 2                          # one part per outer placement: the nested instancer stays nested
    assert world.scene.primitive_count() == 2
    for x in (-3.0, 3.0):
        assert hit(world, (x, -1, 10)) is not None and hit(world, (x, 1, 10)) is not None
        assert hit(world, (x, 0, 10)) is None          # id 8 is invisible
    assert hit(world, (0, 0, 10)) is None              # prototypes are not drawn on their own


def test_volume_prims_are_regions_not_geometry_and_bad_ones_are_skipped(orc):
    world, _, _ = load(orc, '''
def Cube "fog" { double size = 2
 token crust:volume:type = "homogeneous"
 color3f crust:volume:sigmaS = (0.5, 0.5, 0.5) }
def Cube "badgrid" { token crust:volume:type = "grid"
 int[] crust:volume:gridDims = [2, 2, 2]
 float[] crust:volume:gridData = [1, 2, 3] }
def Cube "nogrid" { token crust:volume:type = "grid" }
def Cube "mystery" { token crust:volume:type = "plasma" }
def Sphere "ball" { double radius = 0.5 }
''')
    assert world.count() == 1 and len(world.volumes) == 1   # only the sphere is geometry; only the well-formed volume survives
    v = orc.Volumes(world.volumes)
    info = v.region(0)
    assert info["homogeneous"] and np.allclose(info["aabb_min"], (-1, -1, -1)) and np.allclose(info["aabb_max"], (1, 1, 1))
    assert abs(info["majorant"] - 0.5) < 1e-7               # defaults: sigmaA = 0, densityScale = 1 (:482-486)


def test_volume_inside_a_prototype_is_not_instanced(orc):
    world, _, _ = load(orc, '''
class Xform "_P" {
  def Sphere "s" { double radius = 0.5 }
  def Cube "smoke" { token crust:volume:type = "homogeneous" }
}
def Xform "I" (instanceable = true; references = </_P>) {}
''')
    assert world.count() == 1 and len(world.volumes) == 0   # usd_import.rs:1497-1506
