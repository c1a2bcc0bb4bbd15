#!/usr/bin/env python
"""Builds tests/golden/medcity_scene.npz from /root/reference/samples/PointInstancedMedCity.usd with this
repo's crate reader (crust-render_b200/usd/usdc.py): the PointInstancer's per-instance arrays, the prototype
meshes (points + fan-triangulated indices + local matrices) and the camera, so BASELINE.json's config 4 runs
on the GPU box, where /root/reference does not exist. scenes.medcity() rebuilds the World with the
importer's placement arithmetic."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crust_render_b200.usd import usdc  # noqa: E402
from crust_render_b200.usd.importer import local_matrix, triangulate, _mul, _ident  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/samples/PointInstancedMedCity.usd"
stage = usdc.load(src)
inst = next(p for p in stage.walk() if p.type == "PointInstancer")
out = {}
m = _ident()
chain = []
p = inst
while p is not None:
    chain.append(p)
    p = p.parent
for q in reversed(chain):
    m = _mul(m, local_matrix(q)[0])
out["instancer_l2w"] = m
out["positions"] = np.asarray(inst.attr("positions"), np.float32)
out["orientations_wxyz"] = np.asarray(inst.attr("orientations"), np.float16)
out["scales"] = np.asarray(inst.attr("scales"), np.float32)
out["proto_indices"] = np.asarray(inst.attr("protoIndices")).astype(np.uint8)
targets = inst.relationships["prototypes"]
out["n_protos"] = np.asarray(len(targets))
for k, t in enumerate(targets):
    root = stage.prim_at(t)
    mesh = next(c for c in root.walk() if c.type == "Mesh")
    pts = np.asarray(mesh.attr("points"), np.float32)
    tris = triangulate(np.asarray(mesh.attr("faceVertexCounts")).astype(np.int64), np.asarray(mesh.attr("faceVertexIndices")).astype(np.int64), len(pts))
    local = _ident()
    q = mesh
    chain = []
    while q is not root:
        chain.append(q)
        q = q.parent
    for c in reversed(chain):
        local = _mul(local, local_matrix(c)[0])
    out[f"proto{k}_points"], out[f"proto{k}_tris"], out[f"proto{k}_local"] = pts, tris.astype(np.uint16), local
cam = next(p for p in stage.walk() if p.type == "Camera")
m = _ident()
chain = []
p = cam
while p is not None:
    chain.append(p)
    p = p.parent
for q in reversed(chain):
    m = _mul(m, local_matrix(q)[0])
out["camera_l2w"] = m
g = lambda n, d: d if cam.attr(n) is None else cam.attr(n)  # noqa: E731
out["camera_params"] = np.asarray([g("focalLength", 50.0), g("horizontalAperture", 20.955), g("verticalAperture", -1.0), g("fStop", 0.0), g("focusDistance", 10.0)], np.float32)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "medcity_scene.npz")
np.savez_compressed(dst, **out)
print(dst, len(out["positions"]), "instances", len(targets), "prototypes", os.path.getsize(dst), "bytes")
