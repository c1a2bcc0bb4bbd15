#!/usr/bin/env python
"""Builds tests/golden/cornellbox_scene.npz from /root/reference/samples/cornellbox.usda with this
repo's importer (crust-render_b200/usd/importer.py): the five meshes' authored points, triangulated
indices and composed local-to-world matrices, plus the camera. scenes.cornellbox() rebuilds the
World from it with the importer's bake-vs-instance rule, so BASELINE.json's config 1 runs on the GPU
box, where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crust_render_b200.usd import usda  # noqa: E402
from crust_render_b200.usd.importer import local_matrix, triangulate, _mul, _ident  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/samples/cornellbox.usda"
stage = usda.load(src)
top = stage.roots[0]
top_m, _ = local_matrix(top)
out = {}
k = 0
for c in top.children:  # >= 4 subtrees: imported in authored order (stream_roots)
    m, _ = local_matrix(c)
    world = _mul(top_m, m)
    if c.type == "Mesh":
        pts = np.asarray(c.attr("points"), np.float64).reshape(-1, 3).astype(np.float32)
        tris = triangulate(np.asarray(c.attr("faceVertexCounts")).astype(np.int64), np.asarray(c.attr("faceVertexIndices")).astype(np.int64), len(pts))
        out[f"mesh{k}_name"] = np.frombuffer(c.name.encode(), np.uint8)
        out[f"mesh{k}_points"], out[f"mesh{k}_tris"], out[f"mesh{k}_l2w"] = pts, tris.astype(np.uint16 if len(pts) < 65536 else np.uint32), world
        k += 1
    elif c.type == "Camera":
        out["camera_l2w"] = world
        out["camera_params"] = np.asarray([c.attr("focalLength"), c.attr("horizontalAperture"), c.attr("verticalAperture"), 0.0 if c.attr("fStop") is None else c.attr("fStop"),
                                           c.attr("focusDistance")], np.float32)
out["n_meshes"] = np.asarray(k)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornellbox_scene.npz")
np.savez_compressed(dst, **out)
print(dst, k, "meshes", sum(len(out[f"mesh{i}_tris"]) for i in range(k)), "triangles", os.path.getsize(dst), "bytes")
