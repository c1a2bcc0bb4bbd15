#!/usr/bin/env python
"""Tuning harness for the instanced traversal variants: times closest-hit / occlusion on K2 (instanced
stress scene), K3 instance_scene and the 40 000-instance MedCity scene, plus the MedCity render, on one
build of libcrt_b200 (CRT_B200_LIB selects it) and prints Mrays/s with an md5 of every output buffer.
Development tool (no oracle involved)."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crust_render_b200 as crt  # noqa: E402
from crust_render_b200 import scenes  # noqa: E402
from crust_render_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--rays-log2", type=int, default=22)
ap.add_argument("--tag", default=os.environ.get("CRT_B200_LIB", "default"))
ap.add_argument("--only", default="", help="comma list of: k2,k3,medcity,render")
ap.add_argument("--once", action="store_true", help="one launch per case, no timing loop (for ncu)")
a = ap.parse_args()
only = set(a.only.split(",")) if a.only else {"k2", "k3", "medcity", "render"}

crt.init(0)
stream = torch.cuda.current_stream()
res = {"tag": a.tag}


def rate(scene, rays, closest):
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda") if closest else torch.empty(n, dtype=torch.uint8, device="cuda")

    def go():
        if closest:
            scene.intersect_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), 0, stream.cuda_stream)
        else:
            scene.occluded_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), stream.cuda_stream)

    go()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(0 if a.once else a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        go()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out = d_o.cpu().numpy()
    return round(n / best / 1e3, 1), hashlib.md5(out.tobytes()).hexdigest()[:10], out


def both(name, scene, rays):
    res[name + "_closest"], res[name + "_closest_md5"], hits = rate(scene, rays, True)
    res[name + "_occluded"], res[name + "_occluded_md5"], _ = rate(scene, rays, False)
    return hits


if "k2" in only:
    s = W.stress_scene_instanced(crt, grid=a.grid).commit()
    both("k2", s, W.ray_batch(1 << a.rays_log2, W.scene_extent(s.bounds())))
    del s
if "k3" in only:
    s = W.instance_scene(crt).commit()
    both("k3inst", s, W.ray_batch(1 << a.rays_log2, W.scene_extent(s.bounds())))
    del s
if "medcity" in only or "render" in only:
    world, cam, st = scenes.medcity(crt)
    if "medcity" in only:  # the reference's LCG box rays over the city's extent (the render below covers camera rays)
        both("medcity_lcg", world.scene, W.ray_batch(1 << a.rays_log2, W.scene_extent(world.scene.bounds())))
    if "render" in only:
        fb = torch.zeros((st.height, st.width, 4), dtype=torch.float32, device="cuda")
        r = crt.Renderer(cam, world, st)
        best = None
        for k in range(1 if a.once else 4):
            stats = r.render_device(fb.data_ptr(), stream.cuda_stream)
            if (k or a.once) and (best is None or r.device_ms < best):
                best = r.device_ms
        res["medcity_render_ms"] = round(best, 2)
        res["medcity_render_mrays"] = round(stats.total_rays() / best / 1e3, 1)
        res["medcity_render_md5"] = hashlib.md5(fb.cpu().numpy().tobytes()).hexdigest()[:10]
res["status"] = crt.device_status()
print(json.dumps(res), flush=True)
