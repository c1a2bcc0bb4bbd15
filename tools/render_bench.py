#!/usr/bin/env python
"""Times crt_render on the stress scene (config 5) and prints paths/s, rays/s and spp/s.
Development tool (no oracle involved)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crust_render_b200 as crt  # noqa: E402
from crust_render_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="stress")
ap.add_argument("--grid", type=int, default=32)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-paths", type=int, default=0)
ap.add_argument("--tag", default="")
a = ap.parse_args()
crt.init(0)
t0 = time.time()
if a.scene == "stress":
    world, cam, st, n_tris = scenes.stress(crt, grid=a.grid, jitter=True, width=a.width, height=a.height, spp=a.spp)
else:
    world, cam, st = getattr(scenes, a.scene)(crt, width=a.width, height=a.height, spp=a.spp)
    n_tris = world.primitive_count()
st.max_paths = a.max_paths
print(f"commit {time.time() - t0:.1f}s", file=sys.stderr)
fb = torch.zeros((a.height, a.width, 4), dtype=torch.float32, device="cuda")
r = crt.Renderer(cam, world, st)
best = None
for k in range(a.reps + 1):
    stats = r.render_device(fb.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if k and (best is None or r.device_ms < best[0]):
        best = (r.device_ms, stats, r.launches, r.rounds)
ms, stats, launches, rounds = best
samples = int(stats.camera_rays)
print(json.dumps({"tag": a.tag, "scene": a.scene, "prims": n_tris, "w": a.width, "h": a.height, "spp": a.spp, "ms": round(ms, 2), "launches": launches, "rounds": rounds,
                  "msamples_s": round(samples / ms / 1e3, 1), "mrays_s": round(stats.total_rays() / ms / 1e3, 1), "rays_per_sample": round(stats.total_rays() / samples, 2),
                  "spp_per_s": round(samples / (a.width * a.height) / ms * 1e3, 2), "mean": float(fb[..., :3].mean().item()), "status": crt.device_status()}))
