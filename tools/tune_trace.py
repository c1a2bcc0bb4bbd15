#!/usr/bin/env python
"""Tuning harness: times the five bench ray sets on one build of libcrt_b200 (CRT_B200_LIB selects
it) and prints Mrays/s per set plus an md5 of every output buffer, so builds with different tunables
can be compared for speed and checked to be result-identical. Not part of the product path."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crust_render_b200 as crt  # noqa: E402
from crust_render_b200 import workloads as W  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=32)
ap.add_argument("--rays-log2", type=int, default=23)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--cache", default="/tmp/tune_rays.npz")
ap.add_argument("--tag", default=os.environ.get("CRT_B200_LIB", "default"))
a = ap.parse_args()

crt.init(0)
b, n_tris = W.stress_scene(crt, grid=a.grid, shards=30000, jitter=True)
scene = b.commit()
if os.path.exists(a.cache):
    z = np.load(a.cache)
    sets = {k: (z[k].view(crt.RAY_DTYPE).reshape(-1), (z[k + "_mask"] if k + "_mask" in z else None)) for k in bench.SETS}
else:
    sets, _ = bench.build_ray_sets(W, scene.bounds(), a.rays_log2, 0, lambda r: scene.intersect_n(r)[0])
    out = {}
    for k, (r, m) in sets.items():
        out[k] = r.view(np.float32).reshape(-1, 8)
        if m is not None:
            out[k + "_mask"] = m
    np.savez(a.cache, **out)
res = {"tag": a.tag, "tris": n_tris}
stream = torch.cuda.current_stream()
for name in bench.SETS:
    rays, mask = sets[name]
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_m = torch.from_numpy(mask.view(np.int32)).cuda() if mask is not None else None
    closest = name.startswith("closest")
    d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda") if closest else torch.empty(n, dtype=torch.uint8, device="cuda")

    def go():
        mk = d_m.data_ptr() if d_m is not None else 0
        if closest:
            scene.intersect_ptr(n, d_r.data_ptr(), 0, mk, d_o.data_ptr(), 0, stream.cuda_stream)
        else:
            scene.occluded_ptr(n, d_r.data_ptr(), 0, mk, d_o.data_ptr(), stream.cuda_stream)

    for _ in range(2):
        go()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        go()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name] = round(n / best / 1e3, 1)
    res[name + "_md5"] = hashlib.md5(d_o.cpu().numpy().tobytes()).hexdigest()[:10]
res["status"] = crt.device_status()
print(json.dumps(res), flush=True)
