#!/usr/bin/env python
"""Times the host-buffer query path (crt_intersect_host_n / crt_occluded_host_n) on pinned buffers for the
current CRT_HOST_CHUNK_LOG2. Development tool."""
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

crt.init(0)
b, n_tris = W.stress_scene(crt, grid=32, shards=30000, jitter=True)
scene = b.commit()
n = 1 << 24
rays = W.ray_batch(n, W.scene_extent(scene.bounds()))
h_in = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).pin_memory()
h_hit = torch.empty((n, 8), dtype=torch.float32).pin_memory()
h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()
out = {"chunk_log2": os.environ.get("CRT_HOST_CHUNK_LOG2", "default")}
for name, fn in (("closest", lambda: scene.intersect_host_ptr(n, h_in.data_ptr(), 0, 0, h_hit.data_ptr(), 0)),
                 ("occluded", lambda: scene.occluded_host_ptr(n, h_in.data_ptr(), 0, 0, h_occ.data_ptr()))):
    fn()
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    out[name + "_mrays"] = round(n / best / 1e6, 1)
    out[name + "_gbs_h2d"] = round(n * 32 / best / 1e9, 1)
print(json.dumps(out))
