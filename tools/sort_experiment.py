#!/usr/bin/env python
"""Experiment: does reordering incoherent rays (direction octant, then Morton code of the origin) speed
up trace_kernel enough to pay for a device-side binning pass? Times the bench's ray sets as given and
host-sorted by several keys. Development tool."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crust_render_b200 as crt  # noqa: E402
from crust_render_b200 import workloads as W  # noqa: E402
import bench  # noqa: E402

crt.init(0)
b, n_tris = W.stress_scene(crt, grid=32, shards=30000, jitter=True)
scene = b.commit()
sets, _ = bench.build_ray_sets(W, scene.bounds(), 24, 0, lambda r: scene.intersect_n(r)[0])
bounds = scene.bounds()
lo, hi = np.asarray(bounds.minimum, np.float32), np.asarray(bounds.maximum, np.float32)
stream = torch.cuda.current_stream()


def part1by2(x):
    x = x.astype(np.uint32) & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x


def keys(rays, mode):
    o, d = rays["o"], rays["d"]
    octant = ((d[:, 0] < 0).astype(np.uint32) | ((d[:, 1] < 0).astype(np.uint32) << 1) | ((d[:, 2] < 0).astype(np.uint32) << 2))
    if mode == "octant":
        return octant
    q = np.clip((o - lo) / (hi - lo) * 1024, 0, 1023)
    m = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
    if mode == "morton":
        return m
    if mode == "octant_morton":
        return (octant.astype(np.uint64) << 30) | m
    if mode == "morton15_octant":  # coarse cell (5 bits per axis), then octant, then fine order
        return ((m >> 15).astype(np.uint64) << 33) | (octant.astype(np.uint64) << 30) | (m & 0x7FFF)
    raise ValueError(mode)


def time_set(rays, closest):
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda") if closest else torch.empty(n, dtype=torch.uint8, device="cuda")
    best = 1e9
    for k in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if closest:
            scene.intersect_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), 0, stream.cuda_stream)
        else:
            scene.occluded_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        if k:
            best = min(best, e0.elapsed_time(e1))
    return n / best / 1e3


for name in ("closest_lcg", "closest_bounce", "closest_camera", "occluded_lcg"):
    rays, mask = sets[name]
    if mask is not None:
        continue
    res = {"set": name, "as_given": round(time_set(rays, name.startswith("closest")), 1)}
    for mode in ("octant", "morton", "octant_morton", "morton15_octant"):
        order = np.argsort(keys(rays, mode), kind="stable")
        res[mode] = round(time_set(np.ascontiguousarray(rays[order]), name.startswith("closest")), 1)
    print(json.dumps(res), flush=True)
