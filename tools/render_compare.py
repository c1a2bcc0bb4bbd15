#!/usr/bin/env python
"""GPU-vs-oracle render comparison on the small test scenes: prints the image metrics the parity
tests assert on (so tolerances in tests/test_gpu_render.py are set from measurements) and writes
PNG previews under gpurun_out/. Development tool; imports the oracle, so never part of the product."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import crust_render_b200 as crt  # noqa: E402
import orc  # noqa: E402
from crust_render_b200 import scenes  # noqa: E402


def metrics(a, b):
    a3, b3 = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    rel = (a3 - b3) ** 2 / (b3 ** 2 + 1e-4)
    absd = np.abs(a3 - b3) / (np.abs(b3) + 1e-3)
    return dict(relmse=float(rel.mean()), frac_gt_1e3=float((absd.max(axis=-1) > 1e-3).mean()), frac_gt_1e5=float((absd.max(axis=-1) > 1e-5).mean()),
                max_rel=float(absd.max()), taken_equal=float((a[..., 3] == b[..., 3]).mean()), mean_a=float(a3.mean()), mean_b=float(b3.mean()))


def save(fb, path):
    try:
        import cv2
    except ImportError:
        return
    img = np.clip(fb[::-1, :, :3], 0, 1) ** (1 / 2.2)
    cv2.imwrite(path, (img[:, :, ::-1] * 255).astype(np.uint8))


crt.init(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
cases = [
    ("room", scenes.furnace_room, dict(width=192, height=128, spp=16, variance_threshold=0.0)),
    ("room_adaptive", scenes.furnace_room, dict(width=192, height=128, spp=64, variance_threshold=0.05, min_samples_per_pixel=16)),
    ("veach", scenes.veach_mis, dict(width=240, height=135, spp=16)),
    ("showcase", scenes.openpbr_showcase, dict(width=320, height=180, spp=64)),
    ("stress", lambda api, **kw: scenes.stress(api, **kw)[:3], dict(grid=2, shards=2000, width=160, height=90, spp=8)),
]
for name, fn, kw in cases:
    wg, cam, st = fn(crt, **kw)
    wo, _, _ = fn(orc, **kw)
    rg = crt.Renderer(cam, wg, st)
    t0 = time.time()
    fg, sg = rg.render_with_stats()
    tg = time.time() - t0
    ro = orc.Renderer(cam, wo, st)
    fo, so = ro.render_with_stats()
    m = metrics(fg, fo)
    m.update(name=name, gpu_s=round(tg, 3), gpu_device_ms=round(rg.device_ms, 3), launches=rg.launches, rounds=rg.rounds, cpu_s=round(ro.seconds, 3),
             stats_gpu=sg.__dict__, stats_cpu=so.__dict__, nan=int(np.isnan(fg).sum()))
    print(json.dumps(m), flush=True)
    save(fg, os.path.join(ROOT, "gpurun_out", f"render_{name}_gpu.png"))
    save(fo, os.path.join(ROOT, "gpurun_out", f"render_{name}_cpu.png"))
print("status", crt.device_status())
