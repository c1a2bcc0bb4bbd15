#!/usr/bin/env python
"""Builds tests/golden/veach_mis_ref_240x135.npy from the reference's own preview render
(/root/reference/images/veach_mis_test.png, 960x540 sRGB, written by crust-render's write_png,
crates/crust-render/src/main.rs:459-468): decode sRGB -> linear, average 4x4 blocks, store float16.
It is the only image the reference repository ships of a bundled sample, so it is the one golden
vector that pins the integrator + OpenPBR + MIS restatement against the real crust-render (the
sample count and sampler it was rendered with are unknown: the comparison is statistical)."""
import os
import sys

import cv2
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/images/veach_mis_test.png"
img = cv2.imread(src, cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float64) / 255.0  # RGB, row 0 = top
lin = np.where(img <= 0.04045, img / 12.92, ((img + 0.055) / 1.055) ** 2.4)
h, w, _ = lin.shape
small = lin.reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "veach_mis_ref_240x135.npy")
np.save(out, small.astype(np.float16))
print(out, small.shape, small.mean())
