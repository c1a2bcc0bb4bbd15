#!/usr/bin/env python
"""tests/golden/sky_env_128x64.npy = /root/reference/samples/sky_env.exr decoded to float32 RGB [64, 128, 3]
(OpenCV's EXR reader; row 0 = top of the sky). The lat-long map of samples/domelight.usda, so the dome-light scene
runs on the GPU box where the reference tree is absent."""
import os
import sys

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
import cv2  # noqa: E402
import numpy as np  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/samples/sky_env.exr"
im = cv2.imread(src, cv2.IMREAD_UNCHANGED)
rgb = np.ascontiguousarray(im[:, :, 2::-1].astype(np.float32))
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sky_env_128x64.npy")
np.save(dst, rgb)
print(dst, rgb.shape, float(rgb.max()))
