#!/usr/bin/env python
"""commit() on the host (reference SAH/SBVH) vs on the device (CRT_COMMIT_DEVICE_BUILD) for the K1 stress mesh:
build time, tree size, and traversal throughput of the five bench ray sets on both trees, with the fraction of
rays whose closest hit is identical. Development / measurement tool."""
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
import bench  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 32
crt.init(0)
meshes = W.stress_meshes(grid, 30000, True)
stream = torch.cuda.current_stream()


def commit(flags):
    b = crt.SceneBuilder()
    for v, i in meshes:
        b.attach(crt.Geometry.TriangleMesh(v, i))
    t0 = time.perf_counter()
    s = b.commit(flags=flags)
    return s, time.perf_counter() - t0


commit(crt.COMMIT_DEVICE_BUILD)  # warm-up (context, cub temp allocations)
dev, t_dev = commit(crt.COMMIT_DEVICE_BUILD)
host, t_host = commit(0)
bi_d, bi_h = dev.build_info(), host.build_info()
out = {"triangles": int(bi_h.primitives), "host": {"commit_s": round(t_host, 3), "build_s": round(bi_h.build_seconds, 3), "upload_s": round(bi_h.upload_seconds, 3),
                                                   "wide_nodes": int(bi_h.wide_nodes), "leaves": int(bi_h.leaves), "leaf_refs": int(bi_h.leaf_refs)},
       "device": {"commit_s": round(t_dev, 4), "device_build_s": round(bi_d.build_seconds, 4), "concat_upload_s": round(bi_d.upload_seconds, 4),
                  "wide_nodes": int(bi_d.wide_nodes), "leaves": int(bi_d.leaves), "leaf_refs": int(bi_d.leaf_refs), "built_on_device": int(bi_d.built_on_device)}}
sets, _ = bench.build_ray_sets(W, host.bounds(), 23, 0, lambda r: host.intersect_n(r)[0])
res = {}
for name in bench.SETS:
    rays, mask = sets[name]
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_m = torch.from_numpy(mask.view(np.int32)).cuda() if mask is not None else None
    closest = name.startswith("closest")
    outs = {}
    for tag, sc in (("host_tree", host), ("device_tree", dev)):
        d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda") if closest else torch.empty(n, dtype=torch.uint8, device="cuda")
        best = 1e9
        for k in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mk = d_m.data_ptr() if d_m is not None else 0
            if closest:
                sc.intersect_ptr(n, d_r.data_ptr(), 0, mk, d_o.data_ptr(), 0, stream.cuda_stream)
            else:
                sc.occluded_ptr(n, d_r.data_ptr(), 0, mk, d_o.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            if k:
                best = min(best, e0.elapsed_time(e1))
        outs[tag] = (round(n / best / 1e3, 1), d_o)
    a, b = outs["host_tree"][1], outs["device_tree"][1]
    same = float((a.view(torch.int32) == b.view(torch.int32)).all(dim=1).float().mean().item()) if closest else float((a == b).float().mean().item())  # bit compare: miss ids are NaN as floats
    res[name] = {"host_tree_mrays": outs["host_tree"][0], "device_tree_mrays": outs["device_tree"][0], "identical": same}
out["traversal"] = res
out["status"] = crt.device_status()
print(json.dumps(out))
