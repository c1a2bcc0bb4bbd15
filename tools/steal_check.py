#!/usr/bin/env python
"""Multi-GPU check of the distributed render paths (run under torchrun, one rank per GPU): the adaptive
openpbr_showcase frame rendered (a) with the static interleaved tile deal + NCCL sum-reduce, (b) with the shared CUDA-IPC
work-stealing cursor + reduce, (c) with the static deal and the peer-memory gather (every rank's film kernel stores into
rank 0's IPC-mapped frame, no frame collective) and (d) stealing + peer gather must be bit-identical on rank 0, and
identical to a single-rank render. Prints one JSON line with the timings (max over ranks)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import crust_render_b200 as crt  # noqa: E402
from crust_render_b200 import scenes  # noqa: E402
from crust_render_b200.distributed import render_distributed  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world_size = int(os.environ.get("WORLD_SIZE", "1"))
if world_size > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
crt.init(local)
rank = dist.get_rank() if world_size > 1 else 0
W, H, SPP = int(os.environ.get("STEAL_W", 1920)), int(os.environ.get("STEAL_H", 1080)), 128
world, cam, st = scenes.openpbr_showcase(crt, width=W, height=H, spp=SPP)
st.steal_chunk_tiles = int(os.environ.get("STEAL_CHUNK", 0))
r = crt.Renderer(cam, world, st)
fb = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
d = dist if world_size > 1 else None
out = {}
frames = {}
for mode, steal, gather in (("static", False, "reduce"), ("steal", True, "reduce"), ("peer", False, "peer"), ("peer_steal", True, "peer")):
    render_distributed(r, d, fb=fb, stream=stream, steal=steal, gather=gather)  # warm-up
    best = 1e9
    for _ in range(3):
        if d:
            d.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frame, stats = render_distributed(r, d, fb=fb, stream=stream, steal=steal, gather=gather)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if d:
            d.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    out[mode + "_ms"] = round(best, 2)
    frames[mode] = frame.clone()
    out[mode + "_camera_rays"] = stats.camera_rays
if rank == 0:
    single, s1 = r.render_with_stats()
    out["static_equals_steal"] = bool(torch.equal(frames["static"], frames["steal"]))
    out["static_equals_peer"] = bool(torch.equal(frames["static"], frames["peer"]))
    out["static_equals_peer_steal"] = bool(torch.equal(frames["static"], frames["peer_steal"]))
    out["equals_single_rank"] = bool(np.array_equal(frames["steal"].cpu().numpy(), single))
    out.update(world=world_size, width=W, height=H, spp=SPP, mean_taken=float(single[..., 3].mean()))
    print(json.dumps(out), flush=True)
if d:
    d.barrier()
    d.destroy_process_group()
