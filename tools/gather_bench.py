#!/usr/bin/env python
"""How should the per-rank frames be assembled? Times reduce / all_reduce of the full 4K RGBA32F frame and a
gather of only the owned rows (same bytes as owned tiles) under NCCL. Development tool (torchrun)."""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
rank, world = dist.get_rank(), dist.get_world_size()
fb = torch.rand((2160, 3840, 4), device="cuda")
part = fb[: 2160 // world].contiguous()
parts = [torch.empty_like(part) for _ in range(world)] if rank == 0 else None


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 3)


out = {"world": world}
out["reduce_ms"] = timed(lambda: dist.reduce(fb, dst=0, op=dist.ReduceOp.SUM))
out["all_reduce_ms"] = timed(lambda: dist.all_reduce(fb, op=dist.ReduceOp.SUM))
out["gather_owned_ms"] = timed(lambda: dist.gather(part, parts, dst=0))
full = torch.empty((world,) + tuple(part.shape), device="cuda")
out["all_gather_owned_ms"] = timed(lambda: dist.all_gather_into_tensor(full, part))
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
