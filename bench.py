#!/usr/bin/env python
"""bench.py — rays/s of the closest-hit + occlusion hot path on the 10 M-triangle stress mesh.

One step = one pass of the traversal kernels over five resident ray sets on the K1 scene
(SURVEY.md §8(d)): closest-hit on (i) LCG box rays, (ii) camera rays, (iii) cosine bounce
rays; occlusion on (i) and on (iv) RectLight shadow rays. `value` = rays of all five
launches / device time, inputs resident in HBM. `e2e` = the same five queries through the
host-buffer C ABI (crt_*_host_n) with pinned host rays in and hits out, copies inside the
timed region. `--impl reference` times the CPU restatement of crust-rt (oracle/) on all
host threads on a bounded sample of the same sets.

The same line carries the second half of BASELINE.json's metric, spp/s, under "render": config 5
(the stress mesh lit by its RectLight, 3840x2160, depth 6) rendered by the wavefront integrator
(crt_render) `--render-spp` samples per pixel per step; with N GPUs the frame is dealt out in
interleaved 32x32 tiles and sum-reduced to rank 0 over NCCL inside the timed region (strong scaling:
the frame is fixed). spp_per_s = samples per pixel of the whole 4K frame per second.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]; one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Mrays/s closest-hit+occlusion per GPU"


def workload_config(n_tris, rays_log2, rays_per_step, info):
    """The `config` object of the JSON line — identical in both arms (they run the same scene and the same ray sets)."""
    return {"workload": f"K1 stress mesh ({n_tris} triangles, flat BVH4), ray sets (i)-(iv) of SURVEY 8(d)", "triangles": n_tris,
            "rays_per_set": 1 << rays_log2, "rays_per_step_per_gpu": rays_per_step,
            "l2": "inputs larger than L2 (rays+hits 2.7 GB, BVH 0.9 GB per step)", "scene": "replicated per GPU", **info}
SETS = ["closest_lcg", "closest_camera", "closest_bounce", "occluded_lcg", "occluded_shadow"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: native libraries (NCCL's version banner) write to fd 1, so fd 1
# is pointed at stderr for the whole run and the line goes to a private duplicate of the real stdout.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=32, help="stress-scene sphere grid (32 = K1, 10.07 M triangles)")
    ap.add_argument("--rays-log2", type=int, default=24, help="rays per set (camera set: 3840x2160x2 spp when 24)")
    ap.add_argument("--cpu-sample-log2", type=int, default=24, help="rays per set for the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--render-spp", type=int, default=8, help="samples per pixel per render step (a 256-spp frame is 32 such steps)")
    ap.add_argument("--render-steps", type=int, default=3)
    ap.add_argument("--render-width", type=int, default=3840)
    ap.add_argument("--render-height", type=int, default=2160)
    ap.add_argument("--cpu-render", default="1920x1080x6", help="bounded sample of the render workload for the CPU baseline: WxHxSPP")
    return ap.parse_args()


def cpu_render_sample(args, log_prefix):
    """The oracle integrator on all host threads over a bounded sample of config 5 (same scene, camera
    and settings at reduced resolution / spp). Returns the `render.cpu_baseline` object."""
    import orc
    from crust_render_b200 import scenes

    w, h, spp = (int(v) for v in args.cpu_render.lower().split("x"))
    t0 = time.time()
    world, cam, st, n_tris = scenes.stress(orc, grid=args.grid, jitter=True, width=w, height=h, spp=spp)
    log(f"{log_prefix} oracle world commit ({n_tris} triangles): {time.time() - t0:.1f}s")
    r = orc.Renderer(cam, world, st)
    _, stats = r.render_with_stats(threads=0)
    cores = os.cpu_count() or 1
    return {"value": spp / r.seconds * (w * h) / (args.render_width * args.render_height), "unit": "spp/s of the 4K frame (scaled from the sample by pixel count)",
            "msamples_per_s": w * h * spp / r.seconds / 1e6, "mrays_per_s": stats.total_rays() / r.seconds / 1e6, "cores": cores, "kind": "port",
            "sample": f"{w}x{h} x {spp} spp of the same scene/camera/settings, {r.seconds:.1f}s on {cores} threads (rows dealt to threads)"}


SAMPLE_CONFIGS = [
    # (key, label, builder name, kwargs of the BASELINE size, kwargs of the scaled-up frame used for the multi-GPU curve)
    ("1_cornellbox", "config 1 samples/cornellbox.usda (822 baked triangles + 2 instances, sky only, adaptive, depth 32)", "cornellbox",
     dict(width=512, height=512, spp=64), dict(width=1920, height=1080, spp=64)),
    ("2_veach_mis", "config 2 samples/veach_mis.usda (4 sphere lights, balance MIS, depth 8, adaptive off)", "veach_mis",
     dict(width=960, height=540, spp=128), dict(width=1920, height=1080, spp=128)),
    ("3_openpbr_showcase", "config 3 samples/openpbr_showcase.usda (all OpenPBR lobes, 2 sphere lights, adaptive)", "openpbr_showcase",
     dict(width=640, height=360, spp=128), dict(width=1920, height=1080, spp=128)),
    ("4_medcity", "config 4 samples/PointInstancedMedCity.usd (40 000 instances, two-level BVH, sky only, adaptive, depth 32)", "medcity",
     dict(width=640, height=360, spp=128), dict(width=1920, height=1080, spp=128)),
]


def timed_render(renderer, dist, world, fb, stream, steps, gather, torch, dev):
    """`steps` frames through render_distributed after one warm-up; device time by CUDA events on the launching stream,
    max over ranks. Returns (ms per frame, RayStats of the last frame (whole job on rank 0), launches per frame on this rank)."""
    from crust_render_b200.distributed import render_distributed

    d = dist if world > 1 else None
    render_distributed(renderer, d, fb=fb, stream=stream.cuda_stream, gather=gather)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, stats = 0, None
    e0.record(stream)
    for _ in range(steps):
        _, stats = render_distributed(renderer, d, fb=fb, stream=stream.cuda_stream, gather=gather)
        launches += renderer.launches
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, stats, launches // steps


def time_sample_configs(crt, torch, dist, world, rank, dev, stream, gather, with_cpu, steps=2):
    """BASELINE.json configs 1-4 through crt_render on `world` GPUs (32x32 tiles dealt round-robin, frame gathered into rank
    0's HBM inside the timed region): each at the size the config list names and on a 1920x1080 frame (9x to 16x the
    pixels — the BASELINE frames are 0.23-0.52 Mpixel, i.e. 28-64 tiles per rank at 8 GPUs, too few to say anything about
    scaling). spp/s of the frame, Msamples/s, Mrays/s. Beside each BASELINE-size number (rank 0, N = 1 only): the CPU
    restatement of the integrator on all host threads at the same full settings."""
    from crust_render_b200 import scenes

    out = {}
    for key, label, fn, base_kw, big_kw in SAMPLE_CONFIGS:
        rec = {"workload": label}
        for tag, kw in (("baseline_size", base_kw), ("frame_1920x1080", big_kw)):
            world_k, cam_k, st_k = getattr(scenes, fn)(crt, **kw)
            r = crt.Renderer(cam_k, world_k, st_k)
            fb = torch.zeros((st_k.height, st_k.width, 4), dtype=torch.float32, device=dev)
            ms, stats, launches = timed_render(r, dist, world, fb, stream, steps, gather, torch, dev)
            if rank == 0:
                taken = stats.camera_rays / (st_k.width * st_k.height)
                rec[tag] = {"width": st_k.width, "height": st_k.height, "max_spp": st_k.samples_per_pixel, "ms": ms, "spp_per_s": taken / (ms * 1e-3),
                            "mean_spp_taken": taken, "msamples_per_s": stats.camera_rays / ms / 1e3, "mrays_per_s": stats.total_rays() / ms / 1e3,
                            "launches_rank0": launches, "n_gpus": world}
            del fb, r, world_k
        if with_cpu and rank == 0:
            import orc
            w_o, cam_o, st_o = getattr(scenes, fn)(orc, **base_kw)
            ro = orc.Renderer(cam_o, w_o, st_o)
            _, so = ro.render_with_stats(threads=0)
            taken = so.camera_rays / (st_o.width * st_o.height)
            rec["cpu_baseline"] = {"spp_per_s": taken / ro.seconds, "msamples_per_s": so.camera_rays / ro.seconds / 1e6, "mrays_per_s": so.total_rays() / ro.seconds / 1e6,
                                   "seconds": ro.seconds, "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"the whole frame at BASELINE size and full settings ({st_o.width}x{st_o.height}, up to {st_o.samples_per_pixel} spp), rows dealt to threads"}
            rec["gpu_over_cpu_at_baseline_size"] = rec["baseline_size"]["spp_per_s"] / rec["cpu_baseline"]["spp_per_s"]
            log(f"[bench] {key}: GPU {rec['baseline_size']['ms']:.1f} ms, CPU {ro.seconds:.1f} s on {os.cpu_count()} threads")
        out[key] = rec
    return out


def copy_ceiling(torch, dist, world, dev, h2d_bytes, d2h_bytes, mb=256):
    """What the box's PCIe / host-memory path needs for one step's traffic with no kernels at all: `h2d_bytes` up and
    `d2h_bytes` down per rank as plain pinned copies (chunks of `mb` MB) running concurrently on two streams, all ranks at
    once. Returns the wall time of that (max over ranks) and the rates it implies, whole job."""
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(up, down):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_up):
            for off in range(0, up, n):
                c = min(n, up - off)
                d_in[:c].copy_(h_in[:c], non_blocking=True)
        with torch.cuda.stream(s_dn):
            for off in range(0, down, n):
                c = min(n, down - off)
                h_out[:c].copy_(d_out[:c], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    run(h2d_bytes, d2h_bytes)
    both = min(run(h2d_bytes, d2h_bytes) for _ in range(3))
    up = min(run(h2d_bytes, 0) for _ in range(2))
    down = min(run(0, d2h_bytes) for _ in range(2))
    return {"copy_only_ms_per_step": both * 1e3, "h2d_alone_gbs": world * h2d_bytes / up / 1e9, "d2h_alone_gbs": world * d2h_bytes / down / 1e9,
            "h2d_gbs_with_d2h_running": world * h2d_bytes / both / 1e9,
            "method": f"one step's bytes per rank as plain pinned copies in {mb} MB chunks, H2D and D2H on two streams, all {world} rank(s) at once, wall clock, max over ranks, best of 3"}


def time_other_scenes(crt, W, torch, stream, grid, with_cpu):
    """SURVEY 8(d)'s other kernel inputs, timed resident like the main step (best of 3, CUDA events): K2 — the stress scene
    with identical spheres, i.e. one 9 800-triangle BLAS x grid^2 instances over the flat shards/floor/slab (two-level
    traversal); K3 — the reference's three throughput-probe scenes (ray_throughput.rs:50-100) with its LCG box rays
    (:103-116). Beside each: the CPU restatement on one thread (the reference's own figure is per core) and on all, on a
    2^18-ray sample, with hit agreement on that sample."""
    import orc

    def gpu_rate(scene, rays, closest):
        n = len(rays)
        d_r = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
        d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda") if closest else torch.empty(n, dtype=torch.uint8, device="cuda")
        best = None
        for k in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if closest:
                scene.intersect_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), 0, stream.cuda_stream)
            else:
                scene.occluded_ptr(n, d_r.data_ptr(), 0, 0, d_o.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            if k:
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        return n / best / 1e3, d_o.cpu().numpy()

    out = {}
    cases = [("K2 instanced stress", lambda api: W.stress_scene_instanced(api, grid=grid).commit(), 24),
             ("K3 tri_spheres", lambda api: W.triangle_scene(api).commit(), 22),
             ("K3 sphere_grid", lambda api: W.sphere_grid_scene(api).commit(), 22),
             ("K3 instance_scene", lambda api: W.instance_scene(api).commit(), 22)]
    for name, make, log2 in cases:
        scene = make(crt)
        rays = W.ray_batch(1 << log2, W.scene_extent(scene.bounds()))
        pb = scene.primitive_breakdown()
        rec = {"rays": len(rays), "top_level_primitives": scene.primitive_count(), "instances": pb.instances}
        rec["closest_lcg_mrays"], hits = gpu_rate(scene, rays, True)
        rec["occluded_lcg_mrays"], occ = gpu_rate(scene, rays, False)
        if with_cpu:
            o = make(orc)
            sub = slice(0, None, len(rays) >> 18)
            oh, _ = o.intersect_n(rays[sub], threads=0)
            rec["cpu_all_threads_mrays"] = len(oh) / o.last_seconds / 1e6
            o.intersect_n(rays[sub][: 1 << 16], threads=1)
            rec["cpu_one_thread_mrays"] = (1 << 16) / o.last_seconds / 1e6
            g = hits.view(np.uint32).reshape(-1, 8)[sub]
            rec["oracle_agreement_on_sample"] = float(((g[:, 3] == oh["geom_id"]) & (g[:, 4] == oh["prim_id"]) & (g[:, 0].view(np.float32) == oh["t"])).mean())
            rec["occlusion_agreement_on_sample"] = float((occ[sub] == o.occluded_n(rays[sub], threads=0)).mean())
        out[name] = rec
        del scene
    return out


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def camera_dims(rays_log2):
    # 3840x2160 x 2 spp = 16.6 M ~ 2^24; smaller runs shrink the image, keeping 16:9 and 2 spp
    n = 1 << rays_log2
    h = max(9, int(round((n / 2 / (16 / 9)) ** 0.5)))
    if rays_log2 == 24:
        return 3840, 2160, 2
    return int(h * 16 / 9), h, 2


def build_ray_sets(W, scene_bounds, rays_log2, rank, trace_closest):
    """Returns dict name -> (rays RAY_DTYPE[n], mask or None). `trace_closest(rays)` gives hits (HIT_DTYPE)."""
    n = 1 << rays_log2
    ext = W.scene_extent(scene_bounds)
    sets = {}
    lcg = W.ray_batch(n, ext, seed=(0x2545F491 + 7919 * rank) & 0xFFFFFFFF)
    sets["closest_lcg"] = (lcg, None)
    sets["occluded_lcg"] = (lcg, None)
    w, h, spp = camera_dims(rays_log2)
    cam = W.stress_camera(w, h)
    parts = [cam.rays(spp, seed=rank, rows=(j, min(j + 256, h))) for j in range(0, h, 256)]
    cam_rays = np.concatenate(parts)
    sets["closest_camera"] = (cam_rays, None)
    hits = trace_closest(cam_rays)
    sets["closest_bounce"] = (W.bounce_rays(cam_rays, hits, seed=1 + rank), None)
    sr, sm = W.shadow_rays(cam_rays, hits, seed=2 + rank)
    sets["occluded_shadow"] = (sr, sm)
    return sets, {"lcg_extent": ext, "camera": [w, h, spp]}


def bytes_per_ray(stats, n, any_hit):
    """B_ray = 128 N_node + 48 N_tri + 32 N_sph + 48 N_curve + 64 N_inst + 32 (ray in) + B_out."""
    per = lambda k: stats[k] / n  # noqa: E731
    nodes = per("nodes_top") + per("nodes_inst")
    b = 128 * nodes + 48 * per("tri_lanes") + 32 * per("sphere_tests") + 48 * (per("curve_tests") + per("cubic_tests")) + 64 * per("instance_tests")
    return b + 32 + (1 if any_hit else 32), nodes, per("tri_lanes")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference arm: the CPU restatement of crust-rt on all host threads
# ---------------------------------------------------------------------------
def cpu_sample(sets, sample_log2):
    out = {}
    m = 1 << sample_log2
    for name in SETS:
        rays, mask = sets[name]
        stride = max(1, len(rays) // m)
        out[name] = (np.ascontiguousarray(rays[::stride][:m]), None if mask is None else np.ascontiguousarray(mask[::stride][:m]))
    return out


def run_cpu(o_scene, sample, threads, with_stats=False):
    """One pass of the five queries on the oracle. Returns (rays, seconds, per-set dict)."""
    total_rays, total_s, per = 0, 0.0, {}
    for name in SETS:
        rays, mask = sample[name]
        if name.startswith("closest"):
            res = o_scene.intersect_n(rays, None, mask, threads=threads, stats=with_stats)
        else:
            res = o_scene.occluded_n(rays, None, mask, threads=threads, stats=with_stats)
        s = o_scene.last_seconds
        per[name] = {"rays": len(rays), "seconds": s, "stats": res[-1] if with_stats else None}
        total_rays += len(rays)
        total_s += s
    return total_rays, total_s, per


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import orc
    from crust_render_b200 import workloads as W

    cores = os.cpu_count() or 1
    t0 = time.time()
    b, n_tris = W.stress_scene(orc, grid=args.grid, shards=30000, jitter=True)
    o_scene = b.commit(host_threads=0)
    log(f"[reference] oracle commit of {n_tris} triangles: {time.time() - t0:.1f}s")
    sets, info = build_ray_sets(W, o_scene.bounds(), args.rays_log2, 0, lambda r: o_scene.intersect_n(r, threads=0)[0])
    sample = cpu_sample(sets, args.cpu_sample_log2)
    for _ in range(args.warmup):
        run_cpu(o_scene, sample, 0)
    rays = secs = 0
    for _ in range(args.steps):
        r, s, _per = run_cpu(o_scene, sample, 0)
        rays, secs = rays + r, secs + s
    v = rays / secs / 1e6
    desc = f"{rays // args.steps} rays/step: 2^{args.cpu_sample_log2} strided rays of each of the 5 ray sets, {cores} threads, contiguous slices"
    line = {"metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(n_tris, args.rays_log2, sum(len(sets[n][0]) for n in SETS), info),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_render:
        del o_scene, sets, sample
        line["render"] = {"cpu_baseline": cpu_render_sample(args, "[reference]")}
        line["render"]["spp_per_s"] = line["render"]["cpu_baseline"]["value"]
    emit(line)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def main():
    args = parse()
    protect_stdout()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    import crust_render_b200 as crt
    from crust_render_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    crt.init(local)
    dev = torch.device("cuda", local)
    cores = os.cpu_count() or 1

    # ---- scene: replicated on every GPU; built once by rank 0 on all host cores and broadcast over NCCL ----
    from crust_render_b200.distributed import shared_commits
    share = shared_commits(dist if world > 1 else None, dev)
    share.__enter__()  # stays on for every commit of this process (scene, render worlds, sample configs)
    t0 = time.time()
    b, n_tris = W.stress_scene(crt, grid=args.grid, shards=30000, jitter=True)
    t_gen = time.time() - t0
    scene = b.commit(host_threads=cores)
    commit_s = time.time() - t0 - t_gen
    bi = scene.build_info()
    tc = torch.tensor([commit_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    commit_s_max = float(tc.item())
    if rank == 0:
        log(f"[bench] commit of {n_tris} triangles: {commit_s_max:.1f}s on the slowest rank (mesh generation {t_gen:.1f}s before it; host build {bi.build_seconds:.1f}s on rank 0 with {cores} threads, "
            f"upload {bi.upload_seconds:.2f}s" + (f", then one NCCL broadcast of {share.bytes / 1e6:.0f} MB to {world - 1} ranks" if world > 1 else "") +
            f"), {bi.wide_nodes} BVH4 nodes ({bi.node_bytes / 1e6:.0f} MB), {bi.leaf_refs} leaf refs ({bi.leaf_bytes / 1e6:.0f} MB)")

    t0 = time.time()
    sets, info = build_ray_sets(W, scene.bounds(), args.rays_log2, rank, lambda r: scene.intersect_n(r)[0])
    if rank == 0:
        log(f"[bench] ray sets: " + ", ".join(f"{k}={len(v[0])}" for k, v in sets.items()) + f" ({time.time() - t0:.1f}s)")

    # ---- device-resident inputs / outputs ------------------------------------
    d_in, d_mask, d_out, h_in, h_mask, h_out = {}, {}, {}, {}, {}, {}
    for name in SETS:
        rays, mask = sets[name]
        h = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).pin_memory()
        h_in[name] = h
        d_in[name] = h.to(dev, non_blocking=True)
        if mask is not None:
            hm = torch.from_numpy(mask.view(np.int32)).pin_memory()
            h_mask[name] = hm
            d_mask[name] = hm.to(dev, non_blocking=True)
        if name.startswith("closest"):
            d_out[name] = torch.empty((len(rays), 8), dtype=torch.float32, device=dev)
            h_out[name] = torch.empty((len(rays), 8), dtype=torch.float32).pin_memory()
        else:
            d_out[name] = torch.empty(len(rays), dtype=torch.uint8, device=dev)
            h_out[name] = torch.empty(len(rays), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    rays_per_step = sum(len(sets[n][0]) for n in SETS)
    stream = torch.cuda.current_stream()

    def launch(name):
        n = len(sets[name][0])
        mk = d_mask[name].data_ptr() if name in d_mask else 0
        if name.startswith("closest"):
            scene.intersect_ptr(n, d_in[name].data_ptr(), 0, mk, d_out[name].data_ptr(), 0, stream.cuda_stream)
        else:
            scene.occluded_ptr(n, d_in[name].data_ptr(), 0, mk, d_out[name].data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        for name in SETS:
            launch(name)
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream ------------
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in SETS] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = crt.launch_count()
    barrier()
    e_start.record(stream)
    torch.cuda.nvtx.range_push("crt_timed_steps")  # `ncu --nvtx --nvtx-include "crt_timed_steps/"` captures exactly these launches
    for k in range(args.steps):
        for j, name in enumerate(SETS):
            ev[k][j][0].record(stream)
            launch(name)
            ev[k][j][1].record(stream)
    torch.cuda.nvtx.range_pop()
    e_end.record(stream)
    barrier()
    launches = crt.launch_count() - launches0
    clk = clocks.stop()
    total_ms = e_start.elapsed_time(e_end)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = rays_per_step * world / (ms_per_step * 1e-3) / 1e6
    k_ms = {name: float(np.mean([ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(args.steps)])) for j, name in enumerate(SETS)}
    status = crt.device_status()
    # what the timed launches wrote, on the strided sample the CPU baseline runs: checked against the oracle below
    timed_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = 1 << args.cpu_sample_log2
        timed_sample = {name: d_out[name][:: max(1, len(sets[name][0]) // m)][:m].cpu().numpy() for name in SETS}

    # ---- what the kernel really touched: one untimed pass of the counting build over the same resident sets ----
    device_counts = None
    if rank == 0:
        device_counts = {}
        crt.trace_counters_read(reset=True)
        crt.trace_counters_enable(True)
        try:
            for name in SETS:
                launch(name)
                c = crt.trace_counters_read(reset=True)
                n_r = max(1, c["rays"])
                any_hit = not name.startswith("closest")
                device_counts[name] = {"nodes_fetched_per_ray": c["nodes_fetched"] / n_r, "entries_culled_on_pop_per_ray": c["entries_culled_on_pop"] / n_r,
                                       "leaves_per_ray": c["leaves_entered"] / n_r, "triangle_records_per_ray": c["triangle_records"] / n_r,
                                       "bytes_per_ray_touched": 128 * c["nodes_fetched"] / n_r + 48 * c["triangle_records"] / n_r + 32 + (1 if any_hit else 32)}
        finally:
            crt.trace_counters_enable(False)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            for name in SETS:
                n = len(sets[name][0])
                mk = h_mask[name].data_ptr() if name in h_mask else 0
                if name.startswith("closest"):
                    scene.intersect_host_ptr(n, h_in[name].data_ptr(), 0, mk, h_out[name].data_ptr(), 0)
                else:
                    scene.occluded_host_ptr(n, h_in[name].data_ptr(), 0, mk, h_out[name].data_ptr())
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(len(sets[n][0]) * 32 + (len(sets[n][0]) * 4 if n in h_mask else 0) for n in SETS)
        d2h = sum(len(sets[n][0]) * (32 if n.startswith("closest") else 1) for n in SETS)
        e2e = {"value": rays_per_step * world * args.steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt / args.steps * 1e3}
        # the ceiling of this box for that traffic: plain pinned copies in both directions at once, all ranks together
        ceil = copy_ceiling(torch, dist, world, dev, h2d, d2h)
        e2e["copy_ceiling"] = ceil
        e2e["h2d_gbs_achieved"] = world * h2d / (dt / args.steps) / 1e9
        e2e["frac_of_copy_ceiling"] = ceil["copy_only_ms_per_step"] / e2e["ms_per_step"]

    # ---- render: config 5 through the wavefront integrator, frame split over the ranks ----------
    render = None
    other_scenes = None
    sample_sets = cpu_sample(sets, args.cpu_sample_log2) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    full_sets_sizes = {n: len(sets[n][0]) for n in SETS}
    if not args.no_render:
        from crust_render_b200 import scenes
        from crust_render_b200.distributed import rank_settings, shared_host_frame

        del d_in, d_out, d_mask, h_in, h_out, h_mask
        torch.cuda.empty_cache()
        rw, rh, rspp = args.render_width, args.render_height, args.render_spp
        t0 = time.time()
        r_world, r_cam, r_st, r_tris = scenes.stress(crt, grid=args.grid, jitter=True, width=rw, height=rh, spp=rspp)
        if rank == 0:
            log(f"[bench] render world commit ({r_tris} triangles incl. the RectLight): {time.time() - t0:.1f}s")
        renderer = crt.Renderer(r_cam, r_world, r_st)
        fb = torch.zeros((rh, rw, 4), dtype=torch.float32, device=dev)
        # N > 1: peer-memory gather — every rank's film kernel stores its tiles into rank 0's IPC-mapped frame over NVLink
        # (no frame collective); the RayStats reduce is the only collective in the step
        r_gather = "peer" if world > 1 else "reduce"
        if world > 1:
            # peer access needs every rank to see the owner's GPU; where it does not (e.g. one visible device per process)
            # the frame is sum-reduced over NCCL instead, and `split` says which path ran
            from crust_render_b200.distributed import shared_frame
            ok = torch.ones(1, device=dev)
            try:
                shared_frame(dist, dev, rh, rw)
            except Exception as e:  # noqa: BLE001
                log(f"[bench] rank {rank}: peer-memory frame unavailable ({e}); using the NCCL reduce")
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) == 0.0:
                r_gather = "reduce"
        r_ms, st0, r_launches = timed_render(renderer, dist, world, fb, stream, args.render_steps, r_gather, torch, dev)
        dev_ms = renderer.device_ms
        if r_gather == "peer":
            from crust_render_b200.distributed import shared_frame
            r_frame = shared_frame(dist, dev, rh, rw)[1]
        else:
            r_frame = fb
        mean_rgb = [float(v) for v in r_frame[..., :3].double().mean(dim=(0, 1)).tolist()] if rank == 0 else None
        # e2e: the same frame through crt_render_host into ONE host frame. N = 1: pinned caller memory, the resolved frame is
        # copied straight into it. N > 1: the frame lives in POSIX shared memory mapped by every rank; each rank's
        # crt_render_host downloads only the pixels it owns (packed D2H over its own PCIe link, host-side scatter), so the
        # frame is complete in host memory when the last rank returns. Wall clock, barrier on both sides, max over ranks.
        if world > 1:
            host_fb = shared_host_frame(dist, rh, rw)
        else:
            host_fb = torch.empty((rh, rw, 4), dtype=torch.float32).pin_memory().numpy()
        renderer.settings = rank_settings(r_st, rank, world)
        renderer.render_with_stats(out=host_fb)  # warm-up: device frame / staging allocation, page faults of the host frame
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.render_steps):
            renderer.render_with_stats(out=host_fb)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.render_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        renderer.settings = r_st
        host_mean = [float(v) for v in host_fb[..., :3].mean(axis=(0, 1), dtype=np.float64)] if rank == 0 else None
        if rank == 0:
            render = {"workload": f"config 5: stress mesh ({r_tris} triangles) + 12x12 RectLight, {rw}x{rh}, depth 6, {rspp} spp per step, power MIS, triangle filter",
                      "spp_per_s": rspp / (r_ms * 1e-3), "msamples_per_s": rw * rh * rspp / r_ms / 1e3, "mrays_per_s": st0.total_rays() / r_ms / 1e3,
                      "ms_per_step": r_ms, "steps": args.render_steps, "rays_per_sample": st0.total_rays() / max(1, st0.camera_rays), "n_gpus": world,
                      "split": "whole frame" if world == 1 else (f"32x32 tiles dealt round-robin over {world} ranks; each rank's film kernel stores its tiles into rank 0's frame through CUDA-IPC peer memory (NVLink), RayStats reduced over NCCL, all inside the timed region"
                                if r_gather == "peer" else f"32x32 tiles dealt round-robin over {world} ranks, NCCL sum-reduce of the frame to rank 0 inside the timed region"),
                      "scaling": "strong", "gpu_launches_per_step_rank0": r_launches, "device_ms_rank0_render_only": dev_ms,
                      "e2e": {"spp_per_s": rspp / e2e_s, "unit": "spp/s", "ms_per_step": e2e_s * 1e3, "frac_of_device_spp_per_s": (rspp / e2e_s) / (rspp / (r_ms * 1e-3)),
                              "d2h_bytes_per_step": rw * rh * 16, "h2d_bytes_per_step": 0,
                              "note": "crt_render_host: every rank downloads only the pixels it owns into one host frame (pinned memory at N = 1, a POSIX shared-memory frame mapped by all ranks at N > 1); the frame is complete in host memory inside the timed region",
                              "mean_rgb_host_frame": host_mean},
                      "mean_rgb": mean_rgb, "ray_stats": st0.__dict__}
        del fb, renderer, r_world, host_fb
        torch.cuda.empty_cache()
        # configs 1-4 of BASELINE.json on the same N GPUs, each with the CPU restatement beside it (N = 1)
        cfgs = time_sample_configs(crt, torch, dist, world, rank, dev, stream, r_gather, with_cpu=(world == 1 and not args.no_cpu_baseline), steps=args.render_steps)
        if rank == 0:
            render["configs"] = cfgs
            if world == 1:
                other_scenes = time_other_scenes(crt, W, torch, stream, args.grid, not args.no_cpu_baseline)

    # ---- CPU baseline + algorithmic bytes per ray (rank 0, N = 1 only) -----------
    cpu_baseline, roofline, per_set = None, None, {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        stats_by_set = None
        if world == 1 and not args.no_cpu_baseline:
            import orc
            t0 = time.time()
            ob, _ = W.stress_scene(orc, grid=args.grid, shards=30000, jitter=True)
            o_scene = ob.commit(host_threads=0)
            log(f"[bench] oracle commit: {time.time() - t0:.1f}s")
            sample = sample_sets
            # parity check of the TIMED buffers (checker role only): what the timed launches left in d_out, on the strided
            # sample, against the oracle's answers for the same rays
            for name in SETS:
                rays, mask = sample[name]
                g = timed_sample[name]
                if name.startswith("closest"):
                    c = o_scene.intersect_n(rays, None, mask, threads=0)[0]
                    gu = g.view(np.uint32).reshape(-1, 8)
                    agree = float(((gu[:, 3] == c["geom_id"]) & (gu[:, 4] == c["prim_id"]) & (gu[:, 0].view(np.float32) == c["t"])).mean())
                else:
                    agree = float((g == o_scene.occluded_n(rays, None, mask, threads=0)).mean())
                per_set.setdefault(name, {})["oracle_agreement_of_timed_output"] = agree
            run_cpu(o_scene, sample, 0)
            r_all, s_all, _ = run_cpu(o_scene, sample, 0)
            r_1, s_1, _ = run_cpu(o_scene, {k: (v[0][: 1 << 17], None if v[1] is None else v[1][: 1 << 17]) for k, v in sample.items()}, 1)
            _, _, st = run_cpu(o_scene, {k: (v[0][: 1 << 17], None if v[1] is None else v[1][: 1 << 17]) for k, v in sample.items()}, 0, with_stats=True)
            stats_by_set = {k: (v["stats"], v["rays"]) for k, v in st.items()}
            cpu_baseline = {"value": r_all / s_all / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                            "sample": f"2^{args.cpu_sample_log2} strided rays of each of the 5 sets ({r_all} rays, {s_all:.1f}s of CPU work); "
                                      f"C++ restatement of crust-rt, one ray per call, contiguous slices per thread",
                            "one_thread_mrays": r_1 / s_1 / 1e6}
        # roofline of the dominant kernel (closest-hit): algorithmic bytes / measured duration
        closest = [n for n in SETS if n.startswith("closest")]
        if stats_by_set:
            tot_bytes = 0.0
            for name in SETS:
                bpr, nodes, tris = bytes_per_ray(stats_by_set[name][0], stats_by_set[name][1], not name.startswith("closest"))
                per_set.setdefault(name, {}).update({"bytes_per_ray": bpr, "nodes_per_ray": nodes, "tri_lanes_per_ray": tris})
            for name in closest:
                tot_bytes += per_set[name]["bytes_per_ray"] * len(sets[name][0])
            dur = sum(k_ms[n] for n in closest) * 1e-3
            ach = tot_bytes / dur / 1e9
            traffic, traffic_src = None, None
            try:  # DRAM bytes per closest-hit launch from the committed `ncu --set full` capture of this command — quoted only
                # while the traversal sources are the ones it was captured on (tools/trace_traffic.py stamps their hash)
                from tools.trace_traffic import source_sha
                tj = json.load(open(os.path.join(ROOT, "profiles", "trace_traffic.json")))
                if tj.get("source_sha") == source_sha():
                    traffic, traffic_src = float(tj["mean_traffic_per_closest_launch"]), tj["source"]
                else:
                    traffic_src = "profiles/trace_traffic.json was captured on other traversal sources (hash mismatch): not quoted"
            except (OSError, KeyError, ValueError, ImportError):
                pass
            roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "note": "cache-oblivious figure (SURVEY 8(d)): the algorithmic bytes are mostly served by L1/L2 (the 0.83 GB tree is hot in the 126 MB L2); "
                                "ncu (profiles/trace_traffic.json) shows the kernel issue- and divergence-bound, DRAM at a few percent of peak",
                        "kernel": "trace_kernel<closest-hit, triangles>", "peak_source": peak_src,
                        "bytes_per_launch": tot_bytes / len(closest), "launch_ms": dur * 1e3 / len(closest)}
        for name in SETS:
            per_set.setdefault(name, {}).update({"rays": len(sets[name][0]), "kernel_ms": k_ms[name], "mrays": len(sets[name][0]) / k_ms[name] / 1e3})
            if device_counts:
                per_set[name]["device_counters"] = device_counts[name]
        if roofline is not None and device_counts:
            # the same roofline with the bytes the device kernel really touched (it culls popped entries without fetching
            # them, so it touches fewer nodes than the reference algorithm the contract's figure describes)
            touched = sum(device_counts[n]["bytes_per_ray_touched"] * len(sets[n][0]) for n in closest)
            roofline["achieved_touched_bytes"] = touched / (sum(k_ms[n] for n in closest) * 1e-3) / 1e9
            roofline["frac_touched_bytes"] = roofline["achieved_touched_bytes"] / peak

        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(n_tris, args.rays_log2, rays_per_step, info),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "sets": per_set, "closest_hit_mrays": sum(len(sets[n][0]) for n in closest) / sum(k_ms[n] for n in closest) / 1e3,
                "occlusion_mrays": sum(len(sets[n][0]) for n in SETS if n not in closest) / sum(k_ms[n] for n in SETS if n not in closest) / 1e3,
                "device_status": status, "host_cores": cores,
                "commit_seconds_slowest_rank": commit_s_max,
                "commit_sharing": ("one process" if world == 1 else f"rank 0 builds with {cores} host threads, crt_scene_export -> NCCL broadcast -> crt_scene_import on the other {world - 1} ranks")}
        if other_scenes is not None:
            line["other_scenes"] = other_scenes
        if world == 1:
            # commit() itself: the reference's SAH/SBVH host build (what the numbers above ran on) next to the device build
            def timed_commit(flags):
                bb, _ = W.stress_scene(crt, grid=args.grid, shards=30000, jitter=True)
                t0 = time.perf_counter()
                sc = bb.commit(flags=flags, host_threads=cores)
                return sc, time.perf_counter() - t0
            timed_commit(crt.COMMIT_DEVICE_BUILD)  # warm-up: scratch arenas, pinned staging
            dev_scene, dev_s = timed_commit(crt.COMMIT_DEVICE_BUILD)
            dbi = dev_scene.build_info()
            dev_rates = {}
            for name in SETS:
                rays, mask = sample_sets[name] if sample_sets else (sets[name][0][: 1 << 22], None if sets[name][1] is None else sets[name][1][: 1 << 22])
                dr = torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).to(dev)
                dm = torch.from_numpy(np.ascontiguousarray(mask).view(np.int32)).to(dev) if mask is not None else None
                closest_k = name.startswith("closest")
                do = torch.empty((len(rays), 8), dtype=torch.float32, device=dev) if closest_k else torch.empty(len(rays), dtype=torch.uint8, device=dev)
                rates = {}
                for tag, sc in (("reference_tree", scene), ("device_tree", dev_scene)):
                    best = None
                    for k in range(4):
                        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ea.record(stream)
                        mk = dm.data_ptr() if dm is not None else 0
                        if closest_k:
                            sc.intersect_ptr(len(rays), dr.data_ptr(), 0, mk, do.data_ptr(), 0, stream.cuda_stream)
                        else:
                            sc.occluded_ptr(len(rays), dr.data_ptr(), 0, mk, do.data_ptr(), stream.cuda_stream)
                        eb.record(stream)
                        torch.cuda.synchronize()
                        if k:
                            best = ea.elapsed_time(eb) if best is None else min(best, ea.elapsed_time(eb))
                    rates[tag] = len(rays) / best / 1e3
                dev_rates[name] = rates
            line["commit"] = {"triangles": n_tris, "host_sbvh_build_s": bi.build_seconds, "host_upload_s": bi.upload_seconds, "host_threads": cores,
                              "device_build_commit_s": dev_s, "device_build_kernels_s": dbi.build_seconds, "device_concat_upload_s": dbi.upload_seconds,
                              "device_tree": {"wide_nodes": int(dbi.wide_nodes), "leaves": int(dbi.leaves), "leaf_refs": int(dbi.leaf_refs)},
                              "reference_tree": {"wide_nodes": int(bi.wide_nodes), "leaves": int(bi.leaves), "leaf_refs": int(bi.leaf_refs)},
                              "mrays_on_each_tree": dev_rates,
                              "note": "CRT_COMMIT_DEVICE_BUILD (csrc/device_build.cu) is opt-in; every other number in this line uses the reference-identical host tree"}
            del dev_scene
        if render is not None:
            if world == 1 and not args.no_cpu_baseline:
                del o_scene
                render["cpu_baseline"] = cpu_render_sample(args, "[bench]")
            line["render"] = render
            line["spp_per_s"] = render["spp_per_s"]
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
