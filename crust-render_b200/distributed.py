"""Multi-GPU rendering: image-space split with a replicated scene.

The reference shards a frame over rayon workers by pixel / 16x16 tile with no shared mutable state
(crates/crust-core/src/tracer.rs:423-502). Here the workers are processes, one per GPU: square tiles
are dealt round-robin (tile t -> rank t % world, so adaptive-sampling cost is decorrelated from
rank), every rank renders its tiles into a zero-initialised full-frame buffer, and one collective —
a sum-reduce to rank 0 over NCCL/NVLink, exact because pixel ownership is disjoint — assembles the
frame. Per-pixel sampler streams depend only on (i, j, frame, sample), so the image is bit-identical
for every world size. Traversal and shading never communicate.
"""
from __future__ import annotations

import copy
import dataclasses

import numpy as np

from .render_types import RayStats, STATS_DTYPE


def rank_settings(settings, rank: int, world: int, tile_size: int = 32):
    s = copy.copy(settings)
    s.tile_size, s.tile_rank, s.tile_world = tile_size, rank, world
    return s


def owned_mask(width: int, height: int, tile_size: int, rank: int, world: int) -> np.ndarray:
    """[h, w] bool: pixels whose tile id % world == rank (the split crt_render applies)."""
    if world <= 1:
        return np.ones((height, width), dtype=bool)
    tx = (width + tile_size - 1) // tile_size
    jj, ii = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    return ((jj // tile_size) * tx + ii // tile_size) % world == rank


_CURSORS = {}


def shared_cursor(dist, device):
    """One StealCursor per process group: rank 0 creates it, the 64-byte CUDA IPC handle is broadcast, the other
    ranks open it (peer access over NVLink). Cached for the life of the process."""
    import torch

    from .render import StealCursor

    key = id(dist.group.WORLD) if dist is not None and dist.is_initialized() else None
    if key in _CURSORS:
        return _CURSORS[key]
    world = dist.get_world_size() if key is not None else 1
    rank = dist.get_rank() if world > 1 else 0
    cur = StealCursor() if rank == 0 else None
    if world > 1:
        t = torch.zeros(64, dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cur.handle), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        if rank != 0:
            cur = StealCursor(bytes(t.cpu().numpy().tobytes()))
    _CURSORS[key] = cur
    return cur


_FRAMES = {}


def shared_frame(dist, device, height: int, width: int):
    """The peer-mapped final frame of this process group (render.SharedFrame): rank 0 owns it, the others map it through
    its CUDA IPC handle. Returns (SharedFrame, torch view of it). Cached per (group, size)."""
    import torch

    from .render import SharedFrame

    gkey = id(dist.group.WORLD) if dist is not None and dist.is_initialized() else None
    key = (gkey, height, width)
    if key in _FRAMES:
        return _FRAMES[key]
    world = dist.get_world_size() if gkey is not None else 1
    rank = dist.get_rank() if world > 1 else 0
    frame, err = None, None
    if rank == 0:
        try:
            frame = SharedFrame(height, width)
        except Exception as e:  # noqa: BLE001  (the peers must still get their broadcast)
            err = e
    if world > 1:
        t = torch.zeros(64, dtype=torch.uint8, device=device)
        if frame is not None:
            t.copy_(torch.frombuffer(bytearray(frame.handle), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        handle = bytes(t.cpu().numpy().tobytes())
        if rank != 0:
            if not any(handle):
                raise RuntimeError("the owning rank could not create the shared frame")
            frame = SharedFrame(height, width, handle)
    if err is not None:
        raise err
    view = torch.as_tensor(frame, device=device)
    _FRAMES[key] = (frame, view)
    return _FRAMES[key]


class shared_commits:
    """Build once, share: inside this context every `SceneBuilder.commit()` of the product binding is performed by rank 0
    alone — with all the host's cores instead of cores / world — and the committed scene (flattened BVH4 nodes, leaf
    records, id table; crt_scene_export) is broadcast to the other ranks over NCCL as one device buffer and unpacked there
    with device-to-device copies (crt_scene_import). The other ranks discard their builders without building. All ranks must
    commit the same scenes in the same order (SPMD), which scenes.py / the importer do. The reference builds once per
    process (scene.rs:226); one process per GPU would otherwise repeat the SAH/SBVH build N times on 1/N of the cores."""

    def __init__(self, dist, device):
        self.dist, self.device = dist, device
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.commits, self.bytes = 0, 0

    def _commit(self, builder, flags, host_threads):
        import os

        import torch

        from . import rt

        if self.world <= 1 or (flags & rt.COMMIT_HOST_ONLY):
            return builder.commit_local(flags, host_threads)
        children = builder.distinct_children()
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        scene = None
        if self.rank == 0:
            scene = builder.commit_local(flags, os.cpu_count() or 1)
            n[0] = scene.export_size()
        else:
            builder.discard()
        self.dist.broadcast(n, src=0)
        nbytes = int(n.item())
        blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            scene.export_to(blob.data_ptr(), nbytes, True, torch.cuda.current_stream().cuda_stream)
        self.dist.broadcast(blob, src=0)
        if self.rank != 0:
            torch.cuda.current_stream().synchronize()
            scene = rt.Scene.import_from(blob.data_ptr(), nbytes, True, children)
        self.commits += 1
        self.bytes += nbytes
        return scene

    def __enter__(self):
        from . import rt
        self._prev = rt._commit_override
        rt._commit_override = self._commit
        return self

    def __exit__(self, *exc):
        from . import rt
        rt._commit_override = self._prev
        return False


_HOST_FRAMES = {}


def shared_host_frame(dist, height: int, width: int) -> np.ndarray:
    """One [h, w, 4] float32 frame in POSIX shared memory mapped by every rank of the (single-node) process group: rank 0
    creates it, the name is broadcast, the others attach. Each rank's crt_render_host then downloads the pixels it owns
    straight into this frame (its own PCIe link, its own host threads), so the whole frame is assembled in host memory
    without any cross-rank copy. Cached per (group, size); unlinked when the creating process exits."""
    import atexit
    from multiprocessing import shared_memory

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    key = (id(dist.group.WORLD) if world > 1 else None, height, width)
    if key in _HOST_FRAMES:
        return _HOST_FRAMES[key][1]
    nbytes = height * width * 16
    name = [None]
    shm = None
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        name[0] = shm.name
        atexit.register(lambda s=shm: (s.close(), s.unlink()))
    if world > 1:
        dist.broadcast_object_list(name, src=0)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name[0])
        try:  # the creator unlinks; attaching processes must not (Python < 3.13 registers every attachment for cleanup)
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:  # noqa: BLE001
            pass
        atexit.register(shm.close)
    frame = np.ndarray((height, width, 4), dtype=np.float32, buffer=shm.buf)
    if rank == 0:
        frame.fill(0.0)
    if world > 1:
        dist.barrier()
    _HOST_FRAMES[key] = (shm, frame)
    return frame


def render_distributed(renderer, dist=None, tile_size: int = 32, fb=None, stream: int = 0, steal: bool = False, gather: str = "reduce"):
    """Renders this rank's tiles and sum-reduces frame + RayStats to rank 0.

    steal=True (device framebuffer only): instead of the static round-robin deal, every rank pulls chunks of
    `settings.steal_chunk_tiles` tiles off a cursor shared through CUDA IPC until the frame is exhausted —
    balances adaptive sampling, whose cost per tile is not known in advance. The image is the same either way.

    `renderer` is a Renderer (CUDA binding; the oracle's in CPU tests). With `fb` a CUDA torch tensor
    [h, w, 4] the render writes device memory and the reduce runs on it (NCCL); otherwise host
    buffers are used (gloo). Returns (frame, RayStats) — complete on rank 0, this rank's share elsewhere.

    gather="peer" (device path): no per-rank frame and no frame collective at all — every rank's film kernel stores
    its tiles directly into rank 0's frame, which the ranks map through CUDA IPC (NVLink / NVSwitch peer stores;
    `shared_frame`). The frame is complete on rank 0 when the RayStats reduce (the only collective left) returns.
    `fb` then only gives the device and the size; the returned frame is the shared one.
    """
    import torch

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    base = renderer.settings
    renderer.settings = rank_settings(base, rank, world, tile_size)
    peer = gather == "peer" and fb is not None and world > 1
    if peer:
        _, fb = shared_frame(dist, fb.device, base.height, base.width)
    try:
        if peer:
            # tiles are disjoint and every pixel of the frame belongs to some rank: nothing to zero, nothing to reduce.
            # The barrier keeps a fast rank from overwriting the previous frame while rank 0's consumer still reads it.
            dist.barrier()
            if steal:
                cursor = shared_cursor(dist, fb.device)
                if rank == 0:
                    cursor.reset()
                dist.barrier()
                stats = renderer.render_steal(fb.data_ptr(), cursor, stream)
            else:
                stats = renderer.render_device(fb.data_ptr(), stream)
            frame = fb
        elif fb is not None and steal:
            cursor = shared_cursor(dist, fb.device)
            fb.zero_()
            if world > 1:
                dist.barrier()          # nobody is still pulling from the previous frame
            if rank == 0:
                cursor.reset()
            if world > 1:
                dist.barrier()          # the cursor is zero before anyone pulls
            stats = renderer.render_steal(fb.data_ptr(), cursor, stream)
            frame = fb
        elif fb is not None:
            fb.zero_()
            stats = renderer.render_device(fb.data_ptr(), stream)
            frame = fb
        else:
            host, stats = renderer.render_with_stats()
            frame = torch.from_numpy(host)
    finally:
        renderer.settings = base
    if world > 1:
        if not peer:
            dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
        st = torch.tensor([getattr(stats, n) for n in STATS_DTYPE.names], dtype=torch.int64, device=frame.device)
        dist.reduce(st, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            stats = RayStats(*[int(v) for v in st.tolist()])
    return (frame if fb is not None else frame.numpy()), stats
