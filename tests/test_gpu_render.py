"""-m gpu: the CUDA wavefront integrator (csrc/render.cu, through the C ABI of include/crt_render.h)
against the CPU oracle (oracle/crt_oracle_render.hpp) on identical scenes, settings and sampler
seeds. BASELINE.json's bar is "at equal spp and identical Sobol seeds, images match within a stated
relative-MSE tolerance": the tolerance stated here is relMSE <= 1e-6 with <= 1% of pixels off by more
than 1e-3 relative (measured on B200: relMSE 4e-12 .. 1.2e-7, see profiles/r01_render_parity.jsonl).
Differences come from CUDA's vs glibc's transcendental functions and the forward-vs-backward
summation order; a handful of paths per million flip a discrete decision (lobe pick, roulette)."""
import math

import numpy as np
import pytest

from crust_render_b200 import scenes
from crust_render_b200.distributed import owned_mask, rank_settings
from crust_render_b200.render_types import PixelFilter, PixelFilterKind, SamplingStrategy

pytestmark = pytest.mark.gpu

RELMSE_TOL = 1e-6


def compare(a, b, frac_tol=0.01):
    a3, b3 = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    relmse = float(((a3 - b3) ** 2 / (b3 ** 2 + 1e-4)).mean())
    rel = (np.abs(a3 - b3) / (np.abs(b3) + 1e-3)).max(axis=-1)
    assert np.all(np.isfinite(a)), "NaN/Inf in the GPU frame"
    assert relmse <= RELMSE_TOL, relmse
    assert (rel > 1e-3).mean() <= frac_tol, (rel > 1e-3).mean()
    return relmse


def stats_close(g, c, tol=1e-4, adaptive=False):
    if not adaptive:  # with adaptive stopping a 1-ulp luminance difference can move one pixel's stop by 4 samples
        assert g.camera_rays == c.camera_rays
    for name in ("camera_rays", "closest_hit", "shadow_rays", "vertices", "rr_tested", "rr_killed", "ended_escaped", "ended_depth"):
        a, b = getattr(g, name), getattr(c, name)
        assert abs(a - b) <= max(4, tol * max(a, b)), (name, a, b)


def both(gpu, orc, fn, **kw):
    wg, cam, st = fn(gpu, **kw)[:3]
    wo = fn(orc, **kw)[0]
    rg = gpu.Renderer(cam, wg, st)
    fg, sg = rg.render_with_stats()
    fo, so = orc.Renderer(cam, wo, st).render_with_stats()
    return fg, sg, fo, so, rg


def test_every_integrator_branch_scene(gpu, orc):
    """Diffuse, metal, coat, thick glass with an absorbing medium, thin-walled delta lobe, fuzz on
    instanced meshes, sphere + rect lights, lens sampling."""
    fg, sg, fo, so, rg = both(gpu, orc, scenes.furnace_room, width=192, height=128, spp=16, variance_threshold=0.0)
    compare(fg, fo)
    stats_close(sg, so)
    assert np.array_equal(fg[..., 3], fo[..., 3]) and rg.launches > 0 and gpu.device_status() == 0


def test_adaptive_sampling_decisions_match(gpu, orc):
    fg, sg, fo, so, rg = both(gpu, orc, scenes.furnace_room, width=192, height=128, spp=64, variance_threshold=0.05, min_samples_per_pixel=16)
    assert (fg[..., 3] == fo[..., 3]).mean() >= 0.999  # a 1-ulp luminance difference may move a threshold decision
    same = fg[..., 3] == fo[..., 3]
    compare(fg[same], fo[same])
    assert rg.rounds > 1 and fg[..., 3].min() >= 16 and (fg[..., 3] < 64).any()
    stats_close(sg, so, tol=2e-3, adaptive=True)


def test_config2_veach_mis(gpu, orc):
    fg, sg, fo, so, _ = both(gpu, orc, scenes.veach_mis, width=240, height=135, spp=16)
    compare(fg, fo)
    stats_close(sg, so)


@pytest.mark.parametrize("strategy", list(SamplingStrategy))
def test_sampling_strategies(gpu, orc, strategy):
    fg, sg, fo, so, _ = both(gpu, orc, scenes.veach_mis, width=96, height=54, spp=8, sampling_strategy=strategy)
    compare(fg, fo)
    stats_close(sg, so)
    assert (sg.shadow_rays == 0) == (strategy == SamplingStrategy.BsdfOnly)


def test_config3_openpbr_showcase_adaptive(gpu, orc):
    fg, sg, fo, so, _ = both(gpu, orc, scenes.openpbr_showcase, width=320, height=180, spp=64)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)


def test_config5_stress_scene_small(gpu, orc):
    fn = lambda api, **kw: scenes.stress(api, **kw)  # noqa: E731
    fg, sg, fo, so, _ = both(gpu, orc, fn, grid=2, shards=2000, width=160, height=90, spp=8)
    compare(fg, fo)
    stats_close(sg, so)


def test_config1_cornellbox(gpu, orc):
    """samples/cornellbox.usda via the importer's fixture: 822 baked triangles + two instances of one
    760-triangle BLAS, no lights (no shadow rays at all), depth 32 with Russian roulette, adaptive on."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.cornellbox, width=128, height=128, spp=48)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)
    assert sg.shadow_rays == 0 and sg.rr_tested > 0


def test_config4_point_instanced_medcity(gpu, orc):
    """samples/PointInstancedMedCity.usd via the crate reader's fixture: one TLAS of 40 000 instances over 8
    prototype BLASes (two-level traversal in one kernel), sky light only, importer defaults (adaptive, depth 32)."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.medcity, width=192, height=108, spp=40)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)
    assert sg.shadow_rays == 0 and sg.rr_tested > 0


# ---- BASELINE.json configs 1-3 at the sizes the config list names (the reduced-size tests above run first and fail faster) ----
def test_config1_cornellbox_at_baseline_size(gpu, orc):
    """512x512, 64 spp, adaptive (threshold 0.05, min 32), depth 32: the full config, every pixel against the oracle."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.cornellbox, width=512, height=512, spp=64)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)


def test_config2_veach_mis_at_baseline_size(gpu, orc):
    """960x540, 128 spp, balance heuristic, depth 8, adaptive off (samples/veach_mis.usda:206-216)."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.veach_mis, width=960, height=540, spp=128)
    compare(fg, fo)
    stats_close(sg, so)
    assert np.array_equal(fg[..., 3], fo[..., 3])


def test_config3_openpbr_showcase_at_baseline_size(gpu, orc):
    """640x360, 128 spp, adaptive, depth 32: every OpenPBR lobe of the sample."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.openpbr_showcase, width=640, height=360, spp=128)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)


def test_config4_medcity_at_baseline_size(gpu, orc):
    """640x360, up to 128 spp adaptive, depth 32, 40 000 instances: the full config (importer defaults), every pixel."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.medcity, width=640, height=360, spp=128)
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.999
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=2e-3, adaptive=True)


def test_dome_light_with_environment_map_and_distant_light(gpu, orc):
    """samples/domelight.usda: an importance-sampled lat-long HDRI dome (rotated) plus a 1.5-degree
    distant light; both are found by NEE and by escaping bounce rays with matching MIS densities."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.domelight, width=160, height=90, spp=16, variance_threshold=0.0)
    compare(fg, fo)
    stats_close(sg, so)
    assert sg.shadow_rays == sg.vertices and sg.ended_escaped > 0


# ---- volume regions: shade_kernel<VOL> / nee_resolve_kernel<VOL> against the oracle ----------------
@pytest.mark.parametrize("kind", ["absorb", "emit", "furnace", "furnace_grid", "furnace_noise"])
def test_volume_slab_probes(gpu, orc, kind):
    fg, sg, fo, so, _ = both(gpu, orc, scenes.volume_slab, kind=kind, spp=16)
    compare(fg, fo)
    stats_close(sg, so)
    if kind.startswith("furnace"):
        assert abs(float(fg[..., :3].mean()) - 1.0) < 0.01  # the analytic answer, on the GPU itself
    assert gpu.device_status() == 0


def test_fog_sample_scene(gpu, orc):
    """samples/fog.usda: homogeneous fog, volume NEE through the analytic transmittance fast path, adaptive rounds."""
    fg, sg, fo, so, rg = both(gpu, orc, scenes.fog, variance_threshold=0.0, samples_per_pixel=32)
    compare(fg, fo)
    stats_close(sg, so)
    fg, sg, fo, so, rg = both(gpu, orc, scenes.fog)  # as authored: adaptive
    same = fg[..., 3] == fo[..., 3]
    assert same.mean() >= 0.998 and rg.rounds > 1
    compare(fg[same], fo[same])
    stats_close(sg, so, tol=4e-3, adaptive=True)


def test_smoke_sample_scene(gpu, orc):
    """samples/smoke.usda: fBm smoke (delta + ratio tracking), an emissive ember, a voxel grid, overlapping majorants."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.smoke, variance_threshold=0.0, samples_per_pixel=16)
    compare(fg, fo)
    stats_close(sg, so)


def test_volumes_can_be_replaced_and_cleared(gpu, orc):
    wg, cam, st = scenes.volume_slab(gpu, "furnace", spp=4)
    a, _ = gpu.Renderer(cam, wg, st).render_with_stats()
    wg.with_volumes([])
    b, sb = gpu.Renderer(cam, wg, st).render_with_stats()
    assert np.all(b[..., :3] == 1.0) and sb.vertices == 0  # the bare dome
    wg.with_volumes(scenes.volume_slab(orc, "furnace", spp=4)[0].volumes)
    c, _ = gpu.Renderer(cam, wg, st).render_with_stats()
    assert a.tobytes() == c.tobytes()


def test_any_number_of_volume_regions(gpu, orc):
    """`Volumes` holds a Vec of regions with no cap (volume.rs:356-366). 48 regions tiled side by side under the 1-degree
    camera of the slab probe (a ray crosses one or two of them) render like the oracle and like one wide slab; the one limit
    left is per ray — a segment that crosses more than CRT_MAX_VOLUMES regions at once walks the first 16 and sets bit 1 of
    the device status word."""
    from crust_render_b200.render_types import DensityField, VolumeRegion

    def tiles(n):
        out = []
        for k in range(n):
            m = np.eye(4, dtype=np.float32)
            m[0, 3] = -12.0 + 0.5 * k + 0.25  # 0.5-wide columns next to each other along x (column-vector l2w)
            out.append(VolumeRegion(m, (0.25, 50.0, 0.5), (0,) * 3, (0.7,) * 3, 0.0, (0,) * 3, 1.0, DensityField.Homogeneous()))
        return out

    wg, cam, st = scenes.volume_slab(gpu, "absorb", spp=16)
    wo = scenes.volume_slab(orc, "absorb", spp=16)[0]
    one, _ = gpu.Renderer(cam, wg, st).render_with_stats()
    wg.with_volumes(tiles(48))
    wo.with_volumes(tiles(48))
    many, sm = gpu.Renderer(cam, wg, st).render_with_stats()
    ref, so = orc.Renderer(cam, wo, st).render_with_stats()
    compare(many, ref)
    stats_close(sm, so)
    # the tracker's transmittance estimate is stochastic (weights of 0 or 1 per walk): 16 Ki walks => sigma 0.004
    assert abs(float(many[..., :3].mean()) - math.exp(-0.7)) < 0.016 and abs(float(one[..., :3].mean()) - math.exp(-0.7)) < 0.016
    assert gpu.device_status() == 0
    wg.with_volumes(scenes.volume_slab(orc, "furnace", spp=4)[0].volumes * 17)  # 17 coincident regions: every ray crosses them all
    gpu.Renderer(cam, wg, st).render_with_stats()
    assert gpu.device_status() & 2
    gpu.device_status_reset()
    assert gpu.device_status() == 0


@pytest.mark.parametrize("name", ["instancing", "nested_instancing", "curves"])
def test_instancing_and_curve_samples(gpu, orc, name):
    """samples/instancing.usda (native instances of a two-material class prototype + a PointInstancer), nested_instancing.usda
    (three-level descent: instancer -> nested instancer sub-scene -> leaf BLAS) and curves.usda (round cones + cubic Bezier spans)."""
    fg, sg, fo, so, _ = both(gpu, orc, scenes.sample_stage, name=name, width=160, height=90, samples_per_pixel=16, variance_threshold=0.0)
    compare(fg, fo)
    stats_close(sg, so)
    assert gpu.device_status() == 0


def test_synthetic_instanced_city(gpu, orc):
    """The scalable stand-in generator (any instance count) used for instancing stress runs."""
    fn = lambda api, **kw: scenes.point_instanced_city(api, **kw)  # noqa: E731
    fg, sg, fo, so, _ = both(gpu, orc, fn, n_instances=2500, width=160, height=90, spp=8, variance_threshold=0.0)
    compare(fg, fo)
    stats_close(sg, so)


@pytest.mark.parametrize("kind,radius", [(PixelFilterKind.Box, 0.5), (PixelFilterKind.Gaussian, 1.5), (PixelFilterKind.Mitchell, 2.0)])
def test_pixel_filters(gpu, orc, kind, radius):
    fg, _, fo, _, _ = both(gpu, orc, scenes.furnace_room, width=64, height=48, spp=8, variance_threshold=0.0, pixel_filter=PixelFilter(kind, radius))
    compare(fg, fo)


def test_motion_blur_and_depth_limits(gpu, orc):
    """A moving instance makes has_motion() true, so the shutter dimension is drawn (tracer.rs:556,
    579-583); max_depth 1 and 2 exercise the depth-exhausted emission-only trace (:1123-1148)."""
    from crust_render_b200.render_types import Camera, LightList, OpenPBR, RenderSettings
    from crust_render_b200.rt_types import Affine3A, Geometry
    from crust_render_b200 import workloads as W

    def build(api, depth):
        wb = api.WorldBuilder()
        lights = LightList()
        wb.attach(Geometry.TriangleMesh(np.array([(-6, 0, -6), (6, 0, -6), (6, 0, 6), (-6, 0, 6)], np.float32), np.array([[0, 2, 1], [0, 3, 2]], np.uint32)), OpenPBR())
        sv, si = W.uv_sphere((0, 0, 0), 0.7, 12, 6)
        sb = api.SceneBuilder()
        sb.attach(Geometry.TriangleMesh(sv, si))
        child = sb.commit()
        wb.attach(Geometry.Instance(child, Affine3A.from_translation((-1, 0.7, 0)), Affine3A.from_translation((1.5, 0.7, 0))), OpenPBR(base_color=(0.9, 0.2, 0.2)))
        scenes._sphere_light(wb, lights, (0, 4, 1), 0.6, (1, 1, 1), 25.0, True)
        return wb.commit(lights), Camera((0, 2, 6), (0, 0.7, 0), (0, 1, 0), 40.0, 1.5), RenderSettings(8, depth, 96, 64, variance_threshold=0.0)

    for depth in (1, 2, 5):
        wg, cam, st = build(gpu, depth)
        wo, _, _ = build(orc, depth)
        assert wg.has_motion() and wo.has_motion()
        fg, sg = gpu.Renderer(cam, wg, st).render_with_stats()
        fo, so = orc.Renderer(cam, wo, st).render_with_stats()
        compare(fg, fo)
        stats_close(sg, so)
        assert sg.ended_depth > 0


def test_max_depth_zero_and_empty_world(gpu, orc):
    from crust_render_b200.render_types import Camera, RenderSettings
    wg, cam, st = scenes.furnace_room(gpu, width=32, height=16, spp=2, max_depth=0, variance_threshold=0.0)
    fg, sg = gpu.Renderer(cam, wg, st).render_with_stats()
    assert np.all(fg[..., :3] == 0) and np.all(fg[..., 3] == 2) and sg.ended_depth == sg.camera_rays == 32 * 16 * 2 and sg.closest_hit == 0
    for api in (gpu, orc):  # nothing attached: every ray sees the sky gradient (tracer.rs:1330-1337)
        w = api.WorldBuilder().commit()
        f, s = api.Renderer(Camera((0, 0, 0), (0, 0, -1), (0, 1, 0), 60.0, 2.0), w, RenderSettings(2, 4, 32, 16, variance_threshold=0.0)).render_with_stats()
        assert s.ended_escaped == s.camera_rays == s.closest_hit and s.shadow_rays == 0
        if api is gpu:
            sky = f
    compare(sky, f)


def test_bit_determinism_pool_size_and_tile_split(gpu):
    """GPU vs GPU, bit for bit: repeated renders, a path pool so small that every round is cut into
    many wavefronts, and the multi-GPU tile deal (each rank's tiles == the same pixels of the full frame)."""
    world, cam, st = scenes.furnace_room(gpu, width=96, height=64, spp=12, variance_threshold=0.05, min_samples_per_pixel=4)
    full, sf = gpu.Renderer(cam, world, st).render_with_stats()
    again, sa = gpu.Renderer(cam, world, st).render_with_stats()
    assert full.tobytes() == again.tobytes() and sf == sa
    st_small = scenes.furnace_room(gpu, width=96, height=64, spp=12, variance_threshold=0.05, min_samples_per_pixel=4)[2]
    st_small.max_paths = 1500
    small, ss = gpu.Renderer(cam, world, st_small).render_with_stats()
    assert full.tobytes() == small.tobytes() and sf == ss
    acc = np.zeros_like(full)
    rays = 0
    for rank in range(3):
        part, sp = gpu.Renderer(cam, world, rank_settings(st, rank, 3, tile_size=16)).render_with_stats()
        m = owned_mask(96, 64, 16, rank, 3)
        assert np.all(part[~m] == 0) and np.array_equal(part[m], full[m])
        acc += part
        rays += sp.total_rays()
    assert acc.tobytes() == full.tobytes() and rays == sf.total_rays()


def test_cached_pixel_list_is_keyed_by_the_real_tile_deal(gpu):
    """One World renders rank 0 of 2, then rank 0 of 3 (same image, same tile size), then rank 1 of 2 and the whole
    frame: the device-resident pixel list of the previous call must never be reused for a different deal
    (tile_world 2 and 3 once aliased in a hashed key)."""
    world, cam, st = scenes.furnace_room(gpu, width=96, height=64, spp=4, variance_threshold=0.0)
    full = gpu.Renderer(cam, world, st).render_with_stats()[0]
    for rank, n in ((0, 2), (0, 3), (0, 3), (1, 2), (0, 4), (0, 5), (3, 4)):
        part = gpu.Renderer(cam, world, rank_settings(st, rank, n, tile_size=16)).render_with_stats()[0]
        m = owned_mask(96, 64, 16, rank, n)
        assert np.all(part[~m] == 0) and np.array_equal(part[m], full[m]), (rank, n)
    again = gpu.Renderer(cam, world, st).render_with_stats()[0]
    assert again.tobytes() == full.tobytes()


def test_host_frame_tile_split_downloads_only_owned_pixels(gpu):
    """crt_render_host with a tile split resolves, downloads and scatters only the pixels this rank owns: the other
    ranks' pixels of the caller's frame stay as given (no upload, no whole-frame download), also when a device-frame
    render of another deal ran on the same World in between (the host-side scatter order carries its own key)."""
    import torch
    world, cam, st = scenes.furnace_room(gpu, width=96, height=64, spp=4, variance_threshold=0.0)
    full = gpu.Renderer(cam, world, st).render_with_stats()[0]
    dev = torch.zeros((64, 96, 4), dtype=torch.float32, device="cuda")
    frame = np.full((64, 96, 4), -7.0, dtype=np.float32)
    for rank, n in ((0, 2), (1, 2)):
        gpu.Renderer(cam, world, rank_settings(st, rank, n, tile_size=16)).render_with_stats(out=frame)
        # a device-frame render of a different deal with the same pixel count on the same World
        gpu.Renderer(cam, world, rank_settings(st, 1 - rank, n, tile_size=16)).render_device(dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
        m = owned_mask(96, 64, 16, rank, n)
        assert np.array_equal(frame[m], full[m])
        if rank == 0:
            assert np.all(frame[~m] == -7.0)
        gpu.Renderer(cam, world, rank_settings(st, rank, n, tile_size=16)).render_with_stats(out=frame)  # cached deal, host copy must be this deal's
        assert np.array_equal(frame[m], full[m])
    assert frame.tobytes() == full.tobytes()
    pinned = torch.empty((64, 96, 4), dtype=torch.float32).pin_memory()
    out = gpu.Renderer(cam, world, st).render_with_stats(out=pinned.numpy())[0]  # whole frame straight into pinned caller memory
    assert out.tobytes() == full.tobytes()


def test_device_framebuffer_entry_and_distributed_wrapper(gpu):
    import torch
    from crust_render_b200.distributed import render_distributed
    world, cam, st = scenes.furnace_room(gpu, width=64, height=48, spp=4, variance_threshold=0.0)
    host, sh = gpu.Renderer(cam, world, st).render_with_stats()
    fb = torch.empty((48, 64, 4), dtype=torch.float32, device="cuda")
    frame, sd = render_distributed(gpu.Renderer(cam, world, st), None, fb=fb, stream=torch.cuda.current_stream().cuda_stream)
    assert frame.cpu().numpy().tobytes() == host.tobytes() and sd == sh


def test_work_stealing_cursor_single_process(gpu):
    """crt_render_steal with a local cursor: the frame is pulled chunk by chunk (strided tile sets) and is
    bit-identical to the one-shot render, adaptive sampling included; a second frame needs a reset."""
    import torch
    from crust_render_b200.distributed import render_distributed
    world, cam, st = scenes.furnace_room(gpu, width=160, height=96, spp=24, variance_threshold=0.05, min_samples_per_pixel=8)
    full, sf = gpu.Renderer(cam, world, st).render_with_stats()
    fb = torch.zeros((96, 160, 4), dtype=torch.float32, device="cuda")
    for chunk in (1, 4, 1000):
        st.steal_chunk_tiles = chunk
        r = gpu.Renderer(cam, world, st)
        frame, ss = render_distributed(r, None, tile_size=32, fb=fb, stream=torch.cuda.current_stream().cuda_stream, steal=True)
        assert frame.cpu().numpy().tobytes() == full.tobytes(), chunk
        assert ss.camera_rays == sf.camera_rays and ss.total_rays() == sf.total_rays()
    # an exhausted cursor hands out nothing: the frame stays as it was given
    cur = gpu.StealCursor()
    r = gpu.Renderer(cam, world, rank_settings(st, 0, 1, 32))
    fb.fill_(7.0)
    r.render_steal(fb.data_ptr(), cur)
    assert not bool((fb == 7.0).any())
    fb.fill_(7.0)
    s2 = r.render_steal(fb.data_ptr(), cur)
    assert bool((fb == 7.0).all()) and s2.camera_rays == 0
    cur.reset()
    assert r.render_steal(fb.data_ptr(), cur).camera_rays == sf.camera_rays


def test_shared_frame_is_a_render_target(gpu):
    """crt_shared_frame_create: the IPC-exportable frame the peer-memory gather renders into (the cross-process half runs
    under torchrun: tools/steal_check.py). Here: it is zero-initialised device memory torch can view, a render into it
    equals a render into a torch buffer, and tile-split renders of two 'ranks' into the one frame assemble the whole image."""
    import torch
    world, cam, st = scenes.furnace_room(gpu, width=96, height=64, spp=4, variance_threshold=0.0)
    r = gpu.Renderer(cam, world, st)
    ref, _ = r.render_with_stats()
    frame = gpu.SharedFrame(64, 96)
    view = torch.as_tensor(frame, device="cuda")
    assert view.shape == (64, 96, 4) and float(view.abs().sum()) == 0.0 and len(frame.handle) == 64
    r.render_device(frame.ptr)
    assert np.array_equal(view.cpu().numpy(), ref)
    view.zero_()
    for rank in range(2):
        r.settings = rank_settings(st, rank, 2)
        r.render_device(frame.ptr)
    assert np.array_equal(view.cpu().numpy(), ref)
    del view
    frame.close()


def test_render_argument_errors(gpu):
    world, cam, st = scenes.furnace_room(gpu, width=16, height=16, spp=1)
    for field, value in (("width", 0), ("samples_per_pixel", 0), ("samples_per_pixel", 70000)):
        bad = scenes.furnace_room(gpu, width=16, height=16, spp=1)[2]
        setattr(bad, field, value)
        with pytest.raises(gpu.CrtError) as e:
            gpu.Renderer(cam, world, bad).render_device(0)
        assert e.value.status == 1
    bad = rank_settings(st, 3, 2)
    with pytest.raises(gpu.CrtError):
        gpu.Renderer(cam, world, bad).render_with_stats()


@pytest.mark.parametrize("name", ["furnace_room", "openpbr_showcase"])
def test_converged_renders_agree_across_sampler_seeds(gpu, orc, name):
    """BASELINE: "converged 4096-spp renders must match within 1e-3 relMSE". The sampler's values are unpinned against the
    reference (openqmc-rs is not vendored), so convergence must not depend on them: the GPU at 4096 spp and seed 0 against
    the ORACLE at 4096 spp and another seed (independent sample sets; the oracle against itself across these seeds measures
    2e-4 / 6e-5) — and, at the same seed, the usual tight bound."""
    fn = getattr(scenes, name)
    kw = dict(width=48, height=32 if name == "furnace_room" else 27, spp=4096, variance_threshold=0.0)
    wg, cam, st = fn(gpu, **kw)[:3]
    fg, _ = gpu.Renderer(cam, wg, st).render_with_stats()
    wo, _, st_o = fn(orc, **kw)[:3]
    fo_same, _ = orc.Renderer(cam, wo, st_o).render_with_stats()
    compare(fg, fo_same)
    st_o.frame = 7
    fo, _ = orc.Renderer(cam, wo, st_o).render_with_stats()
    a, b = fg[..., :3].astype(np.float64), fo[..., :3].astype(np.float64)
    relmse = float(((a - b) ** 2 / (b ** 2 + 1e-2)).mean())
    assert relmse <= 1e-3, relmse
    assert abs(a.mean() / b.mean() - 1.0) < 2e-3
