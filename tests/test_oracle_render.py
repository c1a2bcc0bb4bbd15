"""CPU tests of the integrator restatement (oracle/crt_oracle_render.hpp) and of the host-side logic
around the render entry: the reference's own veach_mis preview as a golden image, RayStats
bookkeeping (stats.rs:128-152), the adaptive-sampling rule (tracer.rs:609-617), tile splitting and
the world_size-2 frame reduce over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from crust_render_b200 import scenes
from crust_render_b200.distributed import owned_mask, rank_settings
from crust_render_b200.render_types import SamplingStrategy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def srgb_roundtrip(fb):
    """write_png of the reference CLI (crates/crust-render/src/main.rs:459-468), then back to linear; row 0 = top."""
    lin = np.clip(fb[::-1, :, :3].astype(np.float64), 0, 1)
    srgb = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055)
    q = np.floor(srgb * 255 + 0.5) / 255
    return np.where(q <= 0.04045, q / 12.92, ((q + 0.055) / 1.055) ** 2.4)


def test_veach_mis_matches_the_reference_preview_image(orc):
    """Golden vector from the reference itself: images/veach_mis_test.png (960x540), reduced 4x4 by
    tests/golden/make_veach_reference.py. Statistical comparison — the reference's sampler and spp are
    unknown, and its image is offset by a fraction of a pixel along silhouettes, so edges are masked."""
    ref = np.load(os.path.join(ROOT, "tests", "golden", "veach_mis_ref_240x135.npy")).astype(np.float64)
    world, cam, st = scenes.veach_mis(orc, spp=16)
    fb, stats = orc.Renderer(cam, world, st).render_with_stats()
    ours = srgb_roundtrip(fb).reshape(135, 4, 240, 4, 3).mean(axis=(1, 3))
    d = ours - ref
    assert abs(ours.mean() / ref.mean() - 1.0) < 0.01            # measured 0.0007
    assert np.abs(d).mean() < 0.012                               # measured 0.0069 (both images are noisy)
    gy, gx = np.gradient(ref.mean(axis=2))
    interior = np.hypot(gx, gy) < 0.01
    assert interior.mean() > 0.5
    assert np.abs(d[interior]).mean() < 0.008
    assert np.all(np.abs(d[interior].mean(axis=0)) < 0.003)       # no per-channel bias
    assert stats.camera_rays == 960 * 540 * 16 and stats.ended_depth < stats.camera_rays // 100


def test_ray_stats_bookkeeping(orc):
    world, cam, st = scenes.furnace_room(orc, width=64, height=48, spp=4, variance_threshold=0.0)
    fb, s = orc.Renderer(cam, world, st).render_with_stats()
    assert s.camera_rays == 64 * 48 * 4
    assert s.shadow_rays == s.vertices                 # one NEE sample per surface vertex (two lights, always samplable)
    assert s.closest_hit >= s.camera_rays and s.closest_hit <= s.camera_rays + s.vertices
    assert s.rr_killed <= s.rr_tested
    assert s.total_rays() == s.closest_hit + s.shadow_rays
    assert np.all(fb[..., 3] == 4) and np.all(np.isfinite(fb)) and fb[..., :3].min() >= 0.0
    # every path ends exactly one way: escaped, depth, roulette, or absorbed (no sample)
    assert s.ended_escaped + s.ended_depth + s.rr_killed <= s.camera_rays


def test_render_is_deterministic_and_thread_count_invariant(orc):
    world, cam, st = scenes.furnace_room(orc, width=48, height=32, spp=4, variance_threshold=0.0)
    r = orc.Renderer(cam, world, st)
    a, sa = r.render_with_stats(threads=1)
    b, sb = r.render_with_stats(threads=5)
    assert a.tobytes() == b.tobytes() and sa == sb


def test_strategies_agree_in_expectation(orc):
    """tracer.rs:49-60: the four estimators are unbiased and differ only in variance."""
    means = {}
    for strat in SamplingStrategy:
        world, cam, st = scenes.furnace_room(orc, width=48, height=32, spp=64, variance_threshold=0.0, sampling_strategy=strat)
        fb, s = orc.Renderer(cam, world, st).render_with_stats()
        means[strat] = fb[..., :3].mean()
        assert (s.shadow_rays == 0) == (strat == SamplingStrategy.BsdfOnly)
    ref = means[SamplingStrategy.PowerMis]
    assert abs(means[SamplingStrategy.BalanceMis] / ref - 1) < 0.03
    assert abs(means[SamplingStrategy.LightOnly] / ref - 1) < 0.06
    assert abs(means[SamplingStrategy.BsdfOnly] / ref - 1) < 0.15  # small lights: BSDF-only is the noisy one


def test_adaptive_sampling_rule(orc):
    """tracer.rs:602-617: never fewer than min_spp, stops only on multiples of 4, and the stop decision
    follows from the samples taken (a second render with spp = taken reproduces the pixel)."""
    world, cam, st = scenes.furnace_room(orc, width=48, height=32, spp=64, min_samples_per_pixel=10, variance_threshold=0.05)
    fb, s = orc.Renderer(cam, world, st).render_with_stats()
    taken = fb[..., 3].astype(int)
    assert taken.min() >= 12 and taken.max() <= 64
    early = taken < 64
    assert early.any() and np.all(taken[early] % 4 == 0)
    assert s.camera_rays == taken.sum()
    st.adaptive = False
    full, _ = orc.Renderer(cam, world, st).render_with_stats()
    assert np.all(full[..., 3] == 64)
    same = taken == 64
    assert np.array_equal(full[same][:, :3], fb[same][:, :3])
    # min_spp below 2 is raised to 2; threshold 0 disables the test entirely
    st2 = scenes.furnace_room(orc, width=16, height=8, spp=8, min_samples_per_pixel=0, variance_threshold=0.0)[2]
    assert np.all(orc.Renderer(cam, world, st2).render_with_stats()[0][..., 3] == 8)


def test_tile_split_union_equals_the_full_frame(orc):
    world, cam, st = scenes.furnace_room(orc, width=80, height=48, spp=4, variance_threshold=0.0)
    full, sf = orc.Renderer(cam, world, st).render_with_stats()
    acc = np.zeros_like(full)
    cams = 0
    for rank in range(3):
        part, sp = orc.Renderer(cam, world, rank_settings(st, rank, 3, tile_size=16)).render_with_stats()
        m = owned_mask(80, 48, 16, rank, 3)
        assert np.all(part[~m] == 0) and np.array_equal(part[m], full[m])
        acc += part
        cams += sp.camera_rays
    assert acc.tobytes() == full.tobytes() and cams == sf.camera_rays


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch.distributed as dist
import orc
from crust_render_b200 import scenes
from crust_render_b200.distributed import render_distributed
dist.init_process_group("gloo")
world, cam, st = scenes.furnace_room(orc, width=80, height=48, spp=4, variance_threshold=0.0)
frame, stats = render_distributed(orc.Renderer(cam, world, st), dist, tile_size=16)
if dist.get_rank() == 0:
    np.save({out!r}, frame)
    print("STATS", stats.camera_rays, stats.closest_hit, stats.shadow_rays, flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_process_render_over_gloo_matches_single_process(orc, tmp_path):
    """The N>1 host path (tile deal + sum-reduce of frame and RayStats) on CPU: world_size 2, gloo."""
    out = str(tmp_path / "frame.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    world, cam, st = scenes.furnace_room(orc, width=80, height=48, spp=4, variance_threshold=0.0)
    full, sf = orc.Renderer(cam, world, st).render_with_stats()
    assert np.load(out).tobytes() == full.tobytes()
    line = [l for l in res.stdout.splitlines() if l.startswith("STATS")][0].split()
    assert [int(v) for v in line[1:]] == [sf.camera_rays, sf.closest_hit, sf.shadow_rays]


# ---- volume regions (volume.rs, tracer.rs:1016-1077, 1166-1249): analytic probes of the oracle integrator ----------
def _slab_mean(orc, kind, **kw):
    w, c, s = scenes.volume_slab(orc, kind, **kw)
    fb, st = orc.Renderer(c, w, s).render_with_stats()
    return fb[..., :3].reshape(-1, 3).astype(np.float64).mean(axis=0), st


def test_absorbing_slab_is_beer_lambert(orc):
    mean, st = _slab_mean(orc, "absorb")
    assert np.abs(mean - np.exp(-0.7)).max() < 0.008  # 65 536 Bernoulli(e^-0.7) samples: sigma = 0.002
    assert st.vertices == 0 and st.shadow_rays == 0 and st.ended_escaped == st.camera_rays


def test_emissive_slab_matches_the_analytic_source_term(orc):
    mean, _ = _slab_mean(orc, "emit")
    expect = np.array([4.0, 1.5, 0.3]) * (1.0 - np.exp(-0.7)) + np.exp(-0.7)
    assert np.abs(mean / expect - 1.0).max() < 0.02


@pytest.mark.parametrize("kind", ["furnace", "furnace_grid", "furnace_noise"])
@pytest.mark.parametrize("strategy", [SamplingStrategy.PowerMis, SamplingStrategy.BalanceMis, SamplingStrategy.BsdfOnly])
def test_white_furnace_in_a_scattering_volume(orc, kind, strategy):
    """Albedo-1 medium under a uniform dome: volume NEE (phase value as the bsdf, tracer.rs:1040-1077) and the
    phase-sampled escapes (PrevVertex::Phase in escaped_emission, :995) must sum to exactly the dome radiance."""
    mean, st = _slab_mean(orc, kind, spp=16, sampling_strategy=strategy)
    assert np.abs(mean - 1.0).max() < 0.01, mean
    assert st.vertices > 0 and (st.shadow_rays == 0) == (strategy == SamplingStrategy.BsdfOnly)


def test_fog_scene_renders_and_fog_dims_the_room(orc):
    w, c, s = scenes.fog(orc, width=64, height=36, samples_per_pixel=16, variance_threshold=0.0)
    fb, st = orc.Renderer(c, w, s).render_with_stats()
    clear = scenes.fog(orc, width=64, height=36, samples_per_pixel=16, variance_threshold=0.0)[0].with_volumes([])
    fc, sc = orc.Renderer(c, clear, s).render_with_stats()
    assert np.isfinite(fb).all() and st.vertices > sc.vertices  # scatter vertices inside the fog add to the surface ones
    # in-scattering brightens the dark open front of the room, extinction dims the lit back wall
    assert fb[..., :3].mean() > 0.0 and abs(fb[..., :3].mean() / fc[..., :3].mean() - 1.0) > 0.02


def test_smoke_scene_renders_deterministically(orc):
    w, c, s = scenes.smoke(orc, width=48, height=27, samples_per_pixel=8, variance_threshold=0.0)
    a, sa = orc.Renderer(c, w, s).render_with_stats(threads=1)
    b, sb = orc.Renderer(c, w, s).render_with_stats(threads=4)
    assert a.tobytes() == b.tobytes() and sa == sb and np.isfinite(a).all()
