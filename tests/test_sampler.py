"""The path sampler both the CUDA integrator and the oracle use (csrc/shade/base.h): a 4-D
Owen-scrambled Sobol sequence behind OpenQMC's interface (the reference's `openqmc-rs` crate is not
vendored, so sample VALUES are unpinned — SURVEY §8c; these tests pin the properties the integrator
relies on: (0, m, 2)-net stratification per pixel, decorrelated domains, determinism)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pts(orc, n, keys=(), x=3, y=5, frame=0):
    return np.stack([orc.sampler_draw(x, y, frame, i, keys)[0] for i in range(n)])


def test_range_and_determinism(orc):
    a, b = pts(orc, 256), pts(orc, 256)
    assert np.array_equal(a, b)
    assert a.min() >= 0.0 and a.max() < 1.0
    r = [orc.sampler_draw(1, 2, 3, i, (5,))[1] for i in range(256)]
    assert 0.0 <= min(r) and max(r) < 1.0 and 0.4 < float(np.mean(r)) < 0.6


def test_first_two_dimensions_form_a_net(orc):
    """Owen scrambling preserves the (0, m, 2)-net property of Sobol dims (0, 1): every elementary
    interval of area 2^-m holds exactly one of the first 2^m points."""
    for keys in ((), (0,), (1, 2, 0)):
        p = pts(orc, 256, keys)
        for m in (4, 6, 8):
            q = p[: 1 << m, :2]
            for a in range(m + 1):
                cell = (np.floor(q[:, 0] * (1 << a)).astype(int) << (m - a)) + np.floor(q[:, 1] * (1 << (m - a))).astype(int)
                assert np.array_equal(np.sort(cell), np.arange(1 << m)), (keys, m, a)


def test_every_dimension_is_stratified_in_1d(orc):
    p = pts(orc, 128, (2,))
    for d in range(4):
        assert np.array_equal(np.sort(np.floor(p[:, d] * 128).astype(int)), np.arange(128))


def test_domains_pixels_and_frames_are_decorrelated(orc):
    base = pts(orc, 256, (0,))
    for other in (pts(orc, 256, (1,)), pts(orc, 256, (0,), x=4), pts(orc, 256, (0,), y=6), pts(orc, 256, (0,), frame=1), pts(orc, 256, (0, 0))):
        assert not np.array_equal(base, other)
        c = np.corrcoef(base[:, 0], other[:, 0])[0, 1]
        assert abs(c) < 0.25
    # mean of each dimension over many pixels ~ 0.5 (no pixel-to-pixel bias)
    m = np.mean([orc.sampler_draw(x, y, 0, 0, (0,))[0] for x in range(32) for y in range(32)], axis=0)
    assert np.all(np.abs(m - 0.5) < 0.05)


def test_direction_numbers_match_the_generator_script(orc):
    """tools/gen_sobol.py regenerates the Joe-Kuo matrices baked into base.h."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_sobol.py")], capture_output=True, text=True, check=True).stdout
    src = open(os.path.join(ROOT, "crust-render_b200", "csrc", "shade", "base.h")).read()
    nums = [w.strip(",{}") for w in out.replace("\n", " ").split() if w.strip(",{}").startswith("0x")]
    assert len(nums) >= 48
    for w in nums[:48]:
        assert w.lower() in src.lower(), w
