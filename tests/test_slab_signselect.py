"""The traversal kernel picks the near / far plane of each axis by the sign of the ray's slab reciprocal
(csrc/traverse.cu node_visit, CRT_SIGNSEL) where the reference takes min / max of the two plane distances
(crates/crust-rt/src/bvh.rs:790-808 `slab4`, :662-668 `safe_inv3`). This restates both in numpy f32 and checks
that the lane decision `tnear <= tfar` and the entry distance agree for every input the kernel can see
(value-equal: the sign of a zero plays no part in any comparison that follows)."""
import numpy as np

f32 = np.float32


def safe_inv(d):  # bvh.rs:662-668
    with np.errstate(divide="ignore"):
        return np.where(np.abs(d) < f32(1e-20), np.copysign(f32(1e20), d), f32(1.0) / d).astype(f32)


def slab_minmax(lo, hi, o, inv, t_min, t_max):
    t0 = ((lo - o) * inv).astype(f32)
    t1 = ((hi - o) * inv).astype(f32)
    near = np.maximum(np.minimum(t0, t1).max(axis=-1), t_min)
    far = np.minimum(np.maximum(t0, t1).min(axis=-1), t_max)
    return near, far


def slab_signsel(lo, hi, o, inv, t_min, t_max):
    neg = inv < 0  # the sign of the reciprocal, not of the direction (-0.0 travels backwards like copysign(1e20, -0.0))
    tn = ((np.where(neg, hi, lo) - o) * inv).astype(f32)
    tf = ((np.where(neg, lo, hi) - o) * inv).astype(f32)
    return np.maximum(tn.max(axis=-1), t_min), np.minimum(tf.min(axis=-1), t_max)


def _cases(rng, n):
    lo = rng.uniform(-50, 50, (n, 3)).astype(f32)
    ext = rng.uniform(0, 20, (n, 3)).astype(f32)
    ext[rng.random((n, 3)) < 0.1] = 0  # flat boxes: lo == hi
    hi = (lo + ext).astype(f32)
    o = rng.uniform(-60, 60, (n, 3)).astype(f32)
    o[rng.random((n, 3)) < 0.05] = 0
    sel = rng.random((n, 3)) < 0.05  # origins exactly on a face
    o[sel] = lo[sel]
    d = rng.normal(size=(n, 3)).astype(f32)
    kind = rng.random((n, 3))
    d[kind < 0.04] = f32(0.0)
    d[(kind >= 0.04) & (kind < 0.08)] = f32(-0.0)
    d[(kind >= 0.08) & (kind < 0.10)] = f32(1e-30)
    d[(kind >= 0.10) & (kind < 0.12)] = f32(-1e-30)
    d[(kind >= 0.12) & (kind < 0.14)] = f32(3e38)
    return lo, hi, o, d


def test_sign_selected_planes_equal_min_max():
    rng = np.random.default_rng(20251018)
    lo, hi, o, d = _cases(rng, 400_000)
    inv = safe_inv(d)
    assert np.all(np.isfinite(inv)) and not np.any(inv == 0)
    for t_min, t_max in ((f32(0.001), f32(np.inf)), (f32(0.0), f32(7.5)), (f32(-3.0), f32(1e30))):
        with np.errstate(over="ignore", invalid="ignore"):
            n0, f0 = slab_minmax(lo, hi, o, inv, t_min, t_max)
            n1, f1 = slab_signsel(lo, hi, o, inv, t_min, t_max)
        ok = ~(np.isnan(n0) | np.isnan(f0))  # inf - inf can only come from an infinite origin, which no finite ray has
        assert ok.all()
        assert np.array_equal(n0 <= f0, n1 <= f1)
        assert np.array_equal(n0, n1) and np.array_equal(f0, f1)  # == treats -0.0 and +0.0 alike, as every later comparison does


def test_empty_lane_encoding_never_hits():
    """Empty lanes of the device copy carry lo = +inf, hi = -inf (scene.cu mark_empty_lanes): the sign-selected test
    must reject them for every ray without consulting the child word (the reference excludes them by flag bits,
    bvh.rs:229-241,464)."""
    rng = np.random.default_rng(7)
    _, _, o, d = _cases(rng, 100_000)
    inv = safe_inv(d)
    lo = np.full_like(o, np.inf)
    hi = np.full_like(o, -np.inf)
    with np.errstate(over="ignore", invalid="ignore"):
        n1, f1 = slab_signsel(lo, hi, o, inv, f32(0.0), f32(np.inf))
    assert not np.any(n1 <= f1)
