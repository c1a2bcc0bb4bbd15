"""The triangle test's hot path works in the frame (kx, ky, kz) = ((kz+1)%3, (kz+2)%3, kz) and leaves out the reference's swap
of kx and ky for d[kz] < 0 (triangle.rs:47-60); csrc/traverse.cuh `tri_eval` argues that the swap only negates e0, e1, e2, det
and t_scaled together and that no traversal decision sees that sign. This file checks the argument itself in numpy float32
(every operation rounded separately, like the kernel's -fmad=false arithmetic) on random and adversarial inputs: the five
values negate bit for bit, and the zero tests, the sign-mix test, both range tests and t come out identical."""
import numpy as np

F = np.float32


def _eval(o, d, v0, v1, v2, swap):
    """triangle.rs:41-80 + 123-136 in f32; `swap`: apply the reference's kx / ky exchange for d[kz] < 0."""
    n = len(o)
    ad = np.abs(d)
    kz = np.where(ad[:, 0] > ad[:, 1], np.where(ad[:, 0] > ad[:, 2], 0, 2), np.where(ad[:, 1] > ad[:, 2], 1, 2))
    kx = (kz + 1) % 3
    ky = (kx + 1) % 3
    i = np.arange(n)
    dz = d[i, kz]
    if swap:
        neg = dz < 0
        kx, ky = np.where(neg, ky, kx), np.where(neg, kx, ky)
    with np.errstate(all="ignore"):
        sx, sy, sz = d[i, kx] / dz, d[i, ky] / dz, F(1.0) / dz
        a, b, c = v0 - o, v1 - o, v2 - o
        akz, bkz, ckz = a[i, kz], b[i, kz], c[i, kz]
        ax, ay = a[i, kx] - sx * akz, a[i, ky] - sy * akz
        bx, by = b[i, kx] - sx * bkz, b[i, ky] - sy * bkz
        cx, cy = c[i, kx] - sx * ckz, c[i, ky] - sy * ckz
        e0 = bx * cy - by * cx
        e1 = cx * ay - cy * ax
        e2 = ax * by - ay * bx
        det = (e0 + e1) + e2
        ts = (e0 * (sz * akz) + e1 * (sz * bkz)) + e2 * (sz * ckz)
    return e0, e1, e2, det, ts


def _decisions(e0, e1, e2, det, ts, t_min, t_max):
    with np.errstate(all="ignore"):
        on_edge = (e0 == 0) | (e1 == 0) | (e2 == 0)
        mixed = ((e0 < 0) | (e1 < 0) | (e2 < 0)) & ((e0 > 0) | (e1 > 0) | (e2 > 0))
        nonzero = det != 0
        adet = np.abs(det)
        tss = np.where(det < 0, -ts, ts)
        in_packet = (tss >= t_min * adet) & (tss <= t_max * adet)  # triangle.rs:327-331
        out_scalar = ((det < 0) & ((ts > t_min * det) | (ts < t_max * det))) | ((det > 0) & ((ts < t_min * det) | (ts > t_max * det)))  # :163-168
        t = ts * (F(1.0) / det)
    return on_edge, mixed, nonzero, in_packet, out_scalar, t


def _cases(n, rng):
    o = rng.uniform(-5, 5, (n, 3)).astype(F)
    d = rng.normal(0, 1, (n, 3)).astype(F)
    v0 = rng.uniform(-5, 5, (n, 3)).astype(F)
    v1 = v0 + rng.normal(0, 1, (n, 3)).astype(F)
    v2 = v0 + rng.normal(0, 1, (n, 3)).astype(F)
    # adversarial thirds: rays through a vertex / along an edge (exact zeros in the edge functions), axis-aligned directions,
    # grid coordinates (exact products), tiny and huge magnitudes
    k = n // 6
    d[:k] = (v0[:k] - o[:k])
    t = rng.uniform(0, 1, (k, 1)).astype(F)
    d[k:2 * k] = (v0[k:2 * k] + t * (v1[k:2 * k] - v0[k:2 * k])) - o[k:2 * k]
    ax = rng.integers(0, 3, k)
    d[2 * k:3 * k] = 0
    d[2 * k + np.arange(k), ax] = rng.choice(np.array([-1, 1, -3, 0.5], F), k)
    for arr in (o, v0, v1, v2):
        arr[3 * k:4 * k] = np.round(arr[3 * k:4 * k])
    d[3 * k:4 * k] = np.round(d[3 * k:4 * k] * 2) + np.where(np.round(d[3 * k:4 * k] * 2) == 0, F(1), F(0))
    d[4 * k:5 * k] *= F(1e-18)
    d[5 * k:] *= F(1e18)
    return o, d, v0, v1, v2


def test_swap_negates_everything_and_changes_no_decision():
    rng = np.random.default_rng(5)
    o, d, v0, v1, v2 = _cases(600_000, rng)
    n = len(o)
    t_min = rng.choice(np.array([-2.0, 0.0, 1e-3, 0.5], F), n)
    t_max = rng.choice(np.array([0.75, 3.0, np.inf], F), n)
    ref = _eval(o, d, v0, v1, v2, swap=True)
    hot = _eval(o, d, v0, v1, v2, swap=False)
    ad = np.abs(d)
    kz = np.where(ad[:, 0] > ad[:, 1], np.where(ad[:, 0] > ad[:, 2], 0, 2), np.where(ad[:, 1] > ad[:, 2], 1, 2))
    swapped = d[np.arange(n), kz] < 0
    assert swapped.sum() > n // 4 and (~swapped).sum() > n // 4
    for name, a, b in zip(("e0", "e1", "e2", "det", "t_scaled"), ref, hot):
        same = ~swapped
        assert a[same].tobytes() == b[same].tobytes(), name
        x, y = a[swapped], b[swapped]
        nz = ~((x == 0) & (y == 0)) & ~(np.isnan(x) & np.isnan(y))
        assert np.array_equal(x[nz].view(np.uint32), (-y[nz]).view(np.uint32)), f"{name} is not the exact negation under the swap"
    for name, a, b in zip(("on_edge", "mixed_signs", "det != 0", "in_range_packet", "out_of_range_scalar"), _decisions(*ref, t_min, t_max), _decisions(*hot, t_min, t_max)):
        assert np.array_equal(a, b), name
    ta, tb = _decisions(*ref, t_min, t_max)[5], _decisions(*hot, t_min, t_max)[5]
    ok = (ref[3] != 0) & np.isfinite(ta)
    assert np.array_equal(ta[ok], tb[ok])  # numerically equal; only the sign of a zero may differ, which finalize re-derives in the reference's frame
    assert (np.signbit(ta[ok]) != np.signbit(tb[ok])).sum() == ((ta[ok] == 0) & (np.signbit(ta[ok]) != np.signbit(tb[ok]))).sum()
