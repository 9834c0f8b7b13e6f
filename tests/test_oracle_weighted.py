"""oracle/oracle_weighted.c (FilterParams::GAUSSIAN / CROSS on the reference's RGBA8 planes) against closed-form
answers and an independent numpy float64 evaluation.  The reference enumerates both types and reads neither
(include/filter.cuh:12), so these known answers are what pins the oracle; the CUDA kernels are then held
to the oracle bit for bit (tests/test_gpu_box.py)."""
import ctypes

import numpy as np
import pytest

from oracle import pyoracle as po


def _np_weighted(render, type, radius, sS=0.0, sC=0.0, sA=0.0, sN=0.0, albedo=None, normal=None):
    """float64 evaluation of DESIGN.md §3b, exp() from numpy (independent of the oracle's polynomial)."""
    H, W, _ = render.shape
    img = render[..., :3].astype(np.float64)
    ss = sS if sS > 0 else 0.5 * max(radius, 1)
    acc = np.zeros((H, W, 3))
    ws = np.zeros((H, W))
    for dx in range(-radius, radius + 1):
        for dy in range(-radius, radius + 1):
            ys, ye = max(0, -dy), min(H, H - dy)
            xs, xe = max(0, -dx), min(W, W - dx)
            p = (slice(ys, ye), slice(xs, xe))
            q = (slice(ys + dy, ye + dy), slice(xs + dx, xe + dx))
            e = np.full((ye - ys, xe - xs), (dx * dx + dy * dy) / (2 * ss * ss))
            if type == 2:
                for plane, s in ((render, sC), (albedo, sA), (normal, sN)):
                    if s > 0:
                        d = plane[..., :3].astype(np.float64)
                        e = e + ((d[p] - d[q]) ** 2).sum(-1) / (2 * s * s * 65025.0)
            w = np.exp(-e)
            acc[p] += w[..., None] * img[q]
            ws[p] += w
    return acc / ws[..., None]


def test_exp2_polynomial_accuracy():
    f = po.lib().oracle_exp2_neg
    f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
    xs = np.concatenate([np.linspace(-30, 0, 4001), -np.logspace(-6, 2, 200)]).astype(np.float32)
    got = np.array([f(float(x)) for x in xs], np.float64)
    ref = np.exp2(xs.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 3e-7
    assert f(0.0) == 1.0 and f(-1.0) == 0.5 and f(-200.0) == 0.0


def test_constant_image_is_a_fixed_point():
    img = np.full((40, 52, 4), 0, np.uint8)
    img[..., :3] = (17, 130, 255)
    for t in (1, 2):
        out = po.weighted_filter(img, type=t, radius=3, depth=2, sigmaColor=0.2, sigmaAlbedo=0.1, sigmaNormal=0.3,
                                 albedo=img, normal=img)
        assert np.array_equal(out[..., :3], img[..., :3]) and np.all(out[..., 3] == 0)   # round-to-nearest quotient


@pytest.mark.parametrize("type,kw", [(1, {}), (1, {"sS": 0.8}), (2, {"sS": 2.0, "sC": 0.1}),
                                     (2, {"sC": 0.3, "sA": 0.05, "sN": 0.2})])
def test_against_numpy_float64(type, kw):
    rng = np.random.default_rng(3)
    H, W, r = 37, 45, 2
    render = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    albedo = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    normal = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    out = po.weighted_filter(render, type=type, radius=r, sigmaSpace=kw.get("sS", 0.0), sigmaColor=kw.get("sC", 0.0),
                             sigmaAlbedo=kw.get("sA", 0.0), sigmaNormal=kw.get("sN", 0.0), albedo=albedo, normal=normal)
    ref = _np_weighted(render, type, r, albedo=albedo, normal=normal, **kw)
    # the oracle's fp32 quotient may round to the other code than the float64 one when it sits on a .5 boundary
    d = np.abs(out[..., :3].astype(np.float64) - np.floor(ref + 0.5))
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_cross_does_not_bleed_across_an_albedo_edge():
    H, W = 24, 32
    render = np.zeros((H, W, 4), np.uint8)
    render[:, :16, :3], render[:, 16:, :3] = 40, 200
    albedo = render.copy()
    out = po.weighted_filter(render, type=2, radius=3, sigmaSpace=3.0, sigmaAlbedo=0.02, albedo=albedo)
    assert np.array_equal(out[..., :3], render[..., :3])                  # sharp edge kept
    blur = po.weighted_filter(render, type=1, radius=3, sigmaSpace=3.0)
    assert 60 < int(blur[10, 15, 0]) < 180                                 # the plain Gaussian does bleed


def test_disabled_terms_reduce_cross_to_gaussian_and_depth_iterates():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (30, 41, 4), dtype=np.uint8)
    g = po.weighted_filter(img, type=1, radius=2, depth=1, sigmaSpace=1.3)
    c = po.weighted_filter(img, type=2, radius=2, depth=1, sigmaSpace=1.3)   # every cross term off
    assert np.array_equal(g, c)
    g2 = po.weighted_filter(img, type=1, radius=2, depth=2, sigmaSpace=1.3)
    assert np.array_equal(g2, po.weighted_filter(g, type=1, radius=2, depth=1, sigmaSpace=1.3))  # ping-pong == iteration
