"""Known-answer tests that pin oracle/oracle_svgf.c (parity with the reference is UNPINNED for
the SVGF stages: the reference has no such code and no golden vectors; the oracle
restates the published algorithm, DESIGN.md "SVGF specification").  Each case has a closed-form
answer or is checked against an independent numpy evaluation of the same equations."""
import numpy as np
import pytest

from oracle import pyoracle as po
from raymarchdenoisercuda_b200.synth import synth_frame
from util import flat_gbuffer, lum, make_guide

K = np.array([3 / 8, 1 / 4, 1 / 16])  # reference waveletSpline, src/filter.cu:10


def test_constant_image_is_a_fixed_point_including_borders():
    """Skip-and-renormalise border rule (reference src/filter.cu:38-39): a constant stays constant."""
    H, W = 37, 53
    c, a, g, m = flat_gbuffer(H, W, (0.5, 0.25, 1.0), albedo_u8=128)
    o = po.SvgfOracle(W, H)
    for depth in (0, 1, 5):
        o.reset()
        for _ in range(2):
            out = o.frame(c, a, g, m, depth=depth)
        expect = c[..., :3].astype(np.float32)
        assert np.abs(out[..., :3] - expect).max() < 2e-6, depth
        assert np.abs(out[..., 3]).max() < 1e-9  # no variance


def test_guide_decode_and_slope():
    H, W = 8, 9
    rng = np.random.default_rng(3)
    n = rng.normal(size=(H, W, 3))
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    z = (2 + np.arange(W)[None, :] * 0.25 + np.arange(H)[:, None] * 0.5).astype(np.float32)
    z[2, 3] = 0.0          # sky
    z[5, 5] = -1.0         # sky (z <= 0)
    c, a, _, m = flat_gbuffer(H, W, 1.0)
    o = po.SvgfOracle(W, H)
    o.frame(c, a, make_guide(n, z), m, depth=0)
    g4 = o.plane(po.PLANE_GUIDE)
    valid = z > 0
    assert np.all(g4[~valid] == 0)
    assert np.abs(np.linalg.norm(g4[..., :3][valid], axis=-1) - 1).max() < 1e-6
    assert np.abs(g4[..., :3][valid] - n[valid]).max() < 1e-4  # snorm16 octahedral quantisation
    assert np.array_equal(g4[..., 3][valid], z[valid])
    dz = o.plane(po.PLANE_SLOPE)[..., 0]
    zz = np.where(valid, z, 0)
    zx = np.abs(np.concatenate([zz[:, 1:], zz[:, -1:]], 1) - zz)
    zy = np.abs(np.concatenate([zz[1:], zz[-1:]], 0) - zz)
    assert np.array_equal(dz[valid], np.maximum(zx, zy)[valid])
    assert np.all(dz[~valid] == 0)


def test_orthogonal_normals_do_not_bleed():
    """w_n = max(0, n.n')^sigma_n is exactly 0 across a 90-degree crease: each side keeps its colour."""
    H, W = 24, 32
    rad = np.zeros((H, W, 3), np.float32)
    rad[:, :16] = 1.0
    rad[:, 16:] = 3.0
    nrm = np.zeros((H, W, 3))
    nrm[:, :16] = (0, 0, 1)
    nrm[:, 16:] = (1, 0, 0)
    c, a, _, m = flat_gbuffer(H, W, rad)
    g = make_guide(nrm, np.full((H, W), 4.0, np.float32))
    o = po.SvgfOracle(W, H)
    out = o.frame(c, a, g, m, depth=5)
    assert np.abs(out[:, :16, :3] - 1.0).max() < 1e-6
    assert np.abs(out[:, 16:, :3] - 3.0).max() < 1e-6


def test_sky_passes_through_and_carries_no_weight():
    H, W = 20, 20
    rng = np.random.default_rng(1)
    rad = rng.uniform(0.5, 1.5, (H, W, 3)).astype(np.float16).astype(np.float32)
    z = np.full((H, W), 3.0, np.float32)
    z[5:9, 4:12] = 0.0
    c, a, _, m = flat_gbuffer(H, W, rad, albedo_u8=200)
    c[5:9, 4:12, :3] = 50.0  # very bright sky must not leak into the surface
    g = make_guide(np.broadcast_to((0., 0., 1.), (H, W, 3)), z)
    o = po.SvgfOracle(W, H)
    out = o.frame(c, a, g, m, depth=5)
    assert np.array_equal(out[5:9, 4:12, :3], c[5:9, 4:12, :3].astype(np.float32))
    assert out[z > 0][:, :3].max() < 2.0


def test_level0_matches_numpy_b3_footprint():
    """With luminance/depth/normal terms inert the level is the 5x5 B3-spline convolution with
    skip-and-renormalise borders; the variance channel is filtered with the squared taps."""
    H, W = 19, 23
    rng = np.random.default_rng(5)
    o = po.SvgfOracle(W, H)
    g = make_guide(np.broadcast_to((0., 0., 1.), (H, W, 3)), np.full((H, W), 4.0, np.float32))
    m = np.zeros((H, W, 2), np.float16)
    a = np.full((H, W, 4), 255, np.uint8)
    frames = []
    for f in range(3):  # a few noisy frames so that the variance is non-zero
        c = np.ones((H, W, 4), np.float16)
        c[..., :3] = rng.uniform(0.2, 1.8, (H, W, 3))
        frames.append(c)
        out = o.frame(c, a, g, m, depth=1, sigma_l=1e9, svgf={"short_history": 1})
    tc = o.plane(po.PLANE_TEMPORAL_COLOR).astype(np.float64)
    tv = o.plane(po.PLANE_TEMPORAL_VAR)[..., 0].astype(np.float64)
    assert tv.min() >= 0 and tv.max() > 1e-3
    num = np.zeros((H, W, 3)); den = np.zeros((H, W)); vnum = np.zeros((H, W))
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            h = K[abs(dx)] * K[abs(dy)]
            ys, xs = np.mgrid[0:H, 0:W]
            qy, qx = ys + dy, xs + dx
            ok = (qy >= 0) & (qy < H) & (qx >= 0) & (qx < W)
            qy, qx = np.clip(qy, 0, H - 1), np.clip(qx, 0, W - 1)
            num += (h * ok)[..., None] * tc[qy, qx, :3]
            vnum += (h * h * ok) * tv[qy, qx]
            den += h * ok
    expect = num / den[..., None]
    assert np.abs(out[..., :3] - expect).max() < 2e-5   # exp(-|dL| / (1e9*sqrt(V)+1e-4)) ~ 1 - 1e-8
    assert np.abs(out[..., 3] - vnum / den ** 2).max() < 1e-5
    hist = o.plane(po.PLANE_HISTORY_COLOR)
    assert np.abs(hist[..., :3] - expect).max() < 2e-5
    assert np.abs(hist[..., 3] - lum(expect)).max() < 2e-5


def test_static_scene_history_is_the_running_mean():
    """Zero motion, depth 0: C' follows alpha = max(1/N', 0.05), moments alpha = max(1/N', 0.2),
    N' = min(frame + 1, cap)."""
    H, W = 6, 7
    rng = np.random.default_rng(11)
    o = po.SvgfOracle(W, H)
    g = make_guide(np.broadcast_to((0., 0., 1.), (H, W, 3)), np.full((H, W), 2.0, np.float32))
    m = np.zeros((H, W, 2), np.float16)
    a = np.full((H, W, 4), 128, np.uint8)
    C = M = None
    for f in range(40):
        c = np.ones((H, W, 4), np.float16)
        c[..., :3] = rng.uniform(0.0, 2.0, (H, W, 3))
        out = o.frame(c, a, g, m, depth=0, svgf={"short_history": 1})
        i = c[..., :3].astype(np.float32).astype(np.float64) / np.float64(np.float32(128) * np.float32(1 / 255))
        L = lum(i)
        mu = np.stack([L, L * L], -1)
        N = min(f + 1, 32)
        ac, am = max(np.float32(1) / np.float32(N), np.float32(0.05)), max(np.float32(1) / np.float32(N), np.float32(0.2))
        C = i if C is None else C + float(ac) * (i - C)
        M = mu if M is None else M + float(am) * (mu - M)
        assert np.all(o.plane(po.PLANE_HISTLEN) == N)
        assert np.abs(o.plane(po.PLANE_TEMPORAL_COLOR)[..., :3] - C).max() < 1e-5
        assert np.abs(o.plane(po.PLANE_MOMENTS) - M).max() < 1e-4
        var = np.maximum(0, M[..., 1] - M[..., 0] ** 2)
        assert np.abs(o.plane(po.PLANE_TEMPORAL_VAR)[..., 0] - var).max() < 1e-4
        a_f = np.float64(np.float32(128) * np.float32(1 / 255))
        assert np.abs(out[..., :3] - C * a_f).max() < 1e-5
        C = o.plane(po.PLANE_HISTORY_COLOR)[..., :3].astype(np.float64)  # fp32 storage of the history
        M = o.plane(po.PLANE_MOMENTS).astype(np.float64)


def test_integer_motion_reprojects_history_and_disoccludes_outside():
    """prev = p + motion with integer motion: history is fetched from the shifted texel; texels whose
    source falls outside the image (and has no valid 3x3 neighbour) restart with N' = 1."""
    H, W = 12, 16
    rng = np.random.default_rng(2)
    o = po.SvgfOracle(W, H)
    g = make_guide(np.broadcast_to((0., 0., 1.), (H, W, 3)), np.full((H, W), 2.0, np.float32))
    a = np.full((H, W, 4), 255, np.uint8)
    c0 = np.ones((H, W, 4), np.float16); c0[..., :3] = rng.uniform(0, 2, (H, W, 3))
    c1 = np.ones((H, W, 4), np.float16); c1[..., :3] = rng.uniform(0, 2, (H, W, 3))
    o.frame(c0, a, g, np.zeros((H, W, 2), np.float16), depth=0, svgf={"short_history": 1})
    mv = np.zeros((H, W, 2), np.float16); mv[..., 0] = 3; mv[..., 1] = 2
    o.frame(c1, a, g, mv, depth=0, svgf={"short_history": 1})
    N = o.plane(po.PLANE_HISTLEN)[..., 0]
    tc = o.plane(po.PLANE_TEMPORAL_COLOR)[..., :3]
    i0, i1 = c0[..., :3].astype(np.float32), c1[..., :3].astype(np.float32)
    ys, xs = np.mgrid[0:H, 0:W]
    inside = (xs + 3 < W) & (ys + 2 < H)
    assert np.all(N[inside] == 2)
    src = i0[np.clip(ys + 2, 0, H - 1), np.clip(xs + 3, 0, W - 1)]
    assert np.abs(tc[inside] - 0.5 * (src + i1)[inside]).max() < 1e-6
    far = (xs + 3 > W) | (ys + 2 > H)   # even the 3x3 search window is outside the image
    assert np.all(N[far] == 1)
    assert np.abs(tc[far] - i1[far]).max() == 0


def test_depth_and_normal_disocclusion_tests():
    """History behind a depth discontinuity (|dz| > 0.1 z + 2 slope) or a rotated normal (n.n' < 0.9) is rejected."""
    H, W = 8, 8
    a = np.full((H, W, 4), 255, np.uint8)
    m = np.zeros((H, W, 2), np.float16)
    c = np.ones((H, W, 4), np.float16)
    o = po.SvgfOracle(W, H)
    n0 = np.broadcast_to((0., 0., 1.), (H, W, 3))
    o.frame(c, a, make_guide(n0, np.full((H, W), 2.0, np.float32)), m, depth=0)
    z1 = np.full((H, W), 2.0, np.float32); z1[:, 4:] = 2.25   # 0.25 > 0.1*2.25 = 0.225 -> rejected
    z1[:, :2] = 2.125                                          # 0.125 < 0.2125      -> accepted
    o.frame(c, a, make_guide(n0, z1), m, depth=0)
    N = o.plane(po.PLANE_HISTLEN)[..., 0]
    assert np.all(N[:, 5:] == 1) and np.all(N[:, :2] == 2) and np.all(N[:, 2:3] == 2)
    o.reset()
    o.frame(c, a, make_guide(n0, np.full((H, W), 2.0, np.float32)), m, depth=0)
    n1 = np.zeros((H, W, 3)); n1[...] = (0, 0, 1)
    n1[:, 4:] = (np.sin(0.5), 0, np.cos(0.5))                  # cos(0.5) = 0.8776 < 0.9 -> rejected
    n1[:, :2] = (np.sin(0.3), 0, np.cos(0.3))                  # 0.955 -> accepted
    o.frame(c, a, make_guide(n1, np.full((H, W), 2.0, np.float32)), m, depth=0)
    N = o.plane(po.PLANE_HISTLEN)[..., 0]
    assert np.all(N[:, 4:] == 1) and np.all(N[:, :4] == 2)


def test_short_history_variance_is_spatial_and_boosted():
    """First frame: N' = 1 < 4, so colour/moments are 7x7 cross-bilateral means and the variance is
    4/N' times the spatial estimate; checked against a direct numpy evaluation of spec S3."""
    H, W = 15, 17
    rng = np.random.default_rng(9)
    c = np.ones((H, W, 4), np.float16); c[..., :3] = rng.uniform(0.5, 1.5, (H, W, 3))
    a = np.full((H, W, 4), 255, np.uint8)
    z = (3 + 0.125 * np.arange(W)[None, :] + 0 * np.arange(H)[:, None]).astype(np.float32)
    g = make_guide(np.broadcast_to((0., 0., 1.), (H, W, 3)), z)
    o = po.SvgfOracle(W, H)
    o.frame(c, a, g, np.zeros((H, W, 2), np.float16), depth=0)
    pre = o.plane(po.PLANE_TEMPORAL_COLOR_PRE).astype(np.float64)
    mom = o.plane(po.PLANE_MOMENTS).astype(np.float64)
    dz = o.plane(po.PLANE_SLOPE)[..., 0].astype(np.float64)
    post = o.plane(po.PLANE_TEMPORAL_COLOR)
    var = o.plane(po.PLANE_TEMPORAL_VAR)[..., 0]
    for (y, x) in [(0, 0), (7, 8), (14, 16), (3, 15)]:
        sw, sc, sm = 1.0, pre[y, x, :3].copy(), mom[y, x].copy()
        for dx in range(-3, 4):
            for dy in range(-3, 4):
                qx, qy = x + dx, y + dy
                if (dx == 0 and dy == 0) or not (0 <= qx < W and 0 <= qy < H):
                    continue
                tz = abs(float(z[y, x]) - float(z[qy, qx])) / (1.0 * max(dz[y, x], 1e-8) * np.hypot(dx, dy) + 1e-6)
                tl = abs(pre[y, x, 3] - pre[qy, qx, 3]) / 10.0
                w = np.exp(-tz - tl)
                sw += w; sc += w * pre[qy, qx, :3]; sm += w * mom[qy, qx]
        assert np.abs(post[y, x, :3] - sc / sw).max() < 1e-6
        v = max(0.0, sm[1] / sw - (sm[0] / sw) ** 2) * 4.0
        assert abs(var[y, x] - v) < 1e-6


def test_synthetic_sequence_denoises():
    """End-to-end sanity on the synthetic workload: after a few frames the filtered image is much
    closer to the noise-free irradiance-times-albedo than the 1-spp input is."""
    W, H = 160, 96
    o = po.SvgfOracle(W, H)
    for f in range(6):
        c, a, g, m = synth_frame(W, H, 0x5EED0001, f)
        out = o.frame(c, a, g, m, depth=5)
    assert np.isfinite(out).all()
    noisy = c[..., :3].astype(np.float32)
    # E[g] = 1, so the per-layer clean radiance is albedo*E; estimate it by heavy averaging of `out`
    valid = g[..., 1].view(np.float32) > 0
    assert out[valid][:, :3].std() < 0.6 * noisy[valid].std()
    N = o.plane(po.PLANE_HISTLEN)[..., 0]
    assert N.max() == 6 and (N[valid] == 1).mean() < 0.2  # disocclusions exist but are a minority


def test_cornell_fixture_spatial_only():
    """BASELINE configs[0]: the reference's only fixture through temporal(reset) + 7x7 variance + 5 levels.
    Sanity on the oracle: finite, denoised (less high-frequency energy than the input), mean preserved."""
    import os
    from util import cornell_svgf_inputs
    npz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornell_gbuffer.npz"))
    c, a, g, m = cornell_svgf_inputs(npz)
    H, W, _ = c.shape
    out = po.SvgfOracle(W, H).frame(c, a, g, m, depth=5)
    rad = c[..., :3].astype(np.float32)
    assert np.isfinite(out).all()
    hf_in = np.abs(np.diff(rad, axis=0)).mean()
    hf_out = np.abs(np.diff(out[..., :3], axis=0)).mean()
    assert hf_out < 0.5 * hf_in
    assert abs(float(out[..., :3].mean()) - float(rad.mean())) < 0.02


def test_stress_gbuffer_oracle_is_finite_and_tracks_history():
    from util import stress_gbuffer
    W, H = 96, 64
    o = po.SvgfOracle(W, H)
    for f in range(3):
        c, a, g, m = stress_gbuffer(W, H, 7, f)
        out = o.frame(c, a, g, m, depth=5)
        assert np.isfinite(out).all()
    N = o.plane(po.PLANE_HISTLEN)[..., 0]
    assert N.max() == 3 and N.min() == 0  # long histories where motion allows, sky = 0


def test_oracle_reproduces_the_committed_golden_vectors():
    """tests/golden/svgf_golden.npz (made by make_svgf_golden.py): any change to the oracle's arithmetic or to the
    synthetic scene generator shows up here before it can move the GPU parity target."""
    import os
    import sys
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    import make_svgf_golden as mk
    from raymarchdenoisercuda_b200.synth import synth_frame
    from util import cornell_svgf_inputs
    gold = np.load(os.path.join(gold_dir, "svgf_golden.npz"))
    orc = po.SvgfOracle(mk.SEQ_W, mk.SEQ_H)
    for f in range(mk.SEQ_FRAMES):
        out = orc.frame(*synth_frame(mk.SEQ_W, mk.SEQ_H, mk.SEQ_SEED, f), depth=5)
        assert np.abs(out - gold[f"seq_{f}"]).max() <= 1e-6, f
        assert np.array_equal(orc.plane(po.PLANE_HISTLEN), gold[f"seq_histlen_{f}"]), f
    c, a, g, m = cornell_svgf_inputs(np.load(os.path.join(gold_dir, "cornell_gbuffer.npz")))
    out = po.SvgfOracle(c.shape[1], c.shape[0]).frame(c, a, g, m, depth=5)[::4, ::4]
    ref = gold["cornell_dec4"]
    assert np.abs(out - ref).max() <= 1e-6 * max(1.0, float(np.abs(ref).max()))
