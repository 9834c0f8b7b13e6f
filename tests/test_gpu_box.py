"""GPU parity of the legacy box path (rmd_filter_baseline / rmd_filter_tiled through the C ABI)
against the oracle, the reference-generated golden vectors and — when oracle/_ref/libref_gpu.so
travelled with the snapshot — the reference's own kernels running on the same GPU.  Bit-exact."""
import ctypes
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import pyoracle  # noqa: E402
from util import sha16  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(img, radius, depth, variant):
    import raymarchdenoisercuda_b200 as rmd
    H, W, _ = img.shape
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.full_like(d_in, 0xAB)
    b0, b1 = torch.full_like(d_in, 0xAB), torch.full_like(d_in, 0xAB)
    frame = rmd.GBuffer((W, H), d_in, d_out, buffer=(b0, b1) if depth > 1 else (None, None))
    params = rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=depth, radius=radius)
    (rmd.filter_baseline if variant == "baseline" else rmd.filter_tiled)(frame, params)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("variant", ["tiled", "baseline"])
def test_cornell_matches_reference_golden(variant):
    gold = json.load(open(os.path.join(GOLD, "box_golden.json")))
    img = np.load(os.path.join(GOLD, "cornell_render_rgba.npz"))["render"]
    for depth in (1, 2, 5):
        out = _run(img, 2, depth, variant)
        hashed = out if variant == "tiled" else out[..., :3]
        assert sha16(hashed) == gold[variant][depth - 1]["sha"], (variant, depth)
        assert np.all(out[..., 3] == 0)


@pytest.mark.parametrize("variant", ["tiled", "baseline"])
def test_crop_vectors(variant):
    g = np.load(os.path.join(GOLD, "box_crop_golden.npz"))
    for depth in (1, 5):
        for radius in (1, 2, 3):
            out = _run(g["render"], radius, depth, variant)
            ref = g[f"{variant}_d{depth}_r{radius}"]
            sel = slice(None) if variant == "tiled" else slice(0, 3)
            assert np.array_equal(out[..., sel], ref[..., sel]), (variant, depth, radius)


@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (37, 21), (64, 32), (65, 33), (130, 257), (1080, 1920)])
def test_random_frames_vs_oracle(shape):
    H, W = shape
    rng = np.random.default_rng(H * 7919 + W)
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    radii = (1, 2, 7, 8, 32) if H * W < 100000 else (2,)
    for variant in ("tiled", "baseline"):
        for radius in radii:
            for depth in (1, 3):
                out = _run(img, radius, depth, variant)
                ref = pyoracle.box_filter(img, radius, depth, variant)
                assert np.array_equal(out, ref), (variant, radius, depth)


@pytest.mark.parametrize("shape", [(1, 4), (2, 8), (5, 116), (9, 120), (40, 124), (33, 244), (70, 364), (131, 480)])
def test_register_strip_path_vs_oracle(shape, monkeypatch):
    """W % 4 == 0 and radius 1..4 take box_strip_kernel<R> (warp column strips, 30 output quads per warp): widths
    around the 120-px warp span, heights around the strip height and the ring length, strips of 16 and 5 rows."""
    H, W = shape
    rng = np.random.default_rng(H * 131 + W)
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    img[: H // 2, : W // 2] = 255      # saturated block: the largest window sums
    for strip in ("", "5"):
        monkeypatch.setenv("RMD_BOX_STRIP", strip) if strip else monkeypatch.delenv("RMD_BOX_STRIP", raising=False)
        for variant in ("tiled", "baseline"):
            for radius in (1, 2, 3, 4):
                for depth in (1, 2):
                    out = _run(img, radius, depth, variant)
                    ref = pyoracle.box_filter(img, radius, depth, variant)
                    assert np.array_equal(out, ref), (variant, radius, depth, strip)


def test_strip_and_generic_kernels_agree_at_4k(monkeypatch):
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (2160, 3840, 4), dtype=np.uint8)
    fast = _run(img, 2, 2, "tiled")
    monkeypatch.setenv("RMD_BOX_GENERIC", "1")
    slow = _run(img, 2, 2, "tiled")
    assert np.array_equal(fast, slow)


def test_reference_kernels_on_this_gpu_agree():
    """The reference's own sm_100a build (unmodified src/filter.cu) vs this library vs the oracle."""
    if not os.path.exists(pyoracle.REF_GPU_LIB):
        pytest.skip("oracle/_ref/libref_gpu.so not in the snapshot")
    ref = ctypes.CDLL(pyoracle.REF_GPU_LIB)
    ref.ref_gpu_launch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    img = np.load(os.path.join(GOLD, "cornell_render_rgba.npz"))["render"]
    H, W, _ = img.shape
    d_in = torch.from_numpy(img).cuda()
    for variant, vid in (("baseline", 0), ("tiled", 1)):
        d_out = torch.zeros_like(d_in)
        rc = ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, vid, 0, None)
        torch.cuda.synchronize()
        assert rc == 0
        r = d_out.cpu().numpy()
        ours = _run(img, 2, 1, variant)
        sel = slice(None) if variant == "tiled" else slice(0, 3)
        assert np.array_equal(r[..., sel], ours[..., sel]), variant
        assert np.array_equal(r[..., sel], pyoracle.box_filter(img, 2, 1, variant)[..., sel])


def test_error_codes():
    import raymarchdenoisercuda_b200 as rmd
    d = torch.zeros((16, 16, 4), dtype=torch.uint8, device="cuda")
    frame = rmd.GBuffer((16, 16), d, d.clone())
    with pytest.raises(rmd.RmdError) as e:
        rmd.filter_tiled(frame, rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=2, radius=2))
    assert e.value.code == -1  # depth > 1 needs buffer[0..1] (reference src/filter.cu:24-25)
    with pytest.raises(rmd.RmdError) as e:
        rmd.filter_tiled(frame, rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=1, radius=2))
    assert e.value.code == -5  # the SVGF path needs a per-sequence context (rmd_svgf_frame_gbuffer)


# ---- FilterParams::GAUSSIAN / CROSS (reference include/filter.cuh:12, 16-19) -------------------------------------
def _run_weighted(render, albedo, normal, variant, **kw):
    import raymarchdenoisercuda_b200 as rmd
    H, W, _ = render.shape
    d = [torch.from_numpy(x).cuda() if x is not None else None for x in (render, albedo, normal)]
    d_out = torch.full_like(d[0], 0xAB)
    b0, b1 = torch.full_like(d[0], 0xAB), torch.full_like(d[0], 0xAB)
    depth = kw.get("depth", 1)
    frame = rmd.GBuffer((W, H), d[0], d_out, normal=d[2], albedo=d[1], buffer=(b0, b1) if depth > 1 else (None, None))
    (rmd.filter_baseline if variant == "baseline" else rmd.filter_tiled)(frame, rmd.FilterParams(**kw))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (37, 45), (64, 96), (130, 257), (500, 500)])
@pytest.mark.parametrize("case", [
    dict(type=1, radius=2, depth=1),
    dict(type=1, radius=4, depth=3, sigmaSpace=1.7),
    dict(type=2, radius=2, depth=1, sigmaSpace=2.0, sigmaColor=0.1),
    dict(type=2, radius=3, depth=2, sigmaSpace=1.5, sigmaColor=0.25, sigmaAlbedo=0.05, sigmaNormal=0.3),
    dict(type=2, radius=9, depth=1, sigmaNormal=0.1),
])
def test_gaussian_and_cross_bit_exact_vs_oracle(shape, case):
    """Both types through both entry points, bit for bit against oracle/oracle_weighted.c: every weight is an
    explicit fmaf chain + the shared 2^x polynomial, every squared distance an integer."""
    H, W = shape
    if H * W > 100000 and case["radius"] > 4:
        pytest.skip("large radius on a large frame: covered by the small shapes")
    rng = np.random.default_rng(H * 31 + W + case["radius"])
    render, albedo, normal = (rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in range(3))
    ref = pyoracle.weighted_filter(render, albedo=albedo, normal=normal, **case)
    for variant in ("tiled", "baseline"):
        out = _run_weighted(render, albedo, normal, variant, **case)
        assert np.array_equal(out, ref), (variant, case)


def test_cross_on_the_cornell_fixture():
    """BASELINE configs[0]'s planes (render / albedo / normal of render/cornell/1) through FilterParams::CROSS."""
    npz = np.load(os.path.join(GOLD, "cornell_gbuffer.npz"))
    rgba = lambda a: np.ascontiguousarray(np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2))
    render, albedo, normal = rgba(npz["render"]), rgba(npz["albedo"]), rgba(npz["normal"])
    case = dict(type=2, radius=3, depth=2, sigmaSpace=2.0, sigmaColor=0.2, sigmaAlbedo=0.1, sigmaNormal=0.25)
    out = _run_weighted(render, albedo, normal, "tiled", **case)
    assert np.array_equal(out, pyoracle.weighted_filter(render, albedo=albedo, normal=normal, **case))
    # it denoises: less high-frequency energy than the input, and not the identity
    hf = lambda a: float(np.abs(np.diff(a[..., :3].astype(np.int32), axis=0)).mean())
    assert hf(out) < 0.7 * hf(render)


def test_weighted_argument_validation():
    import raymarchdenoisercuda_b200 as rmd
    img = torch.zeros((16, 16, 4), dtype=torch.uint8, device="cuda")
    out = torch.zeros_like(img)
    frame = rmd.GBuffer((16, 16), img, out)
    with pytest.raises(rmd.RmdError) as e:   # albedo term on, no albedo plane
        rmd.filter_tiled(frame, rmd.FilterParams(type=2, depth=1, radius=2, sigmaAlbedo=0.1))
    assert e.value.code == -1
    with pytest.raises(rmd.RmdError) as e:
        rmd.filter_tiled(frame, rmd.FilterParams(type=1, depth=1, radius=2, sigmaSpace=-1.0))
    assert e.value.code == -3
    with pytest.raises(rmd.RmdError) as e:   # WAVELET needs a per-sequence context
        rmd.filter_tiled(frame, rmd.FilterParams(type=3, depth=1, radius=2))
    assert e.value.code == -5
