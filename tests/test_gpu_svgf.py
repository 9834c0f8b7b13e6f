"""GPU parity of the SVGF path (rmd_svgf_* through the C ABI) against oracle/oracle_svgf.c on the
same seeded inputs.  Tolerances are BASELINE.json's: max-abs <= 1e-3 on linear radiance and
PSNR >= 60 dB; integer/decoded planes (guide, slope, history length) are compared bit-exactly."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import pyoracle as po  # noqa: E402
from raymarchdenoisercuda_b200.synth import synth_frame  # noqa: E402
from util import MAX_ABS_TOL, PSNR_MIN_DB, flat_gbuffer, psnr  # noqa: E402


def _dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).cuda()
            for a in arrs]


def _params(depth, **kw):
    import raymarchdenoisercuda_b200 as rmd
    return rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=depth, radius=2, **kw)


def _run_sequence(W, H, seed, frames, depth, check_planes=False, svgf=None):
    import raymarchdenoisercuda_b200 as rmd
    ctx = rmd.SvgfContext(W, H)
    orc = po.SvgfOracle(W, H)
    out_d = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    worst, worst_psnr = 0.0, 1e9
    _run_sequence.max_histlen = 0
    for f in range(frames):
        c, a, g, m = synth_frame(W, H, seed, f)
        dc, da, dg, dm = _dev(c, a, g, m)
        ctx.frame(dc, da, dg, dm, out_d, _params(depth), rmd.SvgfParams(**svgf) if svgf else None)
        torch.cuda.synchronize()
        ref = orc.frame(c, a, g, m, depth=depth, svgf=svgf)
        got = out_d.cpu().numpy()
        assert np.isfinite(got).all()
        err = float(np.abs(got[..., :3] - ref[..., :3]).max())
        worst = max(worst, err)
        worst_psnr = min(worst_psnr, psnr(got[..., :3], ref[..., :3]))
        if check_planes:
            assert np.array_equal(ctx.read_plane(5), orc.plane(po.PLANE_GUIDE)), f  # decoded guide: bit-exact
            assert np.array_equal(ctx.read_plane(6), orc.plane(po.PLANE_SLOPE)), f
            assert np.array_equal(ctx.read_plane(3), orc.plane(po.PLANE_HISTLEN)), f  # every predicate agreed
            _run_sequence.max_histlen = max(_run_sequence.max_histlen, int(ctx.read_plane(3).max()))
            mo = orc.plane(po.PLANE_MOMENTS)
            assert np.abs(ctx.read_plane(2) - mo).max() <= 1e-5 * max(1.0, float(np.abs(mo).max()))
            assert np.abs(ctx.read_plane(4) - orc.plane(po.PLANE_HISTORY_COLOR)).max() < MAX_ABS_TOL
            # filtered variance (out.w): relative tolerance, it spans orders of magnitude
            vg, vr = got[..., 3], ref[..., 3]
            assert np.abs(vg - vr).max() <= 1e-3 * max(1.0, float(vr.max()))
    ctx.close()
    return worst, worst_psnr


@pytest.mark.parametrize("shape", [(256, 144), (203, 117), (640, 360)])
def test_sequence_parity_full_pipeline(shape):
    W, H = shape
    worst, p = _run_sequence(W, H, 0x5EED0001, 6, 5, check_planes=True)
    assert worst <= MAX_ABS_TOL, worst
    assert p >= PSNR_MIN_DB, p


def test_long_sequence_reaches_history_cap_and_alpha_floor():
    """44 frames with the default parameters: the history length saturates at the cap (min(N+1, 32)) and the
    colour blend factor sits on its floor (max(1/N', 0.05) is active for N' > 20) — spec S2, SURVEY Appendix A.2.
    History length bit-exact on every frame, i.e. every reprojection predicate agreed for the whole sequence."""
    worst, p = _run_sequence(256, 144, 0x5EED0001, 44, 5, check_planes=True)
    assert _run_sequence.max_histlen == 32, _run_sequence.max_histlen
    assert worst <= MAX_ABS_TOL, worst
    assert p >= PSNR_MIN_DB, p


@pytest.mark.parametrize("depth", [0, 1, 2, 3, 4])
def test_every_level_count(depth):
    """Per-level parity: depth = k exposes the output of level k-1 (FilterParams::depth, reference filter.cuh:13)."""
    worst, p = _run_sequence(192, 108, 0x5EED0007, 3, depth)
    assert worst <= MAX_ABS_TOL and p >= PSNR_MIN_DB, (depth, worst, p)


@pytest.mark.parametrize("stop_after", [1, 2])
def test_temporal_and_variance_stages_in_isolation(stop_after):
    """Per-pass parity: four full frames build up a history, then frame 4 stops after the temporal pass
    (stop_after = 1: planes compared with the oracle's PRE-variance planes) or after the variance pass
    (stop_after = 2: post-variance planes).  A stopped frame leaves the context without a level-0 history, so each
    case runs on its own context."""
    import raymarchdenoisercuda_b200 as rmd
    W, H = 224, 120
    ctx = rmd.SvgfContext(W, H)
    orc = po.SvgfOracle(W, H)
    out_d = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for f in range(4):
        c, a, g, m = synth_frame(W, H, 0x5EED0003, f)
        ctx.frame(*_dev(c, a, g, m), out_d, _params(5))
        orc.frame(c, a, g, m, depth=5)
        torch.cuda.synchronize()
        assert np.array_equal(ctx.read_plane(3), orc.plane(po.PLANE_HISTLEN))
    c, a, g, m = synth_frame(W, H, 0x5EED0003, 4)
    ctx.set_stop_after(stop_after)
    ctx.frame(*_dev(c, a, g, m), out_d, _params(5))
    orc.frame(c, a, g, m, depth=5)
    tc, tv = ctx.read_plane(0), ctx.read_plane(1)
    ref_c = orc.plane(po.PLANE_TEMPORAL_COLOR_PRE if stop_after == 1 else po.PLANE_TEMPORAL_COLOR)
    ref_v = orc.plane(po.PLANE_TEMPORAL_VAR_PRE if stop_after == 1 else po.PLANE_TEMPORAL_VAR)
    assert np.abs(tc - ref_c).max() < 2e-4
    assert np.abs(tv - ref_v).max() <= 1e-3 * max(1.0, float(ref_v.max()))
    assert np.array_equal(ctx.read_plane(3), orc.plane(po.PLANE_HISTLEN))
    mo = orc.plane(po.PLANE_MOMENTS)
    assert np.abs(ctx.read_plane(2) - mo).max() <= 1e-5 * max(1.0, float(np.abs(mo).max()))
    if stop_after == 1:   # the two planes must really differ where the 7x7 pass ran, or the comparison is vacuous
        assert np.abs(orc.plane(po.PLANE_TEMPORAL_VAR_PRE) - orc.plane(po.PLANE_TEMPORAL_VAR)).max() > 0
    ctx.close()


def test_first_frame_all_short_history():
    """No history: every pixel takes the 7x7 spatial variance path (worst case of pass 2)."""
    worst, p = _run_sequence(200, 96, 0x5EED0009, 1, 5)
    assert worst <= MAX_ABS_TOL and p >= PSNR_MIN_DB, (worst, p)


def test_reset_restarts_history():
    import raymarchdenoisercuda_b200 as rmd
    W, H = 160, 96
    ctx = rmd.SvgfContext(W, H)
    out1 = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    out2 = torch.empty_like(out1)
    c, a, g, m = synth_frame(W, H, 5, 0)
    d = _dev(c, a, g, m)
    ctx.frame(*d, out1, _params(5))
    ctx.frame(*d, out2, _params(5))
    ctx.reset()
    ctx.frame(*d, out2, _params(5))
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    ctx.close()


def _run_variant(env, W=333, H=190, frames=3):
    import raymarchdenoisercuda_b200 as rmd
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctx = rmd.SvgfContext(W, H)
        out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        for f in range(frames):
            d = _dev(*synth_frame(W, H, 0x5EED0011, f))
            ctx.frame(*d, out, _params(5))
        torch.cuda.synchronize()
        res = out.clone()
        ctx.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return res


def test_tma_and_plain_load_paths_are_bit_identical():
    """Tile kernel: the TMA boxes must deliver exactly what coalesced loads with explicit zero-fill deliver."""
    tma = _run_variant({"RMD_ATROUS_RING": "0", "RMD_NO_TMA": "0"})
    plain = _run_variant({"RMD_ATROUS_RING": "0", "RMD_NO_TMA": "1"})
    assert torch.equal(tma, plain)


def test_ring_kernel_matches_tile_kernel():
    """The persistent ring kernel and the independent-tile kernel evaluate the same taps in a different
    order (columns grouped by |dx|): equal up to fp32 summation order."""
    ring = _run_variant({"RMD_ATROUS_RING": "1", "RMD_NO_TMA": "0"})
    tile = _run_variant({"RMD_ATROUS_RING": "0", "RMD_NO_TMA": "1"})
    assert float((ring[..., :3] - tile[..., :3]).abs().max()) < 2e-5
    assert float((ring[..., 3] - tile[..., 3]).abs().max()) <= 2e-5 * max(1.0, float(tile[..., 3].max()))


@pytest.mark.parametrize("variant", ["1", "3", "6", "7", "8", "7,1,3,0,8"])
def test_tile_kernel_variants_agree(variant):
    """Every compiled tile-kernel variant (RMD_ATROUS_VARIANT: centre terms staged by TMA or loaded from global
    memory, 100-tap or |dx|-grouped body, launch bounds) computes the same level: bit-identical when the tap order
    is the same, equal to fp32 summation order otherwise.  With and without programmatic dependent launch."""
    base = _run_variant({"RMD_ATROUS_VARIANT": "0", "RMD_PDL": "0"})
    got = _run_variant({"RMD_ATROUS_VARIANT": variant, "RMD_PDL": "7"})
    if variant == "1":
        assert torch.equal(got, base)
    else:
        assert float((got[..., :3] - base[..., :3]).abs().max()) < 2e-5
        assert float((got[..., 3] - base[..., 3]).abs().max()) <= 2e-5 * max(1.0, float(base[..., 3].max()))
    plain = _run_variant({"RMD_ATROUS_VARIANT": variant, "RMD_PDL": "5", "RMD_NO_TMA": "1"})
    assert torch.equal(got, plain)   # TMA boxes (incl. the neighbour-phase rows) == explicit loads with zero fill


@pytest.mark.parametrize("variant", ["11", "12", "15", "16", "6,16,15,12,11"])
def test_tile_order_prefetch_and_pair_variants_are_bit_identical(variant):
    """Variants 11-16 change the grouped kernel (6) without touching any accumulator's tap order: tiles enumerated
    lattice-row-major, L2 prefetch of a look-ahead tile, the centre column with the terms shared by two outputs of
    one thread evaluated once (IEEE multiplication and |x - y| are symmetric), and levels that alternate the direction
    of their tile walk.  All must return variant 6's bits,
    with TMA and with plain loads, at a size with ragged tiles and at one where every step has several lattice tiles."""
    for W, H in ((333, 190), (640, 272)):
        base = _run_variant({"RMD_ATROUS_VARIANT": "6", "RMD_PDL": "0"}, W=W, H=H)
        got = _run_variant({"RMD_ATROUS_VARIANT": variant, "RMD_PDL": "7"}, W=W, H=H)
        assert torch.equal(got, base), (W, H)
        ahead = _run_variant({"RMD_ATROUS_VARIANT": variant, "RMD_PDL": "7", "RMD_ATROUS_PREFETCH": "37"}, W=W, H=H)
        assert torch.equal(ahead, base), (W, H)   # L2 prefetch of the tile 37 further along the walk (off by default)
        plain = _run_variant({"RMD_ATROUS_VARIANT": variant, "RMD_PDL": "5", "RMD_NO_TMA": "1"}, W=W, H=H)
        assert torch.equal(plain, base), (W, H)


@pytest.mark.parametrize("frames", [1, 3, 6])
def test_variance_pass_paths_are_bit_identical(frames):
    """The 7x7 estimate has two walks per 32x8 tile — by position, two rows per thread (dense tiles), and through the
    compacted pixel list (sparse tiles) — and two CTA shapes.  A pixel must get the same bits from all of them: in
    band mode the same pixel falls into a differently aligned tile than in the single-context frame.  Frame 0 is all
    short-history (every tile dense), later frames mix dense and sparse tiles."""
    base = _run_variant({"RMD_VAR_DENSE_MIN": "257", "RMD_VAR_THREADS": "256"}, frames=frames)   # list walk only (round-1 shape)
    for env in ({"RMD_VAR_DENSE_MIN": "0", "RMD_VAR_THREADS": "128"},     # position walk only
                {"RMD_VAR_DENSE_MIN": "128", "RMD_VAR_THREADS": "128"},   # shipped mix
                {"RMD_VAR_DENSE_MIN": "64", "RMD_VAR_THREADS": "256"},
                {"RMD_VAR_DENSE_MIN": "257", "RMD_VAR_THREADS": "128"},
                {"RMD_VAR_REVERSE": "0"}):                                # tile list walked first to last
        assert torch.equal(_run_variant(env, frames=frames), base), env


def test_frames_can_be_captured_in_a_cuda_graph():
    """rmd_svgf_frame neither allocates nor synchronises, so a caller can capture it.  The context alternates its
    ping-pong planes by frame parity, hence TWO consecutive frames form one replayable graph (the second returns the
    context to the parity it was captured at).  Replays equal eager frames bit for bit; the PDL edges are captured too."""
    import raymarchdenoisercuda_b200 as rmd
    W, H, N = 320, 180, 7
    seq = [_dev(*synth_frame(W, H, 0x5EED0081, f)) for f in range(N)]
    eager, graphed = rmd.SvgfContext(W, H), rmd.SvgfContext(W, H)
    out_e = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ref = []
    for f in range(N):
        eager.frame(*seq[f], out_e, _params(5))
        ref.append(out_e.clone())
    slots = [[torch.empty_like(p) for p in seq[0]] for _ in range(2)]
    outs = [torch.empty_like(out_e) for _ in range(2)]
    graphed.frame(*seq[0], outs[0], _params(5))          # frame 0 eagerly: the captured frames have a history
    torch.cuda.synchronize()
    assert torch.equal(outs[0], ref[0])
    for s, f in zip(slots, (1, 2)):
        for d, src in zip(s, seq[f]):
            d.copy_(src)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graphed.frame(*slots[0], outs[0], _params(5))
        graphed.frame(*slots[1], outs[1], _params(5))
    for first in (1, 3, 5):
        for s, f in zip(slots, (first, first + 1)):
            for d, src in zip(s, seq[f]):
                d.copy_(src)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(outs[0], ref[first]) and torch.equal(outs[1], ref[first + 1]), first
    eager.close(); graphed.close()


def test_constant_image_fixed_point_1080p():
    """Size-independent property at BASELINE.json's configs[1] size: a constant frame is a fixed point
    of the whole pipeline (skip-and-renormalise borders, reference src/filter.cu:38-39)."""
    import raymarchdenoisercuda_b200 as rmd
    W, H = 1920, 1080
    c, a, g, m = flat_gbuffer(H, W, (0.5, 0.25, 1.0), albedo_u8=128)
    d = _dev(c, a, g, m)
    ctx = rmd.SvgfContext(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for _ in range(3):
        ctx.frame(*d, out, _params(5))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.abs(o[..., :3] - np.array([0.5, 0.25, 1.0], np.float32)).max() < 5e-6
    assert np.abs(o[..., 3]).max() < 1e-9
    ctx.close()


def test_1080p_band_parity_against_oracle_crop():
    """configs[1] size on the GPU, oracle on a crop far from the crop's own borders: the top rows of a
    1920x1080 first frame depend only on rows < 64 + halo, so an oracle run on the top 192 rows must
    agree on rows [0, 64)."""
    import raymarchdenoisercuda_b200 as rmd
    W, H, HC = 1920, 1080, 192
    c, a, g, m = synth_frame(W, H, 0x5EED0001, 0)
    ctx = rmd.SvgfContext(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.frame(*_dev(c, a, g, m), out, _params(5))
    torch.cuda.synchronize()
    orc = po.SvgfOracle(W, HC)
    ref = orc.frame(c[:HC], a[:HC], g[:HC], m[:HC], depth=5)
    got = out[:64].cpu().numpy()
    assert np.abs(got[..., :3] - ref[:64, :, :3]).max() <= MAX_ABS_TOL
    ctx.close()


def test_host_frame_path_matches_device_path():
    import raymarchdenoisercuda_b200 as rmd
    W, H = 320, 180
    ctx_d, ctx_h = rmd.SvgfContext(W, H), rmd.SvgfContext(W, H)
    out_d = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    outs_h = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
    keep = []
    outs_dev = []
    for f in range(4):
        c, a, g, m = synth_frame(W, H, 77, f)
        pinned = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).pin_memory() for x in (c, a, g, m)]
        keep.append(pinned)
        ctx_h.frame_host(*pinned, outs_h[f], _params(5))
        ctx_d.frame(*_dev(c, a, g, m), out_d, _params(5))
        torch.cuda.synchronize()
        outs_dev.append(out_d.cpu())
    ctx_h.host_wait()
    for f in range(4):
        assert torch.equal(outs_h[f], outs_dev[f]), f
    ctx_d.close(); ctx_h.close()


def test_argument_validation():
    import raymarchdenoisercuda_b200 as rmd
    ctx = rmd.SvgfContext(64, 48)
    out = torch.empty((48, 64, 4), dtype=torch.float32, device="cuda")
    d = _dev(*synth_frame(64, 48, 1, 0))
    with pytest.raises(rmd.RmdError) as e:
        ctx.frame(*d, out, rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=5, radius=2))
    assert e.value.code == -5   # AVERAGE / GAUSSIAN / CROSS are served by rmd_filter_*, not by the SVGF context
    with pytest.raises(rmd.RmdError) as e:
        ctx.frame(*d, out, rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=6, radius=2))
    assert e.value.code == -3
    with pytest.raises(rmd.RmdError) as e:
        ctx.frame(*d, out, rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=3))
    assert e.value.code == -3
    ctx.close()
    with pytest.raises(rmd.RmdError):
        rmd.SvgfContext(0, 10)


def test_history_pack_unpack_roundtrip_restores_a_sequence():
    """rmd_svgf_history_pack/unpack carry the complete frame-to-frame state: a second context fed the packed
    history continues the sequence bit-identically."""
    import raymarchdenoisercuda_b200 as rmd
    W, H = 192, 120
    a, b = rmd.SvgfContext(W, H), rmd.SvgfContext(W, H)
    out_a = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    out_b = torch.empty_like(out_a)
    frames = [_dev(*synth_frame(W, H, 21, f)) for f in range(5)]
    for f in range(3):
        a.frame(*frames[f], out_a, _params(5))
    buf = torch.empty(a.history_bytes(H), dtype=torch.uint8, device="cuda")
    a.history_pack(0, H, buf)
    b.frame(*frames[2], out_b, _params(5))   # any frame: makes b "have history" with the same parity as a
    b.frame(*frames[2], out_b, _params(5))
    b.frame(*frames[2], out_b, _params(5))
    b.history_unpack(0, H, buf)
    for f in (3, 4):
        a.frame(*frames[f], out_a, _params(5))
        b.frame(*frames[f], out_b, _params(5))
        torch.cuda.synchronize()
        assert torch.equal(out_a, out_b), f
    a.close(); b.close()


@pytest.mark.parametrize("nbands", [2, 3])
def test_row_banded_frames_match_single_context_bit_exactly(nbands):
    """The N-rank row-band path emulated on one GPU (N band contexts, copies instead of sends): the owned rows
    of every band must equal the single-context result bit for bit, frame after frame."""
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    W, H, halo = 256, 96 * nbands + 40, shard.banded_halo(5)
    assert halo == 80
    full = rmd.SvgfContext(W, H)
    out_full = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    bands = [shard.BandedSvgf(W, H, b, halo) for b in shard.row_bands(H, nbands)]
    outs = [torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda") for b in bands]
    for f in range(5):
        planes = _dev(*synth_frame(W, H, 0x5EED0031, f))
        full.frame(*planes, out_full, _params(5))
        for b, o in zip(bands, outs):
            b.ctx.frame(*[b.slice_rows(p).contiguous() for p in planes], o, _params(5))
        shard.exchange_in_process(bands)
        torch.cuda.synchronize()
        for b, o in zip(bands, outs):
            assert torch.equal(b.owned(o), out_full[b.band.row0:b.band.row0 + b.band.rows]), (f, b.band.rank)
    full.close()
    for b in bands:
        b.ctx.close()


def test_custom_parameters_and_rgba8_output():
    """Non-default sigmas (FilterParams) and SvgfParams, plus the optional RGBA8 `denoised`-format output
    (reference include/gbuffer.h:10): truncation of a value that differs by 1e-6 may differ by one code."""
    import raymarchdenoisercuda_b200 as rmd
    W, H = 200, 112
    svgf = {"alpha_color": 0.1, "alpha_moments": 0.3, "history_cap": 8, "short_history": 3, "depth_tolerance": 0.05,
            "normal_threshold": 0.95, "albedo_floor": 0.01, "variance_lum_scale": 5.0}
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    out8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=4, radius=2, sigmaSpace=2.0, sigmaColor=6.0, sigmaNormal=32.0)
    hmax = 0
    for f in range(12):   # > history_cap frames: the cap and the raised alpha floors are active
        c, a, g, m = synth_frame(W, H, 0x5EED0041, f)
        ctx.frame(*_dev(c, a, g, m), out, p, rmd.SvgfParams(**svgf), out_rgba8=out8)
        torch.cuda.synchronize()
        ref, ref8 = orc.frame(c, a, g, m, depth=4, sigma_z=2.0, sigma_l=6.0, sigma_n=32.0, svgf=svgf, want_rgba8=True)
        got = out.cpu().numpy()
        assert np.abs(got[..., :3] - ref[..., :3]).max() <= MAX_ABS_TOL, f
        assert np.array_equal(ctx.read_plane(3), orc.plane(po.PLANE_HISTLEN)), f
        hmax = max(hmax, int(ctx.read_plane(3).max()))
        assert hmax <= 8
        d8 = np.abs(out8.cpu().numpy().astype(np.int32) - ref8.astype(np.int32))
        assert d8.max() <= 1 and (d8 > 0).mean() < 0.01 and np.all(out8.cpu().numpy()[..., 3] == 255)
    assert hmax == 8   # the cap was reached, not merely respected
    ctx.close()


def test_cornell_fixture_configs0_parity():
    """BASELINE configs[0]: the reference's cornell frame (native 500x500, 5 a-trous levels, no history)."""
    import raymarchdenoisercuda_b200 as rmd
    from util import cornell_svgf_inputs
    npz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornell_gbuffer.npz"))
    c, a, g, m = cornell_svgf_inputs(npz)
    H, W, _ = c.shape
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.frame(*_dev(c, a, g, m), out, _params(5))
    torch.cuda.synchronize()
    ref = orc.frame(c, a, g, m, depth=5)
    got = out.cpu().numpy()
    assert np.abs(got[..., :3] - ref[..., :3]).max() <= MAX_ABS_TOL
    assert psnr(got[..., :3], ref[..., :3]) >= PSNR_MIN_DB
    ctx.close()


def test_stress_gbuffer_parity():
    """Adversarial inputs (tests/util.py:stress_gbuffer).  Radiance reaches ~1e3 here, where fp32 cannot hold an
    absolute 1e-3, so the bound is 1e-3 relative to max(1, |reference|); integer planes stay bit-exact."""
    import raymarchdenoisercuda_b200 as rmd
    from util import stress_gbuffer
    W, H = 208, 144
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for f in range(5):
        c, a, g, m = stress_gbuffer(W, H, 7, f)
        ctx.frame(*_dev(c, a, g, m), out, _params(5))
        torch.cuda.synchronize()
        ref = orc.frame(c, a, g, m, depth=5)
        got = out.cpu().numpy()
        assert np.isfinite(got).all(), f
        assert np.array_equal(ctx.read_plane(5), orc.plane(po.PLANE_GUIDE)), f
        assert np.array_equal(ctx.read_plane(3), orc.plane(po.PLANE_HISTLEN)), f
        rel = np.abs(got[..., :3] - ref[..., :3]) / np.maximum(1.0, np.abs(ref[..., :3]))
        assert rel.max() <= MAX_ABS_TOL, (f, float(rel.max()), np.unravel_index(np.argmax(rel), rel.shape))
    ctx.close()


@pytest.mark.parametrize("nbands", [2, 3])
def test_row_bands_with_per_level_exchange_match_single_context_bit_exactly(nbands):
    """Band mode v2 (no halo recompute for the a-trous levels, per-level boundary rows pushed to the neighbour),
    emulated with all bands on one GPU: owned rows equal the single-context frame bit for bit, 5 frames."""
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    W, H = 256, 100 * nbands + 20
    full = rmd.SvgfContext(W, H)
    out_full = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    bands = [shard.BandedSvgfV2(W, H, b) for b in shard.row_bands(H, nbands)]
    for i, b in enumerate(bands):
        b.connect_local(bands[i - 1] if i > 0 else None, bands[i + 1] if i + 1 < nbands else None)
    outs = [torch.zeros((b.ext_rows, W, 4), dtype=torch.float32, device="cuda") for b in bands]
    for f in range(5):
        planes = _dev(*synth_frame(W, H, 0x5EED0061, f))
        full.frame(*planes, out_full, _params(5))
        shard.frame_in_process_v2(bands, [[b.slice_rows(p).contiguous() for p in planes] for b in bands], outs, _params(5))
        torch.cuda.synchronize()
        for b, o in zip(bands, outs):
            assert torch.equal(b.owned(o), out_full[b.band.row0:b.band.row0 + b.band.rows]), (f, b.band.rank)
    assert all(b.timeouts() == 0 for b in bands)
    full.close()
    for b in bands:
        b.close()


def test_row_band_motion_beyond_the_history_margin_is_defined():
    """Band mode refreshes 21 history rows beyond either band edge.  A reprojection tap that falls beyond them is
    treated as outside the image (disoccluded, N' = 1) instead of reading stale rows: with a uniform motion of -25
    rows (bilinear taps at rows y-25 / y-24, 3x3 fallback search up to row y-24) the first 3 owned rows of the lower
    band restart their history, every other owned row equals the single-GPU frame (include/rmd_b200.h "Motion limit", RMD_BAND_MAX_MOTION_Y)."""
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    W, H = 256, 220
    c, a, g, m = flat_gbuffer(H, W, (0.5, 0.25, 1.0), albedo_u8=128)
    m2 = m.copy()
    m2[..., 1] = np.float16(-25.0)
    full = rmd.SvgfContext(W, H)
    out_full = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    bands = [shard.BandedSvgfV2(W, H, b) for b in shard.row_bands(H, 2)]
    bands[0].connect_local(None, bands[1])
    bands[1].connect_local(bands[0], None)
    outs = [torch.zeros((b.ext_rows, W, 4), dtype=torch.float32, device="cuda") for b in bands]
    for mv in (m, m2):
        planes = _dev(c, a, g, mv)
        full.frame(*planes, out_full, _params(5))
        shard.frame_in_process_v2(bands, [[b.slice_rows(p).contiguous() for p in planes] for b in bands], outs, _params(5))
    torch.cuda.synchronize()
    n_full = full.read_plane(3)[..., 0]
    assert np.all(n_full[:24] == 1) and np.all(n_full[24:] == 2)           # single GPU: rows < 24 reproject off-image
    lo = bands[1]
    n_band = lo.ctx.read_plane(3)[..., 0][lo.top:lo.top + lo.band.rows]
    assert np.all(n_band[:3] == 1)                                         # taps beyond the 21 refreshed rows
    assert np.array_equal(n_band[3:], n_full[lo.band.row0 + 3:])
    up = bands[0]
    assert np.array_equal(up.ctx.read_plane(3)[..., 0][:up.band.rows], n_full[:up.band.rows])
    assert shard.BAND_MAX_MOTION_Y == 13 and all(b.timeouts() == 0 for b in bands)
    full.close()
    for b in bands:
        b.close()


def test_row_band_flag_timeout_is_a_sticky_error():
    """A neighbour that never delivers its halo rows: the bounded flag wait (~2 s) gives up, bumps the host-visible
    word, and every later band call on the context fails with RMD_E_TIMEOUT instead of filtering stale rows
    (the round-1 kernel carried on silently)."""
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    W, H = 256, 320
    bands = [shard.BandedSvgfV2(W, H, b) for b in shard.row_bands(H, 3)]
    mid = bands[1]
    mid.connect_local(bands[0], bands[2])       # the neighbours exist but never run a stage: no push, no flag
    planes = [mid.slice_rows(p).contiguous() for p in _dev(*synth_frame(W, H, 0x5EED0071, 0))]
    out = torch.zeros((mid.ext_rows, W, 4), dtype=torch.float32, device="cuda")
    for s in (0, 1):
        mid.stage(s, *planes, out, _params(5))
    assert mid.timeouts() == 0
    mid.stage(2, *planes, out, _params(5))      # level 1 waits for the neighbours' level-0 rows
    torch.cuda.synchronize()
    assert mid.timeouts() >= 1
    with pytest.raises(rmd.RmdError) as e:
        mid.stage(3, *planes, out, _params(5))
    assert e.value.code == -9
    with pytest.raises(rmd.RmdError) as e:
        mid.frame(*planes, out, _params(5))
    assert e.value.code == -9
    for b in bands:
        b.close()


def test_reference_gbuffer_entry_point_on_cornell():
    """rmd_svgf_frame_gbuffer: the reference's own `GBuffer` (RGBA8 render/albedo/normal -> RGBA8 denoised) through
    SVGF, BASELINE configs[0].  The device conversion is mirrored in numpy fp32 (tests/util.py) and fed to the oracle;
    two calls (history accumulates with zero motion)."""
    import raymarchdenoisercuda_b200 as rmd
    from util import gbuffer_to_svgf_inputs_f32
    npz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornell_gbuffer.npz"))
    rgba = lambda a: np.ascontiguousarray(np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2))
    render, albedo, normal = rgba(npz["render"]), rgba(npz["albedo"]), rgba(npz["normal"])
    H, W, _ = render.shape
    d = [torch.from_numpy(x).cuda() for x in (render, albedo, normal)]
    den = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, H)
    frame = rmd.GBuffer((W, H), d[0], den, normal=d[2], albedo=d[1])
    c, a, g, m = gbuffer_to_svgf_inputs_f32(render, albedo, normal)
    for it in range(2):
        ctx.frame_gbuffer(frame, _params(5), out=out)
        torch.cuda.synchronize()
        ref, ref8 = orc.frame(c, a, g, m, depth=5, want_rgba8=True)
        assert np.array_equal(ctx.read_plane(5), orc.plane(po.PLANE_GUIDE)), it   # conversion + decode bit-exact
        got = out.cpu().numpy()
        assert np.abs(got[..., :3] - ref[..., :3]).max() <= MAX_ABS_TOL, it
        d8 = np.abs(den.cpu().numpy().astype(np.int32) - ref8.astype(np.int32))
        assert d8.max() <= 1 and (d8 > 0).mean() < 0.01, it
    assert int(ctx.read_plane(3).max()) == 2   # history accumulated over the two calls
    ctx.close()


def test_4k_two_frames_crop_parity_against_oracle():
    """configs[2] size (3840x2160, history path): the GPU runs the full frames, the oracle the top 224 rows of the
    same two frames.  Rows [0, 64) of either output depend on rows < 64 + 62 (a-trous reach) + 3 (variance window)
    + the motion of the synthetic camera, far from the crop's own bottom border."""
    import raymarchdenoisercuda_b200 as rmd
    W, H, HC = 3840, 2160, 224
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, HC)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for f in range(2):
        c, a, g, m = synth_frame(W, H, 0x5EED0002, f)
        assert np.abs(m[:HC].astype(np.float32)).max() < 16.0   # the margin argument above
        ctx.frame(*_dev(c, a, g, m), out, _params(5))
        torch.cuda.synchronize()
        ref = orc.frame(c[:HC], a[:HC], g[:HC], m[:HC], depth=5)
        got = out[:64].cpu().numpy()
        assert np.abs(got[..., :3] - ref[:64, :, :3]).max() <= MAX_ABS_TOL, f
        assert np.array_equal(ctx.read_plane(3)[:64], orc.plane(po.PLANE_HISTLEN)[:64]), f
    ctx.close()


def test_4k_eight_frames_history_path_crop_parity():
    """configs[2] (3840x2160 sequence, temporal history path): eight full frames on the GPU, the oracle on the top 384
    rows.  Rows [0, 64) of frame f depend on input rows below 64 + 65 (a-trous + variance reach of the frame)
    + f * 21 (per earlier frame: 16 rows of motion + 5 rows from the level-0 history back to its temporal input),
    i.e. < 276 for f = 7, and the crop's own bottom border (row 384) disturbs at most the rows above 384 - 212."""
    import raymarchdenoisercuda_b200 as rmd
    W, H, HC = 3840, 2160, 384
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, HC)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for f in range(8):
        c, a, g, m = synth_frame(W, H, 0x5EED0002, f)
        assert np.abs(m[:HC].astype(np.float32)).max() < 16.0   # the margin argument above
        ctx.frame(*_dev(c, a, g, m), out, _params(5))
        torch.cuda.synchronize()
        ref = orc.frame(c[:HC], a[:HC], g[:HC], m[:HC], depth=5)
        got = out[:64].cpu().numpy()
        assert np.abs(got[..., :3] - ref[:64, :, :3]).max() <= MAX_ABS_TOL, f
        assert np.array_equal(ctx.read_plane(3)[:64], orc.plane(po.PLANE_HISTLEN)[:64]), f
    assert int(ctx.read_plane(3)[:64].max()) == 8
    ctx.close()


@pytest.mark.parametrize("scheme", ["--v2", "--v1"])
def test_row_bands_across_two_real_ranks_bit_exact(scheme):
    """The N-rank band path on real ranks (one process per GPU, CUDA-IPC peer mappings, NVLink peer stores and
    stream-ordered flags): tools/band_p2p_check.py under torchrun, owned rows bit-identical to the single-GPU frame
    for 8 frames and no flag time-out.  Needs two devices."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(root, "tools", "band_p2p_check.py")] + ([scheme] if scheme == "--v2" else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "P2P BANDED CHECK" in r.stdout and " OK" in r.stdout.split("P2P BANDED CHECK")[-1], r.stdout[-2000:]


def test_8k_row_bands_with_per_level_exchange_bit_exact():
    """configs[3] size (7680x4320 row-banded, per-level halo exchange), the four bands emulated on one GPU: every
    band's owned rows equal the single-context frame bit for bit over two frames (the second uses the history)."""
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    W, H, nbands = 7680, 4320, 4
    full = rmd.SvgfContext(W, H)
    out_full = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    bands = [shard.BandedSvgfV2(W, H, b) for b in shard.row_bands(H, nbands)]
    for i, b in enumerate(bands):
        b.connect_local(bands[i - 1] if i > 0 else None, bands[i + 1] if i + 1 < nbands else None)
    outs = [torch.zeros((b.ext_rows, W, 4), dtype=torch.float32, device="cuda") for b in bands]
    for f in range(2):
        planes = _dev(*synth_frame(W, H, 0x5EED0003, f))
        full.frame(*planes, out_full, _params(5))
        shard.frame_in_process_v2(bands, [[b.slice_rows(p).contiguous() for p in planes] for b in bands], outs, _params(5))
        torch.cuda.synchronize()
        for b, o in zip(bands, outs):
            assert torch.equal(b.owned(o), out_full[b.band.row0:b.band.row0 + b.band.rows]), (f, b.band.rank)
    assert all(b.timeouts() == 0 for b in bands)
    full.close()
    for b in bands:
        b.close()


def test_committed_golden_vectors():
    """The CUDA path against tests/golden/svgf_golden.npz without executing the oracle: the 4-frame synthetic
    sequence (every frame, history length bit-exact) and the decimated cornell frame."""
    import sys
    import raymarchdenoisercuda_b200 as rmd
    from util import cornell_svgf_inputs
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    import make_svgf_golden as mk
    gold = np.load(os.path.join(gold_dir, "svgf_golden.npz"))
    ctx = rmd.SvgfContext(mk.SEQ_W, mk.SEQ_H)
    out = torch.empty((mk.SEQ_H, mk.SEQ_W, 4), dtype=torch.float32, device="cuda")
    for f in range(mk.SEQ_FRAMES):
        ctx.frame(*_dev(*synth_frame(mk.SEQ_W, mk.SEQ_H, mk.SEQ_SEED, f)), out, _params(5))
        torch.cuda.synchronize()
        ref = gold[f"seq_{f}"]
        assert np.abs(out.cpu().numpy()[..., :3] - ref[..., :3]).max() <= MAX_ABS_TOL, f
        assert np.array_equal(ctx.read_plane(3).reshape(ref.shape[:2]), gold[f"seq_histlen_{f}"].reshape(ref.shape[:2])), f
    ctx.close()
    c, a, g, m = cornell_svgf_inputs(np.load(os.path.join(gold_dir, "cornell_gbuffer.npz")))
    H, W, _ = c.shape
    ctx = rmd.SvgfContext(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.frame(*_dev(c, a, g, m), out, _params(5))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy()[::4, ::4, :3] - gold["cornell_dec4"][..., :3]).max() <= MAX_ABS_TOL
    ctx.close()
