"""Source-level drop-in: include/compat + librmd_compat.a.

  * the reference's API is two __global__ symbols the caller launches itself (include/filter.cuh:25-26, src/test.cu:73-75,
    85-87): librmd_compat.a defines them as relocatable device code, include/compat declares them with the reference's
    struct layouts;
  * the reference's OWN harness (src/test.cu + src/main.cpp, unmodified, compiled where it lies) builds against them
    -> oracle/_ref/ref_harness, the reference's `make test` (Makefile:61-62) on this library;
  * examples/compat_check.cu launches the kernels with `<<<grid, block, smem>>>` like the reference's sources do, for
    several block shapes / shared-memory sizes / depths, and its outputs are compared with the oracle bit for bit;
  * CudaGBuffer::openImages (declared, never defined by the reference, include/gbuffer.h:20-33) loads a render
    directory through the library's own PNG codec and runs BASELINE configs[0] through the SVGF path.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raymarchdenoisercuda_b200", "librmd_compat.a")
CHECK = os.path.join(ROOT, "examples", "build", "compat_check")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
GOLD = os.path.join(ROOT, "tests", "golden")


def test_archive_defines_the_reference_kernel_symbols():
    """Same mangled names as the reference's src/filter.cu (signature: GBuffer by value, FilterParams by value)."""
    out = subprocess.run(["nm", "-C", LIB], capture_output=True, text=True, check=True).stdout
    assert " T filterKernelBaseline(GBuffer, FilterParams)" in out
    assert " T filterKernelTiled(GBuffer, FilterParams)" in out
    sass = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
    assert "_Z20filterKernelBaseline7GBuffer12FilterParams" in sass and "_Z17filterKernelTiled7GBuffer12FilterParams" in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout


def test_compat_headers_keep_the_reference_layouts(tmp_path):
    src = tmp_path / "layout.cu"
    src.write_text('#include "filter.cuh"\n#include <cstddef>\n#include <cstdio>\n'
                   'int main() { FilterParams p{}; printf("%zu %zu %zu %zu %d %d\\n", sizeof(GBuffer), sizeof(FilterParams), '
                   'offsetof(GBuffer, buffer), offsetof(FilterParams, sigmaSpace), (int)p.cacheInput, (int)p.cacheBuffer); return 0; }\n')
    exe = tmp_path / "layout"
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-w", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.split() == ["56", "36", "40", "16", "1", "1"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_unmodified_reference_harness_builds_against_the_compat_layer():
    """Rebuilds oracle/_ref/ref_harness from /root/reference/src/{test.cu,main.cpp} (nothing is copied or edited)."""
    r = subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "_ref/ref_harness"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([HARNESS, "-h"], capture_output=True, text=True)   # src/main.cpp:5-10
    assert "-t [label]" in r.stdout


def test_png_codec_roundtrip_against_pil(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (29, 41, 3), dtype=np.uint8)
    for name, pil in (("rgb", Image.fromarray(img)), ("gray", Image.fromarray(img[..., 0])),
                      ("pal", Image.fromarray(img).convert("P", palette=Image.ADAPTIVE, colors=32)),
                      ("rgba", Image.fromarray(np.dstack([img, img[..., :1]])))):
        path = str(tmp_path / f"{name}.png")
        pil.save(path, optimize=True)   # PIL picks per-row filters: exercises Sub / Up / Average / Paeth
        r = subprocess.run([CHECK, "png", path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.array(Image.open(path + ".roundtrip.png")), np.array(Image.open(path).convert("RGBA"))), name
    r = subprocess.run([CHECK, "png", str(tmp_path / "missing.png")], capture_output=True, text=True)
    assert r.returncode == 2 and "Failed to load image" in r.stderr   # std::runtime_error like src/image.cpp:38-39


# ---- GPU ----------------------------------------------------------------------------------------------------------
def _planes(H, W, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in range(3)]


@pytest.mark.gpu
def test_caller_launched_kernels_match_the_oracle_bit_for_bit(tmp_path):
    from oracle import pyoracle as po
    H, W = 150, 203   # ragged against every block shape used
    render, albedo, normal = _planes(H, W, 11)
    for name, a in (("render", render), ("albedo", albedo), ("normal", normal)):
        np.save(tmp_path / f"{name}.npy", a)
    r = subprocess.run([CHECK, "kernels", str(tmp_path / "render.npy"), str(tmp_path / "albedo.npy"), str(tmp_path / "normal.npy"),
                        str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "wavelet error code -5" in r.stdout   # SVGF cannot be served by a stateless launch

    def got(name):
        return np.fromfile(tmp_path / f"{name}.bin", dtype=np.uint8).reshape(H, W, 4)
    box = lambda variant, depth, radius: po.box_filter(render, radius, depth, variant)
    assert np.array_equal(got("baseline_ref_call"), box("baseline", 1, 2))
    assert np.array_equal(got("tiled_ref_call"), box("tiled", 1, 2))
    assert np.array_equal(got("tiled_no_smem"), box("tiled", 1, 2))
    assert np.array_equal(got("tiled_depth5"), box("tiled", 5, 2))          # 5 levels in ONE launch == 5 host-iterated levels
    assert np.array_equal(got("baseline_depth3"), box("baseline", 3, 1))
    assert np.array_equal(got("baseline_subtiles"), box("baseline", 3, 1))
    assert np.array_equal(got("tiled_odd_block"), box("tiled", 2, 2))
    assert np.array_equal(got("tiled_radius7"), box("tiled", 1, 7))
    assert np.array_equal(got("gaussian_depth2"), po.weighted_filter(render, type=1, radius=2, depth=2))
    assert np.array_equal(got("cross_depth2"), po.weighted_filter(render, type=2, radius=3, depth=2, sigmaSpace=1.5, sigmaColor=0.25,
                                                                  sigmaAlbedo=0.05, sigmaNormal=0.3, albedo=albedo, normal=normal))


@pytest.mark.gpu
def test_cudagbuffer_open_images_runs_configs0_through_svgf(tmp_path):
    """CudaGBuffer::openImages on a render directory laid out like the reference's render/cornell/1 (PNG files written
    here from the committed fixture), SVGF through rmd_svgf_frame_gbuffer, denoisedCPU read back and saved as PNG."""
    import torch
    from PIL import Image
    import raymarchdenoisercuda_b200 as rmd
    npz = np.load(os.path.join(GOLD, "cornell_gbuffer.npz"))
    d = tmp_path / "cornell"
    d.mkdir()
    for k in ("render", "albedo", "normal"):
        Image.fromarray(npz[k]).save(d / f"{k}.png")
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([CHECK, "gbuffer", str(d), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "gbuffer 500 x 500" in r.stdout
    H, W = npz["render"].shape[:2]
    got = np.fromfile(out / "svgf_denoised.bin", dtype=np.uint8).reshape(H, W, 4)
    # the same frame through the Python mirror of the same entry point
    rgba = lambda a: np.ascontiguousarray(np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2))
    dev = [torch.from_numpy(rgba(npz[k])).cuda() for k in ("render", "albedo", "normal")]
    den = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    ctx = rmd.SvgfContext(W, H)
    ctx.frame_gbuffer(rmd.GBuffer((W, H), dev[0], den, normal=dev[2], albedo=dev[1]),
                      rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2))
    torch.cuda.synchronize()
    assert np.array_equal(got, den.cpu().numpy())
    assert np.array_equal(np.array(Image.open(out / "svgf_denoised.png")), got)     # Image::save
    from oracle import pyoracle as po
    assert np.array_equal(np.fromfile(out / "box_denoised.bin", dtype=np.uint8).reshape(H, W, 4), po.box_filter(rgba(npz["render"]), 2, 1, "tiled"))
    ctx.close()


@pytest.mark.gpu
def test_reference_make_test_runs_on_this_library():
    """`./build/main -t` of the reference (Makefile:61-62): its unmodified src/test.cu + src/main.cpp linked against
    librmd_compat.a.  The harness itself asserts nothing (src/test.cu:68-90); what is checked here is that both
    registered tests run, take a plausible time on the GPU and that the process exits cleanly."""
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/ref_harness was not built (no reference tree at build time)")
    r = subprocess.run([HARNESS, "-t"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "2 available tests: FILTER_BASELINE FILTER_TILED" in r.stdout
    assert r.stdout.count("Passed with") == 2 and "Fail" not in r.stdout
    r = subprocess.run([HARNESS, "-t", "FILTER_T.*"], capture_output=True, text=True, timeout=300)   # regex filter, src/test.cu:24-29
    assert r.stdout.count("TEST ") == 1 and "TEST FILTER_TILED" in r.stdout
