"""The C-ABI library loads and exports every symbol include/rmd_b200.h declares; the PODs that
cross the boundary keep the reference's layout (GBuffer 56 B, FilterParams 36 B).
No compute calls: runs without a GPU."""
import ctypes
import os
import re

from raymarchdenoisercuda_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for header in ("rmd_b200.h", "rmd_b200_debug.h"):   # the drop-in boundary + the test/inspection hooks
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(rmd_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_debug_hooks_are_not_in_the_product_header():
    text = open(os.path.join(ROOT, "include", "rmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for hook in ("rmd_svgf_set_stop_after", "rmd_svgf_read_plane", "rmd_svgf_last_launch_count"):
        assert hook not in text


def test_header_symbols_all_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rmd_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_pod_layouts_match_reference():
    lib = _lib.load()
    # reference include/gbuffer.h:6-14 and include/filter.cuh:11-23 (sizes probed in SURVEY.md §2.1/§8a)
    assert lib.rmd_sizeof_gbuffer() == 56 == ctypes.sizeof(_lib.RmdGBuffer)
    assert lib.rmd_sizeof_filter_params() == 36 == ctypes.sizeof(_lib.RmdFilterParams)
    assert _lib.RmdGBuffer.render.offset == 8 and _lib.RmdGBuffer.denoised.offset == 16
    assert _lib.RmdGBuffer.normal.offset == 24 and _lib.RmdGBuffer.albedo.offset == 32
    assert _lib.RmdGBuffer.buffer.offset == 40
    assert _lib.RmdFilterParams.radius.offset == 12 and _lib.RmdFilterParams.sigmaSpace.offset == 16
    assert _lib.RmdFilterParams.cacheInput.offset == 32 and _lib.RmdFilterParams.cacheBuffer.offset == 33


def test_error_strings_and_version():
    lib = _lib.load()
    assert lib.rmd_version() == 100
    assert b"null" in lib.rmd_error_string(-1)
    assert lib.rmd_error_string(0) == b"ok"


def test_argument_errors_need_no_gpu():
    """Validation happens before any CUDA call (reference kernels validate nothing, src/test.cu:73-77)."""
    lib = _lib.load()
    assert lib.rmd_filter_baseline(None, None, None) == -1
    g = _lib.RmdGBuffer(0, 0)
    p = _lib.RmdFilterParams(0, 1, 0, 2)
    assert lib.rmd_filter_tiled(ctypes.byref(g), ctypes.byref(p), None) == -2
    g = _lib.RmdGBuffer(16, 16)
    assert lib.rmd_filter_tiled(ctypes.byref(g), ctypes.byref(p), None) == -1  # null planes
    p_bad = _lib.RmdFilterParams(3, 1, 0, 2)
    g.render, g.denoised = 256, 512
    assert lib.rmd_filter_tiled(ctypes.byref(g), ctypes.byref(p_bad), None) == -5  # WAVELET is not the box path
    p_bad = _lib.RmdFilterParams(0, 1, 0, 99)
    assert lib.rmd_filter_tiled(ctypes.byref(g), ctypes.byref(p_bad), None) == -3
    assert lib.rmd_svgf_frame(None, None, None, None, None) == -1
    assert lib.rmd_svgf_reset(None) == -1
    assert lib.rmd_svgf_frame_gbuffer(None, None, None, None, None, None) == -1


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (a CPU fallback would void parity)."""
    pkg = os.path.join(ROOT, "raymarchdenoisercuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for ln in text.splitlines():
                    assert not re.search(r"^\s*(from|import)\s+oracle", ln), (f, ln)
                    assert "liboracle" not in ln, (f, ln)


def test_level_kernel_variant_lists_agree():
    """svgf_atrous_tile.cu is compiled once per variant: the list build.py compiles, the list the dispatcher links
    (RMD_ATROUS_VARIANTS in csrc/svgf.cuh) and the variants the source defines must be the same set, and the shipped
    default must be one of them."""
    from raymarchdenoisercuda_b200 import build
    csrc = os.path.join(ROOT, "raymarchdenoisercuda_b200", "csrc")
    header = open(os.path.join(csrc, "svgf.cuh")).read()
    listed = [int(v) for v in re.findall(r"X\((\d+)\)", re.search(r"#define RMD_ATROUS_VARIANTS\(X\)(.*)", header).group(1))]
    assert sorted(listed) == sorted(build.ATROUS_VARIANTS)
    defined = [int(v) for v in re.findall(r"#(?:el)?if RMD_VARIANT == (\d+)", open(os.path.join(csrc, "svgf_atrous_tile.cu")).read())]
    assert sorted(defined) == sorted(listed)
    default = [int(v) for v in re.search(r"kAtrousDefaultVariant\[[^\]]*\]\s*=\s*\{([^}]*)\}", header).group(1).split(",")]
    assert len(default) == 5 and all(v in listed for v in default)
    for v in listed:   # and every variant's object went into the library
        assert os.path.exists(os.path.join(csrc, "build", f"svgf_atrous_tile_v{v}.o"))
