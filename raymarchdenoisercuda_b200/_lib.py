"""ctypes loader of librmd_b200.so (the C ABI of include/rmd_b200.h).  Fails loudly when the
library has not been built: there is no Python/CPU fallback for any entry point."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librmd_b200.so")
SYNTH_PATH = os.path.join(_HERE, "librmd_synth.so")

# every symbol include/rmd_b200.h declares (tests/test_abi.py parses the header and compares)
SYMBOLS = [
    "rmd_filter_baseline", "rmd_filter_tiled",
    "rmd_svgf_create", "rmd_svgf_destroy", "rmd_svgf_reset", "rmd_svgf_frame", "rmd_svgf_frame_host", "rmd_svgf_frame_gbuffer",
    "rmd_svgf_host_wait", "rmd_svgf_last_launch_count", "rmd_svgf_set_profiling", "rmd_svgf_get_pass_times", "rmd_svgf_read_plane", "rmd_svgf_set_stop_after", "rmd_svgf_history_bytes", "rmd_svgf_history_pack",
    "rmd_svgf_history_unpack", "rmd_p2p_alloc", "rmd_p2p_free", "rmd_p2p_export", "rmd_p2p_open", "rmd_p2p_close",
    "rmd_p2p_signal", "rmd_p2p_wait", "rmd_p2p_timeouts", "rmd_svgf_band_configure", "rmd_svgf_band_recv_bytes",
    "rmd_svgf_band_stage", "rmd_svgf_band_frame", "rmd_svgf_band_timeouts", "rmd_svgf_band_launch_count",
    "rmd_svgf_prepare_host", "rmd_svgf_prepare_gbuffer", "rmd_debug_clock_probe", "rmd_debug_level_cover",
    "rmd_error_string", "rmd_version", "rmd_sizeof_gbuffer", "rmd_sizeof_filter_params",
]


class RmdGBuffer(ctypes.Structure):
    """Mirror of reference `struct GBuffer` (include/gbuffer.h:6-14), 56 bytes."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("render", ctypes.c_void_p),
                ("denoised", ctypes.c_void_p), ("normal", ctypes.c_void_p), ("albedo", ctypes.c_void_p),
                ("buffer", ctypes.c_void_p * 2)]


class RmdFilterParams(ctypes.Structure):
    """Mirror of reference `struct FilterParams` (include/filter.cuh:11-23), 36 bytes."""
    _fields_ = [("type", ctypes.c_int32), ("depth", ctypes.c_int32), ("level", ctypes.c_int32),
                ("radius", ctypes.c_int32), ("sigmaSpace", ctypes.c_float), ("sigmaColor", ctypes.c_float),
                ("sigmaAlbedo", ctypes.c_float), ("sigmaNormal", ctypes.c_float),
                ("cacheInput", ctypes.c_uint8), ("cacheBuffer", ctypes.c_uint8)]


class RmdSvgfFrame(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("color", ctypes.c_void_p),
                ("albedo", ctypes.c_void_p), ("guide", ctypes.c_void_p), ("motion", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("out_rgba8", ctypes.c_void_p)]


class RmdBandLink(ctypes.Structure):
    _fields_ = [("peer_recv", ctypes.c_void_p * 2), ("peer_flag", ctypes.c_void_p * 2), ("recv", ctypes.c_void_p),
                ("flags", ctypes.c_void_p)]


class RmdSvgfParams(ctypes.Structure):
    _fields_ = [("alpha_color", ctypes.c_float), ("alpha_moments", ctypes.c_float),
                ("history_cap", ctypes.c_int32), ("short_history", ctypes.c_int32),
                ("depth_tolerance", ctypes.c_float), ("normal_threshold", ctypes.c_float),
                ("albedo_floor", ctypes.c_float), ("variance_lum_scale", ctypes.c_float)]


_lib = None
_synth = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). raymarchdenoisercuda_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.rmd_filter_baseline.argtypes = [ctypes.POINTER(RmdGBuffer), ctypes.POINTER(RmdFilterParams), P]
    lib.rmd_filter_tiled.argtypes = [ctypes.POINTER(RmdGBuffer), ctypes.POINTER(RmdFilterParams), P]
    lib.rmd_svgf_create.argtypes = [ctypes.POINTER(P), I, I, I]
    lib.rmd_svgf_destroy.argtypes = [P]
    lib.rmd_svgf_reset.argtypes = [P]
    lib.rmd_svgf_frame.argtypes = [P, ctypes.POINTER(RmdSvgfFrame), ctypes.POINTER(RmdFilterParams),
                                   ctypes.POINTER(RmdSvgfParams), P]
    lib.rmd_svgf_frame_host.argtypes = [P, ctypes.POINTER(RmdSvgfFrame), ctypes.POINTER(RmdFilterParams),
                                        ctypes.POINTER(RmdSvgfParams)]
    lib.rmd_svgf_frame_gbuffer.argtypes = [P, ctypes.POINTER(RmdGBuffer), ctypes.POINTER(RmdFilterParams),
                                           ctypes.POINTER(RmdSvgfParams), P, P]
    lib.rmd_svgf_host_wait.argtypes = [P]
    lib.rmd_svgf_last_launch_count.argtypes = [P]
    lib.rmd_svgf_set_profiling.argtypes = [P, I]
    lib.rmd_svgf_get_pass_times.argtypes = [P, ctypes.POINTER(ctypes.c_float), I]
    lib.rmd_svgf_read_plane.argtypes = [P, I, P, ctypes.c_size_t, P]
    lib.rmd_svgf_set_stop_after.argtypes = [P, I]
    lib.rmd_svgf_history_bytes.argtypes = [P, I]
    lib.rmd_svgf_history_bytes.restype = ctypes.c_size_t
    lib.rmd_svgf_history_pack.argtypes = [P, I, I, P, P]
    lib.rmd_svgf_history_unpack.argtypes = [P, I, I, P, P]
    lib.rmd_p2p_alloc.argtypes = [ctypes.POINTER(P), ctypes.c_size_t]
    lib.rmd_p2p_free.argtypes = [P]
    lib.rmd_p2p_export.argtypes = [P, P]
    lib.rmd_p2p_open.argtypes = [P, ctypes.POINTER(P)]
    lib.rmd_p2p_close.argtypes = [P]
    lib.rmd_p2p_signal.argtypes = [P, ctypes.c_ulonglong, P]
    lib.rmd_p2p_wait.argtypes = [P, ctypes.c_ulonglong, P]
    lib.rmd_svgf_band_configure.argtypes = [P, I, I]
    lib.rmd_svgf_band_recv_bytes.argtypes = [P]
    lib.rmd_svgf_band_recv_bytes.restype = ctypes.c_size_t
    lib.rmd_svgf_band_stage.argtypes = [P, ctypes.POINTER(RmdSvgfFrame), ctypes.POINTER(RmdFilterParams),
                                        ctypes.POINTER(RmdSvgfParams), ctypes.POINTER(RmdBandLink), I, P]
    lib.rmd_svgf_band_frame.argtypes = [P, ctypes.POINTER(RmdSvgfFrame), ctypes.POINTER(RmdFilterParams),
                                        ctypes.POINTER(RmdSvgfParams), ctypes.POINTER(RmdBandLink), P]
    lib.rmd_svgf_band_timeouts.argtypes = [P]
    lib.rmd_svgf_band_launch_count.argtypes = [P]
    lib.rmd_svgf_prepare_host.argtypes = [P]
    lib.rmd_svgf_prepare_gbuffer.argtypes = [P, P]
    lib.rmd_debug_clock_probe.argtypes = [P, ctypes.c_uint, P]
    lib.rmd_debug_level_cover.argtypes = [ctypes.c_int] * 12 + [P, P, P]
    lib.rmd_error_string.argtypes = [I]
    lib.rmd_error_string.restype = ctypes.c_char_p
    lib.rmd_sizeof_gbuffer.restype = ctypes.c_size_t
    lib.rmd_sizeof_filter_params.restype = ctypes.c_size_t
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError here means the header and the library diverged
    _lib = lib
    return lib


def load_synth():
    """Host-only synthetic G-buffer generator (workload definition, not the algorithm)."""
    global _synth
    if _synth is not None:
        return _synth
    if not os.path.exists(SYNTH_PATH):
        raise ImportError(f"{SYNTH_PATH} is missing: run __graft_entry__.build()")
    s = ctypes.CDLL(SYNTH_PATH)
    s.rmd_synth_frame.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    s.rmd_synth_frame.restype = None
    _synth = s
    return s
