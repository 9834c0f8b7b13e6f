"""ctypes front-end of the CPU oracle (oracle/liboracle.so) and of the reference's own
filter.cu built for the host (oracle/_ref/libref_cpu.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")
REF_CPU_LIB = os.path.join(HERE, "_ref", "libref_cpu.so")
REF_GPU_LIB = os.path.join(HERE, "_ref", "libref_gpu.so")

PLANE_TEMPORAL_COLOR, PLANE_TEMPORAL_VAR, PLANE_MOMENTS, PLANE_HISTLEN = 0, 1, 2, 3
PLANE_HISTORY_COLOR, PLANE_GUIDE, PLANE_SLOPE = 4, 5, 6
PLANE_TEMPORAL_COLOR_PRE, PLANE_TEMPORAL_VAR_PRE = 100, 101
_PLANES = {0: (np.float32, 4), 1: (np.float32, 1), 2: (np.float32, 2), 3: (np.uint8, 1), 4: (np.float32, 4),
           5: (np.float32, 4), 6: (np.float32, 1), 100: (np.float32, 4), 101: (np.float32, 1)}


class FrameC(ctypes.Structure):  # RmdSvgfFrame
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("color", ctypes.c_void_p),
                ("albedo", ctypes.c_void_p), ("guide", ctypes.c_void_p), ("motion", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("out_rgba8", ctypes.c_void_p)]


class FilterParamsC(ctypes.Structure):  # RmdFilterParams
    _fields_ = [("type", ctypes.c_int32), ("depth", ctypes.c_int32), ("level", ctypes.c_int32),
                ("radius", ctypes.c_int32), ("sigmaSpace", ctypes.c_float), ("sigmaColor", ctypes.c_float),
                ("sigmaAlbedo", ctypes.c_float), ("sigmaNormal", ctypes.c_float),
                ("cacheInput", ctypes.c_uint8), ("cacheBuffer", ctypes.c_uint8)]


class SvgfParamsC(ctypes.Structure):  # RmdSvgfParams
    _fields_ = [("alpha_color", ctypes.c_float), ("alpha_moments", ctypes.c_float),
                ("history_cap", ctypes.c_int32), ("short_history", ctypes.c_int32),
                ("depth_tolerance", ctypes.c_float), ("normal_threshold", ctypes.c_float),
                ("albedo_floor", ctypes.c_float), ("variance_lum_scale", ctypes.c_float)]


def build():
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


_o = None


def lib():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        _o = ctypes.CDLL(ORACLE_LIB)
        _o.oracle_svgf_create.restype = ctypes.c_void_p
        _o.oracle_svgf_create.argtypes = [ctypes.c_int, ctypes.c_int]
        _o.oracle_svgf_destroy.argtypes = [ctypes.c_void_p]
        _o.oracle_svgf_reset.argtypes = [ctypes.c_void_p]
        _o.oracle_svgf_frame.argtypes = [ctypes.c_void_p, ctypes.POINTER(FrameC), ctypes.POINTER(FilterParamsC),
                                         ctypes.POINTER(SvgfParamsC)]
        _o.oracle_svgf_plane.restype = ctypes.c_void_p
        _o.oracle_svgf_plane.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _o.oracle_box_filter.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5
    return _o


def set_threads(n):
    """OpenMP team size of the oracle; returns the size in effect."""
    lib().oracle_set_threads(int(n))
    return int(lib().oracle_num_threads())


def num_threads():
    return int(lib().oracle_num_threads())


def box_filter(render, radius=2, depth=1, variant="tiled"):
    """render: (H,W,4) uint8.  Returns the denoised (H,W,4) uint8 plane (oracle_box.c)."""
    render = np.ascontiguousarray(render, np.uint8)
    H, W, _ = render.shape
    out = np.zeros_like(render)
    b0, b1 = np.zeros_like(render), np.zeros_like(render)
    rc = lib().oracle_box_filter(render.ctypes.data, out.ctypes.data, b0.ctypes.data, b1.ctypes.data, W, H, radius,
                                 depth, 0 if variant == "baseline" else 1)
    assert rc == 0
    return out


def weighted_filter(render, type=1, radius=2, depth=1, sigmaSpace=0.0, sigmaColor=0.0, sigmaAlbedo=0.0, sigmaNormal=0.0,
                    albedo=None, normal=None):
    """FilterParams::GAUSSIAN (type 1) / CROSS (type 2) on (H,W,4) uint8 planes (oracle_weighted.c)."""
    render = np.ascontiguousarray(render, np.uint8)
    H, W, _ = render.shape
    out = np.zeros_like(render)
    b0, b1 = np.zeros_like(render), np.zeros_like(render)
    al = np.ascontiguousarray(albedo, np.uint8) if albedo is not None else None
    no = np.ascontiguousarray(normal, np.uint8) if normal is not None else None
    fp = FilterParamsC(type, depth, 0, radius, sigmaSpace, sigmaColor, sigmaAlbedo, sigmaNormal, 1, 1)
    o = lib()
    o.oracle_weighted_filter.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.POINTER(FilterParamsC)]
    rc = o.oracle_weighted_filter(render.ctypes.data, out.ctypes.data, b0.ctypes.data, b1.ctypes.data,
                                  al.ctypes.data if al is not None else None, no.ctypes.data if no is not None else None,
                                  W, H, ctypes.byref(fp))
    if rc != 0:
        raise RuntimeError(f"oracle_weighted_filter -> {rc}")
    return out


_r = None


def ref_cpu_available():
    return os.path.exists(REF_CPU_LIB)


def ref_cpu_level(render, radius=2, variant="tiled"):
    """One level of the REFERENCE's own kernel body (src/filter.cu compiled for the host)."""
    global _r
    if _r is None:
        _r = ctypes.CDLL(REF_CPU_LIB)
        _r.ref_cpu_filter_level.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
    render = np.ascontiguousarray(render, np.uint8)
    H, W, _ = render.shape
    out = np.zeros_like(render)
    _r.ref_cpu_filter_level(render.ctypes.data, out.ctypes.data, W, H, radius, 0 if variant == "baseline" else 1)
    return out


def ref_cpu_filter(render, radius=2, depth=1, variant="tiled"):
    """depth host-iterated levels of the reference kernel body (its in-kernel depth loop races)."""
    x = np.ascontiguousarray(render, np.uint8)
    for _ in range(depth):
        x = ref_cpu_level(x, radius, variant)
        if variant == "baseline":
            x[..., 3] = 0  # the reference leaves .w uninitialised (src/filter.cu:50)
    return x


class SvgfOracle:
    def __init__(self, width, height):
        self.W, self.H = width, height
        self._h = lib().oracle_svgf_create(width, height)

    def close(self):
        if getattr(self, "_h", None) and _o is not None:
            _o.oracle_svgf_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        lib().oracle_svgf_reset(self._h)

    def frame(self, color, albedo, guide, motion, depth=5, sigma_z=0.0, sigma_l=0.0, sigma_n=0.0, svgf=None,
              want_rgba8=False):
        color, albedo = np.ascontiguousarray(color), np.ascontiguousarray(albedo)
        guide, motion = np.ascontiguousarray(guide), np.ascontiguousarray(motion)
        out = np.zeros((self.H, self.W, 4), np.float32)
        out8 = np.zeros((self.H, self.W, 4), np.uint8) if want_rgba8 else None
        f = FrameC(self.W, self.H, color.ctypes.data, albedo.ctypes.data, guide.ctypes.data, motion.ctypes.data,
                   out.ctypes.data, out8.ctypes.data if want_rgba8 else None)
        fp = FilterParamsC(3, depth, 0, 2, sigma_z, sigma_l, 0.0, sigma_n, 1, 1)
        sp = SvgfParamsC()
        for k, v in (svgf or {}).items():
            setattr(sp, k, v)
        rc = lib().oracle_svgf_frame(self._h, ctypes.byref(f), ctypes.byref(fp), ctypes.byref(sp))
        if rc != 0:
            raise RuntimeError(f"oracle_svgf_frame -> {rc}")
        return (out, out8) if want_rgba8 else out

    def plane(self, plane):
        dt, ch = _PLANES[plane]
        p = lib().oracle_svgf_plane(self._h, plane)
        n = self.W * self.H * ch
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt).reshape(self.H, self.W, ch).copy()
