"""Python mirror of the reference's host interface for the denoise path.

Names and argument meaning follow the reference (`FilterParams` include/filter.cuh:11-23,
`GBuffer` include/gbuffer.h:6-14, `filterKernelBaseline` / `filterKernelTiled`
include/filter.cuh:25-26).  Every call goes through the C ABI of librmd_b200.so with plain
device pointers; torch tensors only provide the memory and the current stream.
"""
import ctypes
import enum

import numpy as np
import torch

from . import _lib
from ._lib import RmdFilterParams, RmdGBuffer, RmdSvgfFrame, RmdSvgfParams


class RmdError(RuntimeError):
    """Raised for any non-zero return of the C ABI (the reference's harness catches
    std::runtime_error, src/test.cu:40-42; this is the Python analogue)."""

    def __init__(self, code):
        self.code = code
        super().__init__(f"rmd error {code}: {_lib.load().rmd_error_string(code).decode()}")


def _check(rc):
    if rc != 0:
        raise RmdError(rc)


class FilterType(enum.IntEnum):  # reference include/filter.cuh:12
    AVERAGE = 0
    GAUSSIAN = 1
    CROSS = 2
    WAVELET = 3


class FilterParams:
    """reference include/filter.cuh:11-23 (defaults: cacheInput = cacheBuffer = true, rest 0)."""

    def __init__(self, type=FilterType.AVERAGE, depth=0, level=0, radius=0, sigmaSpace=0.0, sigmaColor=0.0,
                 sigmaAlbedo=0.0, sigmaNormal=0.0, cacheInput=True, cacheBuffer=True):
        self.type, self.depth, self.level, self.radius = int(type), depth, level, radius
        self.sigmaSpace, self.sigmaColor, self.sigmaAlbedo, self.sigmaNormal = sigmaSpace, sigmaColor, sigmaAlbedo, sigmaNormal
        self.cacheInput, self.cacheBuffer = cacheInput, cacheBuffer

    def c(self):
        return RmdFilterParams(self.type, self.depth, self.level, self.radius, self.sigmaSpace, self.sigmaColor,
                               self.sigmaAlbedo, self.sigmaNormal, int(self.cacheInput), int(self.cacheBuffer))


class SvgfParams:
    def __init__(self, **kw):
        self.kw = kw

    def c(self):
        p = RmdSvgfParams()
        for k, v in self.kw.items():
            setattr(p, k, v)
        return p


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("planes must be contiguous CUDA tensors")
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


class GBuffer:
    """Non-owning view of device planes, reference include/gbuffer.h:6-14.  Planes are
    uint8 tensors of shape (H, W, 4) (uchar4, row-major, pitch 4*W)."""

    def __init__(self, shape, render, denoised, normal=None, albedo=None, buffer=(None, None)):
        self.shape = tuple(shape)  # (W, H) like the reference's int2 shape
        self.render, self.denoised, self.normal, self.albedo, self.buffer = render, denoised, normal, albedo, tuple(buffer)

    def c(self):
        g = RmdGBuffer()
        g.width, g.height = self.shape
        g.render, g.denoised = _ptr(self.render), _ptr(self.denoised)
        g.normal, g.albedo = _ptr(self.normal), _ptr(self.albedo)
        g.buffer[0], g.buffer[1] = _ptr(self.buffer[0]), _ptr(self.buffer[1])
        return g


def filter_baseline(frame: GBuffer, params: FilterParams, stream=None):
    """Drop-in for `filterKernelBaseline<<<grid, block, smem>>>(frame, params)` (reference src/test.cu:73-75)."""
    _check(_lib.load().rmd_filter_baseline(ctypes.byref(frame.c()), ctypes.byref(params.c()), _stream_ptr(stream)))


def filter_tiled(frame: GBuffer, params: FilterParams, stream=None):
    """Drop-in for `filterKernelTiled<<<grid, block, smem>>>(frame, params)` (reference src/test.cu:85-87)."""
    _check(_lib.load().rmd_filter_tiled(ctypes.byref(frame.c()), ctypes.byref(params.c()), _stream_ptr(stream)))


PLANES = {  # id -> (numpy dtype, channels)
    0: (np.float32, 4), 1: (np.float32, 1), 2: (np.float32, 2), 3: (np.uint8, 1), 4: (np.float32, 4),
    5: (np.float32, 4), 6: (np.float32, 1),
}


class SvgfContext:
    """Per-sequence SVGF state (history planes) — the owner the reference sketched as
    `CudaGBuffer` (include/gbuffer.h:20-33) and never implemented."""

    def __init__(self, width, height, device=0):
        self._lib = _lib.load()
        self.width, self.height, self.device = width, height, device
        h = ctypes.c_void_p()
        _check(self._lib.rmd_svgf_create(ctypes.byref(h), width, height, device))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rmd_svgf_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        _check(self._lib.rmd_svgf_reset(self._h))

    def set_stop_after(self, stage):
        _check(self._lib.rmd_svgf_set_stop_after(self._h, stage))

    def frame(self, color, albedo, guide, motion, out, params: FilterParams, svgf: SvgfParams = None, out_rgba8=None,
              stream=None):
        """color (H,W,4) f16 | albedo (H,W,4) u8 | guide (H,W,2) i32/u32 | motion (H,W,2) f16 | out (H,W,4) f32."""
        f = RmdSvgfFrame(self.width, self.height, _ptr(color), _ptr(albedo), _ptr(guide), _ptr(motion), _ptr(out),
                         _ptr(out_rgba8))
        sp = svgf.c() if svgf is not None else None
        _check(self._lib.rmd_svgf_frame(self._h, ctypes.byref(f), ctypes.byref(params.c()),
                                        ctypes.byref(sp) if sp is not None else None, _stream_ptr(stream)))

    def frame_gbuffer(self, frame: "GBuffer", params: FilterParams, svgf: SvgfParams = None, out=None, stream=None):
        """SVGF on the reference's own `GBuffer` (RGBA8 render/albedo/normal in, RGBA8 `denoised` out), the call
        `filterKernel*(GBuffer, FilterParams{.type = WAVELET})` was reserved for (reference include/filter.cuh:12-19)."""
        sp = svgf.c() if svgf is not None else None
        _check(self._lib.rmd_svgf_frame_gbuffer(self._h, ctypes.byref(frame.c()), ctypes.byref(params.c()),
                                                ctypes.byref(sp) if sp is not None else None, _ptr(out),
                                                _stream_ptr(stream)))

    def frame_host(self, color, albedo, guide, motion, out, params: FilterParams, svgf: SvgfParams = None, out_rgba8=None):
        """Same with HOST tensors/arrays (pinned for true overlap); asynchronous — call host_wait()."""
        def hp(a):
            if a is None:
                return None
            return ctypes.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        f = RmdSvgfFrame(self.width, self.height, hp(color), hp(albedo), hp(guide), hp(motion), hp(out), hp(out_rgba8))
        sp = svgf.c() if svgf is not None else None
        _check(self._lib.rmd_svgf_frame_host(self._h, ctypes.byref(f), ctypes.byref(params.c()),
                                             ctypes.byref(sp) if sp is not None else None))

    def host_wait(self):
        _check(self._lib.rmd_svgf_host_wait(self._h))

    def set_profiling(self, enable=True):
        _check(self._lib.rmd_svgf_set_profiling(self._h, int(enable)))

    def pass_times_ms(self):
        """[temporal, variance, level0, ...] GPU milliseconds of the last frame (needs set_profiling)."""
        buf = (ctypes.c_float * 16)()
        n = self._lib.rmd_svgf_get_pass_times(self._h, buf, 16)
        if n < 0:
            raise RmdError(n)
        return [buf[i] for i in range(n)]

    def history_bytes(self, nrows):
        return self._lib.rmd_svgf_history_bytes(self._h, nrows)

    def history_pack(self, row_begin, nrows, buf, stream=None):
        """Packs the frame-to-frame state of rows [row_begin, row_begin+nrows) into the uint8 CUDA tensor `buf`."""
        _check(self._lib.rmd_svgf_history_pack(self._h, row_begin, nrows, self._buf(buf, nrows), _stream_ptr(stream)))

    def history_unpack(self, row_begin, nrows, buf, stream=None):
        _check(self._lib.rmd_svgf_history_unpack(self._h, row_begin, nrows, self._buf(buf, nrows), _stream_ptr(stream)))

    def _buf(self, buf, nrows):
        """uint8 CUDA tensor or a raw device address (int) — e.g. a peer-mapped buffer of another rank."""
        if isinstance(buf, int):
            return ctypes.c_void_p(buf)
        assert buf.is_cuda and buf.numel() * buf.element_size() >= self.history_bytes(nrows)
        return _ptr(buf)

    def last_launch_count(self):
        return self._lib.rmd_svgf_last_launch_count(self._h)

    def read_plane(self, plane, stream=None):
        dt, ch = PLANES[plane]
        a = np.empty((self.height, self.width, ch), dtype=dt)
        _check(self._lib.rmd_svgf_read_plane(self._h, plane, a.ctypes.data, a.nbytes, _stream_ptr(stream)))
        return a


def lib_path():
    return _lib.LIB_PATH


def version():
    return _lib.load().rmd_version()
