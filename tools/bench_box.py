"""Legacy box path: this library vs the reference's own kernels (oracle/_ref/libref_gpu.so, unmodified
src/filter.cu built for sm_100a), radius 2, depth 1, CUDA-event timing, L2 flushed between launches.
Algorithmic bytes: 8 B/px (4 read + 4 written)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raymarchdenoisercuda_b200 as rmd  # noqa: E402
from oracle import pyoracle  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ref = ctypes.CDLL(pyoracle.REF_GPU_LIB) if os.path.exists(pyoracle.REF_GPU_LIB) else None
if ref:
    ref.ref_gpu_launch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.mean(ts))


for W, H in [(500, 500), (1920, 1080), (3840, 2160), (7680, 4320)]:
    g = torch.Generator(device="cuda").manual_seed(1234)
    d_in = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.zeros_like(d_in)
    frame = rmd.GBuffer((W, H), d_in, d_out)
    p = rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=1, radius=2)
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    for st in os.environ.get("BOX_STRIPS", "").split(","):
        if st:
            def fn(st=st):
                os.environ["RMD_BOX_STRIP"] = st
                rmd.filter_tiled(frame, p)
                os.environ.pop("RMD_BOX_STRIP")
            rows.append((f"rmd_filter_tiled strip={st}", fn))
    rows += [("rmd_filter_tiled", lambda: rmd.filter_tiled(frame, p)), ("rmd_filter_baseline", lambda: rmd.filter_baseline(frame, p))]
    if ref and not os.environ.get("BOX_NOREF"):
        rows += [("ref filterKernelBaseline", lambda: ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, 0, 0, s)),
                 ("ref filterKernelTiled cacheInput=0", lambda: ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, 1, 0, s)),
                 ("ref filterKernelTiled cacheInput=1 (wrong output)", lambda: ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, 1, 1, s))]
    for name, fn in rows:
        best, mean = timeit(fn)
        gbs = 8.0 * W * H / (best * 1e-3) / 1e9
        print(f"{W}x{H:5d} {name:52s} best {best*1e3:8.1f} us  mean {mean*1e3:8.1f} us  {W*H/(best*1e-3)/1e6:9.0f} Mpx/s  {gbs:7.0f} GB/s ({gbs/peak*100:4.1f}% of measured HBM peak)")
