"""raymarchdenoisercuda_b200 — B200-native (sm_100a) denoise path behind the filter API of
VictorHerbert/RaymarchDenoiserCuda (reference include/filter.cuh, include/gbuffer.h).

The product is librmd_b200.so (hand-written CUDA kernels + the C ABI declared in
include/rmd_b200.h).  This package is the Python mirror of the reference's host
interface for that path: `FilterParams`, `GBuffer`, the two legacy filter entry
points and the SVGF context.  PyTorch is used only for device memory and streams.
There is no CPU fallback: importing `api` without the built library raises.
"""
from .api import (  # noqa: F401
    FilterParams, FilterType, GBuffer, SvgfContext, SvgfParams, RmdError,
    filter_baseline, filter_tiled, lib_path, version,
)
