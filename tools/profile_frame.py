#!/usr/bin/env python
"""Smallest program that runs the hot path: N frames of one synthetic sequence through rmd_svgf_frame, nothing else
(the command line ncu wraps: `ncu ... python tools/profile_frame.py --workload 4k --frames 4`).  7 kernels per frame:
temporal, variance, 5 a-trous levels."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WORK = {"1080p": (1920, 1080, 0x5EED0001), "4k": (3840, 2160, 0x5EED0002), "8k": (7680, 4320, 0x5EED0003)}

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="4k")
ap.add_argument("--frames", type=int, default=4)
args = ap.parse_args()
import torch  # noqa: E402
import raymarchdenoisercuda_b200 as rmd  # noqa: E402
from raymarchdenoisercuda_b200.synth import synth_frame  # noqa: E402

W, H, seed = WORK[args.workload]
ctx = rmd.SvgfContext(W, H, 0)
out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
for f in range(args.frames):
    planes = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in synth_frame(W, H, seed, f)]
    ctx.frame(*planes, out, params)
torch.cuda.synchronize()
print("checksum", float(out[..., :3].double().mean()))
ctx.close()
