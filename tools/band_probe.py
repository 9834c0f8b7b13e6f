#!/usr/bin/env python
"""Fixed cost of the per-level band path on ONE GPU: an interior band of the 8K frame (540 own rows + 2 x 40 halo rows)
whose two neighbours are the band itself (it pushes into its own receive buffer and waits on its own flags, which are
already set when the wait runs: no skew, no NVLink) against the plain frame on a context of the same rows.
Timing only — the halo contents are meaningless."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import raymarchdenoisercuda_b200 as rmd  # noqa: E402
from raymarchdenoisercuda_b200 import shard  # noqa: E402
from raymarchdenoisercuda_b200.synth import synth_frame  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--no-plain", action="store_true", help="skip the plain-frame comparison")
args = ap.parse_args()
W, H = 7680, 4320
band = shard.row_bands(H, args.ranks)[args.ranks // 2]
b = shard.BandedSvgfV2(W, H, band, 0)
b.connect_local(b, b)
frames = []
for f in range(args.frames):
    frames.append([torch.from_numpy(np.ascontiguousarray(b.slice_rows(x)).view(np.int32) if x.dtype == np.uint32
                                    else np.ascontiguousarray(b.slice_rows(x))).cuda() for x in synth_frame(W, H, 0x5EED0003, f)])
out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
stream = torch.cuda.current_stream()


def timed(fn, n):
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


band_us = timed(lambda i: b.frame(*frames[i % args.frames], out, params), args.steps)
# per stage
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
acc = np.zeros(6)
for i in range(args.steps):
    ev[0].record(stream)
    for s in range(6):
        b.stage(s, *frames[i % args.frames], out, params)
        ev[s + 1].record(stream)
    torch.cuda.synchronize()
    acc += [ev[s].elapsed_time(ev[s + 1]) * 1e3 for s in range(6)]
stages = (acc / args.steps).round(1).tolist()
launches = b.launches_per_frame()
if args.no_plain:
    print({"env": {k: v for k, v in os.environ.items() if k.startswith("RMD_")}, "ranks": args.ranks, "band_frame_us": round(band_us, 1),
           "band_stage_us": stages, "band_launches": launches, "timeouts": b.timeouts()})
    sys.exit(0)
ctx = rmd.SvgfContext(W, b.ext_rows, 0)
plain_us = timed(lambda i: ctx.frame(*frames[i % args.frames], out, params), args.steps)
ctx.set_profiling(True)
acc = None
for i in range(args.steps):
    ctx.frame(*frames[i % args.frames], out, params)
    t = np.array(ctx.pass_times_ms()) * 1e3
    acc = t if acc is None else acc + t
print({"ranks": args.ranks, "own_rows": band.rows, "ext_rows": b.ext_rows, "band_frame_us": round(band_us, 1),
       "band_stage_us(temporal+push, variance+L0, L1, L2, L3, L4+unpack)": stages, "band_launches": launches,
       "plain_frame_us_same_rows": round(plain_us, 1), "plain_pass_us": (acc / args.steps).round(1).tolist(),
       "ideal_us(1/ranks of 8K frame)": None, "timeouts": b.timeouts()})
