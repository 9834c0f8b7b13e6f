#!/usr/bin/env python
"""A/B timing of the variance pass (RMD_VAR_DENSE_MIN / RMD_VAR_THREADS) in ONE process: every configuration gets its
own context and runs the same sequence from an empty history; per-frame pass times come from the context's profiling
marks, frame times from CUDA events.  Frames 0..2 are the dense case (every pixel has a short history), the later ones
the sparse steady state.

  python tools/variance_bench.py [--workload 1080p,4k] [--configs 257/256,128/128,0/128] [--frames 12]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORK = {"1080p": (1920, 1080, 0x5EED0001), "4k": (3840, 2160, 0x5EED0002)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1080p,4k")
    ap.add_argument("--configs", default="257/256,257/128,128/128,64/128,0/128", help="dense_min/threads, ...")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200.synth import synth_frame
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
    stream = torch.cuda.current_stream()
    for wl in args.workload.split(","):
        W, H, seed = WORK[wl]
        dev = [[torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in synth_frame(W, H, seed, f)]
               for f in range(args.frames)]
        out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        ref = None
        for cfg in args.configs.split(","):
            dm, nt = cfg.split("/")
            os.environ["RMD_VAR_DENSE_MIN"], os.environ["RMD_VAR_THREADS"] = dm, nt
            ctx = rmd.SvgfContext(W, H, 0)
            var_ms = np.full((args.reps, args.frames), 1e9)
            frame_ms = np.full((args.reps, args.frames), 1e9)
            for rep in range(args.reps):
                # profiled loop: the variance pass alone
                ctx.reset()
                ctx.set_profiling(True)
                for f in range(args.frames):
                    ctx.frame(*dev[f], out, params)
                    torch.cuda.synchronize()
                    var_ms[rep, f] = ctx.pass_times_ms()[1]
                ctx.set_profiling(False)
                # un-profiled loop: whole frames (PDL overlap on)
                ctx.reset()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.frames + 1)]
                ev[0].record(stream)
                for f in range(args.frames):
                    ctx.frame(*dev[f], out, params)
                    ev[f + 1].record(stream)
                torch.cuda.synchronize()
                frame_ms[rep] = [ev[f].elapsed_time(ev[f + 1]) for f in range(args.frames)]
            res = out.clone()
            if ref is None:
                ref = res
            v, fm = var_ms.min(0), frame_ms.min(0)
            print(json.dumps({"workload": wl, "dense_min": int(dm), "threads": int(nt),
                              "variance_us_first3": [round(float(x) * 1e3, 1) for x in v[:3]],
                              "variance_us_steady": round(float(np.median(v[6:])) * 1e3, 1),
                              "frame_us_first3": [round(float(x) * 1e3, 1) for x in fm[:3]],
                              "frame_us_steady": round(float(np.median(fm[6:])) * 1e3, 1),
                              "bit_identical_to_first_config": bool(torch.equal(res, ref))}), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
