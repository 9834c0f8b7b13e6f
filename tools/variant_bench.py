#!/usr/bin/env python
"""A/B timing of the a-trous tile-kernel variants (RMD_ATROUS_VARIANT) and of PDL (RMD_PDL) in ONE process:
every configuration gets its own context; frame time from CUDA events around K frames (no per-pass marks, so
PDL can overlap), per-pass times from a second profiled loop.

  python tools/variant_bench.py [--workload 1080p,4k] [--variants 0,1,2,3,4,5] [--steps 30]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORK = {"1080p": (1920, 1080, 0x5EED0001), "4k": (3840, 2160, 0x5EED0002), "8k": (7680, 4320, 0x5EED0003)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1080p,4k")
    ap.add_argument("--variants", default="0,1,2,3,4,5")
    ap.add_argument("--pdl", default="1,0")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--extra-env", default="", help="KEY=VAL,KEY=VAL applied to every configuration")
    args = ap.parse_args()
    import torch
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200.synth import synth_frame
    for kv in filter(None, args.extra_env.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
    stream = torch.cuda.current_stream()
    for wl in args.workload.split(","):
        W, H, seed = WORK[wl]
        dev = []
        for f in range(args.frames):
            dev.append([torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
                        for x in synth_frame(W, H, seed, f)])
        out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        ref = None
        for variant in args.variants.split(","):
            for pdl in args.pdl.split(","):
                # "11@370" = variant 11 with an L2 prefetch look-ahead of 370 tiles
                # "17+RMD_VAR_REVERSE=0" = variant 17 with that environment variable set for this configuration only
                variant_id, *extra = variant.split("+")
                for k in [k for k in os.environ if k in getattr(main, "_extra_keys", ())]:
                    os.environ.pop(k)
                main._extra_keys = tuple(kv.split("=")[0] for kv in extra)
                for kv in extra:
                    k, v = kv.split("=")
                    os.environ[k] = v
                vid, _, ahead = variant_id.partition("@")
                os.environ["RMD_ATROUS_VARIANT"] = vid.replace("/", ",")
                if ahead:
                    os.environ["RMD_ATROUS_PREFETCH"] = ahead
                else:
                    os.environ.pop("RMD_ATROUS_PREFETCH", None)
                os.environ["RMD_PDL"] = pdl
                ctx = rmd.SvgfContext(W, H, 0)
                for i in range(args.warmup):
                    ctx.frame(*dev[i % args.frames], out, params)
                torch.cuda.synchronize()
                best = 1e9
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for i in range(args.steps):
                        ctx.frame(*dev[(args.warmup + i) % args.frames], out, params)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / args.steps)
                chk = out[..., :3].double().mean().item()
                ctx.set_profiling(True)
                acc = None
                for i in range(args.steps):
                    ctx.frame(*dev[(args.warmup + i) % args.frames], out, params)
                    t = np.array(ctx.pass_times_ms())
                    acc = t if acc is None else acc + t
                ctx.set_profiling(False)
                p = (acc / args.steps * 1e3).round(1).tolist()
                o = out.clone()
                if ref is None:
                    ref = o
                diff = float((o[..., :3] - ref[..., :3]).abs().max())
                ctx.close()
                print(json.dumps({"workload": wl, "variant": variant, "pdl": int(pdl), "frame_us": round(best * 1e3, 1),
                                  "temporal_us": p[0], "variance_us": p[1], "levels_us": p[2:], "checksum": chk,
                                  "maxdiff_vs_first": diff}), flush=True)
        del dev, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
