#!/usr/bin/env python
"""bench.py — throughput of the full SVGF frame (temporal + variance + 5 a-trous levels).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload 1080p|4k|8k]

Metric (BASELINE.json): Mpixel/s of full SVGF frames.  A "step" is one frame of a synthetic
1-spp G-buffer sequence (camera pan + moving occluders => motion vectors and disocclusions).
  N = 1 : BASELINE.json configs[1] — 1920x1080, one B200.
  N > 1 : one independent 1080p sequence per GPU (BASELINE.json configs[4] sharding; no data-path
          collective; weak scaling), launched by torchrun, one rank per GPU.
`value`  = device-resident throughput (inputs already in HBM; CUDA events on the launch stream).
`e2e`    = the same metric through the host-buffer C-ABI call rmd_svgf_frame_host: pinned host
           G-buffer -> H2D -> frame -> D2H of the float4 result, copies inside the timed region.
`roofline` = dominant kernel (a-trous level): algorithmic bytes / its mean launch time, from CUDA
           events recorded between the passes on the frame's stream (rmd_svgf_set_profiling).
`cpu_baseline` / `--impl reference`: the reference has NO implementation of this path (its kernels
           are an unweighted box filter) and no CPU path at all, so the CPU arm is the
           oracle port (oracle/oracle_svgf.c, OpenMP, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"1080p": (1920, 1080), "4k": (3840, 2160), "8k": (7680, 4320), "tiny": (320, 180)}
DEPTH = 5
# algorithmic bytes per pixel (DESIGN.md "Algorithmic bytes"; one count per distinct plane per kernel)
BYTES_TEMPORAL = 65 + 49
BYTES_LEVEL = 60
BYTES_VARIANCE_STEADY = 0  # 4 B per 32x8 tile
BYTES_FRAME = BYTES_TEMPORAL + BYTES_VARIANCE_STEADY + DEPTH * BYTES_LEVEL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (nvidia-smi equivalent via NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_oracle_rate(W, H, seed, budget_s=15.0, frames=None):
    """Times the CPU oracle (all host threads) on a bounded sample: the top rows of the workload's
    frames.  Returns (Mpixel/s, cores, sample description)."""
    from oracle import pyoracle as po
    from raymarchdenoisercuda_b200.synth import synth_frame
    cores = os.cpu_count() or 1
    hs = min(H, 128)
    orc = po.SvgfOracle(W, hs)
    gen = [tuple(x[:hs] for x in synth_frame(W, H, seed, f)) for f in range(2)]
    t0 = time.perf_counter()
    orc.frame(*gen[0], depth=DEPTH)
    probe = time.perf_counter() - t0  # first frame also takes the all-pixel 7x7 variance path
    n = frames if frames is not None else int(max(2, min(16, budget_s / max(probe, 1e-3))))
    t0 = time.perf_counter()
    for f in range(n):
        orc.frame(*gen[1], depth=DEPTH)
    dt = time.perf_counter() - t0
    orc.close()
    return W * hs * n / dt / 1e6, cores, f"{n} steady-state frames of the top {hs} rows of the {W}x{H} sequence ({W}x{hs} px each)"


def run_reference(args, W, H, rank, world):
    """--impl reference: the CPU arm.  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    from raymarchdenoisercuda_b200.synth import synth_frame
    cores = os.cpu_count() or 1
    hs = min(H, 128)
    orc = po.SvgfOracle(W, hs)
    frames = [tuple(x[:hs] for x in synth_frame(W, H, 0x5EED0001, f)) for f in range(min(args.steps + args.warmup, 8))]
    for i in range(args.warmup):
        orc.frame(*frames[i % len(frames)], depth=DEPTH)
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.frame(*frames[(args.warmup + i) % len(frames)], depth=DEPTH)
    dt = time.perf_counter() - t0
    v = W * hs * args.steps / dt / 1e6
    sample = f"each step = top {hs} rows of a {W}x{H} frame ({W}x{hs} px), oracle port, OpenMP {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "Mpixel/s full SVGF frame", "value": v, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[1]: synthetic {W}x{H} 1-spp G-buffer sequence, temporal+variance+5 a-trous levels",
                   "note": "reference has no SVGF code and no CPU path (SURVEY §0); CPU arm = oracle port"},
        "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_banded(args, W, H, rank, world, local_rank):
    """One W x H sequence, row-banded over the ranks (BASELINE configs[3]): every rank runs the pipeline on its
    band + halo and swaps the history rows of the halo with its neighbours after each frame (NCCL send/recv)."""
    import torch
    import torch.distributed as dist
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    from raymarchdenoisercuda_b200.synth import synth_frame
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup, steps = max(args.warmup, 3), args.steps
    nframes = args.frames or min(steps + warmup, 6)
    band = shard.row_bands(H, world, DEPTH)[rank]
    perlevel = args.scheme == "perlevel"
    if perlevel:
        halo = shard.BAND_HALO
        b = shard.BandedSvgfV2(W, H, band, local_rank)
        if world > 1:
            b.connect_ipc()
    else:
        halo = shard.banded_halo(DEPTH)
        b = shard.BandedSvgf(W, H, band, halo, local_rank)
    dev = []
    for f in range(nframes):
        planes = synth_frame(W, H, 0x5EED0003, f)
        dev.append([torch.from_numpy(np.ascontiguousarray(b.slice_rows(x)).view(np.int32) if x.dtype == np.uint32
                                     else np.ascontiguousarray(b.slice_rows(x))).cuda() for x in planes])
    out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=DEPTH, radius=2)
    stream = torch.cuda.current_stream()

    link = shard.P2PLink(b) if (world > 1 and args.exchange == "p2p" and not perlevel) else None

    def step(i):
        if perlevel:
            b.frame(*dev[i % nframes], out, params)
            return
        b.ctx.frame(*dev[i % nframes], out, params)
        if link is not None:
            link.exchange()
        elif world > 1:
            b.exchange_distributed()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for i in range(steps):
            step(warmup + i)
        e1.record(stream)
        barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    px = W * H
    value = px * steps / (ms * 1e-3) / 1e6
    peak, peak_src = peaks()
    if rank == 0:
        print(json.dumps({
            "metric": "Mpixel/s full SVGF frame", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[3]: synthetic {W}x{H} frame sequence row-banded over {world} GPU(s), "
                                   + (f"per-level halo rows pushed to the neighbour over NVLink peer stores (no recompute of the a-trous "
                                      f"levels; {halo} halo rows per side hold them)" if perlevel else
                                      f"halo {halo} rows recomputed per band, history rows swapped per frame over "
                                      + ("NVLink peer mappings (CUDA IPC, stream-ordered flags)" if link is not None else "NCCL send/recv")),
                       "width": W, "height": H, "levels": DEPTH, "band_rows": band.rows, "ext_rows": b.ext_rows,
                       "frames_resident": nframes,
                       "l2": f"inputs larger than L2: {nframes} distinct frames x {24 * b.ext_rows * W / 1e6:.0f} MB per rank",
                       "parallelism": f"row bands x{world}, neighbour point-to-point only"},
            "roofline": {"bound": "hbm", "achieved": BYTES_FRAME * px / (ms / steps * 1e-3) / 1e9 / world, "peak": peak,
                         "unit": "GB/s", "frac": BYTES_FRAME * px / (ms / steps * 1e-3) / 1e9 / world / peak, "traffic": None,
                         "kernel": "whole frame, per GPU", "peak_source": peak_src},
            "cpu_baseline": None,
            "e2e": None, "gpu_launches": (7 + 2 * DEPTH + 2 if perlevel else b.ctx.last_launch_count()) * steps, "clocks": clk.summary(),
            "scheme": args.scheme,
            "exchange": ("p2p" if perlevel else args.exchange) if world > 1 else None,
            "p2p_wait_timeouts": (b.lib.rmd_p2p_timeouts() if perlevel else (link.timeouts() if link is not None else None)),
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="distinct pre-generated input frames (0 = steps+warmup, max 96)")
    ap.add_argument("--mode", default="sequences", choices=["sequences", "banded"],
                    help="N>1: 'sequences' = one independent sequence per GPU (weak scaling, default); "
                         "'banded' = ONE frame sequence split into row bands over the ranks (strong scaling)")
    ap.add_argument("--scheme", default="perlevel", choices=["perlevel", "halo"],
                    help="banded mode: 'perlevel' = a-trous levels produce only the band's rows and push per-level halo rows "
                         "to the neighbour over NVLink (no recompute); 'halo' = 80 recomputed halo rows + one history swap per frame")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="banded mode: history-row exchange over NVLink peer mappings (CUDA IPC + stream flags) or NCCL send/recv")
    args = ap.parse_args()
    W, H = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, W, H, rank, world)
        return
    if args.mode == "banded":
        run_banded(args, W, H, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200.synth import synth_frame

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(args.warmup, 3)
    steps = args.steps
    nframes = args.frames or min(steps + warmup, 72)
    n_pinned = min(nframes, 8)  # e2e leg: small rotating set of pinned host frames (it is PCIe-bound)
    seed = 0x5EED0001 if world == 1 else 0x5EED0100 + rank  # SURVEY §8d seeds
    px = W * H

    # ---- synthetic sequence: generated on the host, kept in pinned memory, uploaded once ----------
    host, dev = [], []
    for f in range(nframes):
        c, a, g, m = synth_frame(W, H, seed, f)
        planes = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x) for x in (c, a, g, m)]
        if f < n_pinned:
            host.append([p.pin_memory() for p in planes])
            dev.append([p.cuda(non_blocking=True) for p in host[-1]])
        else:
            dev.append([p.cuda() for p in planes])
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=DEPTH, radius=2)
    ctx = rmd.SvgfContext(W, H, local_rank)
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region ---------------------------------------------------------
    for i in range(warmup):
        ctx.frame(*dev[i % nframes], out, params)
    launches_per_frame = ctx.last_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for i in range(steps):
            ctx.frame(*dev[(warmup + i) % nframes], out, params)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * px * steps / (ms * 1e-3) / 1e6

    # ---- per-kernel split with events between the passes (same frames) ----------------------------
    ctx.set_profiling(True)
    acc = None
    for i in range(steps):
        ctx.frame(*dev[(warmup + i) % nframes], out, params)
        t = np.array(ctx.pass_times_ms())
        acc = t if acc is None else acc + t
    ctx.set_profiling(False)
    pass_ms = (acc / steps).tolist()
    level_ms = pass_ms[2:2 + DEPTH]
    dom_ms = float(np.mean(level_ms))
    peak, peak_src = peaks()
    achieved = BYTES_LEVEL * px / (dom_ms * 1e-3) / 1e9
    traffic = None
    warp_inst_1080p = 39.8e6
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        warp_inst_1080p = tj.get("atrous_level_warp_inst_per_launch", warp_inst_1080p)
        if tj.get("workload") == args.workload:
            traffic = tj["atrous_level_dram_bytes_per_launch"]
    # second ceiling: the tap loop is register-file operand-bandwidth bound; FFMA with three
    # distinct operands issues at 0.65 warp-inst/clk/scheduler on B200 (tools/ffma_probe.cu)
    warp_inst = warp_inst_1080p * px / (1920 * 1080)  # ncu smsp__inst_executed.sum per launch, scales with pixels
    sm_clock_hz = 1.965e9
    ipc = warp_inst / (dom_ms * 1e-3 * sm_clock_hz * 148 * 4)
    roofline = {"bound": "hbm", "kernel": "a-trous level: atrous_kernel<1,2,4,8,16> (5 launches/frame, mean)",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "second_ceiling": {"what": "register-file operand bandwidth (3-operand FP32 issue rate)",
                                   "achieved_warp_inst_per_clk_per_scheduler": ipc, "measured_limit": 0.65,
                                   "limit_with_operand_reuse": 0.90, "frac": ipc / 0.65,
                                   "note": "0.65 = FFMA with three distinct registers, 0.90 = two of three reused "
                                           "(a quarter of the tap loop's FFMAs carry .reuse), so frac may pass 1; "
                                           "warp instructions per launch from the 1080p ncu capture, scaled by pixels",
                                   "evidence": "profiles/r1_ffma_probe.txt, profiles/r1_notes.md"},
                "algorithmic_bytes_per_px": BYTES_LEVEL, "launch_ms": dom_ms,
                "frame": {"algorithmic_bytes_per_px": BYTES_FRAME,
                          "achieved": BYTES_FRAME * px / (ms / steps * 1e-3) / 1e9,
                          "frac": BYTES_FRAME * px / (ms / steps * 1e-3) / 1e9 / peak},
                "pass_ms": {"temporal": pass_ms[0], "variance": pass_ms[1], "levels": level_ms}}

    # ---- end to end through the host-buffer C-ABI call --------------------------------------------
    outs_h = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    ctx_h = rmd.SvgfContext(W, H, local_rank)
    for i in range(warmup):
        ctx_h.frame_host(*host[i % n_pinned], outs_h[i & 1], params)
    ctx_h.host_wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        ctx_h.frame_host(*host[(warmup + i) % n_pinned], outs_h[i & 1], params)
    ctx_h.host_wait()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * px * steps / e2e_s / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": 24 * px,
           "d2h_bytes_per_step": 16 * px, "ms_per_step": e2e_s / steps * 1e3,
           "api": "rmd_svgf_frame_host (pinned host planes, 3-stream copy/compute overlap)"}
    checksum = float(outs_h[(steps - 1) & 1][..., :3].double().mean())

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_oracle_rate(W, H, seed)
            cpu = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "Mpixel/s full SVGF frame", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: synthetic {W}x{H} 1-spp G-buffer sequence (albedo/normal/depth/motion), "
                                   f"full SVGF temporal+variance+{DEPTH} a-trous levels"
                                   + ("" if world == 1 else f"; one independent sequence per GPU x{world} (configs[4])"),
                       "width": W, "height": H, "levels": DEPTH, "frames_resident": nframes,
                       "l2": f"inputs larger than L2: {nframes} distinct frames x {24 * px / 1e6:.0f} MB cycle through HBM; "
                             f"internal planes {BYTES_FRAME * px / 1e6:.0f} MB/frame of traffic",
                       "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, one sequence stream each, no collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_frame * steps,
            "launches_per_frame": launches_per_frame, "clocks": clk.summary(), "checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
