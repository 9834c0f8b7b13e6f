#!/usr/bin/env python
"""bench.py — throughput of the full SVGF frame (temporal + variance + 5 a-trous levels).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload 8k|4k|1080p]

Metric (BASELINE.json): Mpixel/s of full SVGF frames.  A "step" is one frame of a synthetic 1-spp G-buffer sequence
(camera pan + moving occluders => motion vectors and disocclusions), generated on the host, resident in HBM.

  N = 1 : the 7680x4320 frame on one B200 — the single-GPU denominator of BASELINE.json configs[3] and the largest
          configuration of the metric ("1080p/4K/8K").  Sub-records of the same JSON line carry configs[1]
          (1920x1080), configs[2] (3840x2160 x 64 frames: steady state AND worst frame) and the legacy box path next
          to the reference's own kernels rebuilt for sm_100a (`reference_gpu_box`).
  N > 1 : the SAME 7680x4320 sequence split into N row bands, one rank per GPU (torchrun), per-level halo rows pushed
          to the neighbour over NVLink peer stores (configs[3]; strong scaling; no collective on the data path).
          Sub-record `replicas`: one independent 1080p sequence per GPU (configs[4], weak scaling).
`value`    device-resident throughput (inputs already in HBM; CUDA events on the launch stream, max over ranks).
`e2e`      the same metric through the host-buffer entry point: pinned host G-buffer -> H2D -> frame -> D2H of the
           RGBA8 result (the reference's `denoised` format), copies inside the timed region; `e2e_fp32` = same with
           the float4 result.
`roofline` dominant kernel (a-trous level): algorithmic bytes / its mean launch time, from CUDA events recorded
           between the passes on the frame's stream (rmd_svgf_set_profiling).
`cpu_baseline` / `--impl reference`: the reference has NO implementation of this path (its kernels are an unweighted
           box filter) and no CPU path at all, so the CPU arm is the oracle port (oracle/oracle_svgf.c,
           OpenMP, thread count set and reported explicitly) on whole frames of the same workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"1080p": (1920, 1080, 0x5EED0001, "configs[1]"), "4k": (3840, 2160, 0x5EED0002, "configs[2]"),
             "8k": (7680, 4320, 0x5EED0003, "configs[3]"), "tiny": (320, 180, 0x5EED0001, "configs[1] (reduced to 320x180)")}
DEPTH = 5
# algorithmic bytes per pixel (DESIGN.md "Algorithmic bytes"; one count per distinct plane per kernel)
BYTES_TEMPORAL = 65 + 49
BYTES_LEVEL = 60
BYTES_VARIANCE_STEADY = 0  # 4 B per 32x8 tile
BYTES_FRAME = BYTES_TEMPORAL + BYTES_VARIANCE_STEADY + DEPTH * BYTES_LEVEL
IN_BYTES_PX = 24  # colour 8 + albedo 4 + guide 8 + motion 4


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(name, world=1, mode="single"):
    """The `config` object: identical for the product arm and the reference arm of the same command line."""
    W, H, _, cfg = WORKLOADS[name]
    c = {"workload": f"{cfg}: synthetic {W}x{H} 1-spp G-buffer sequence (albedo/normal/depth/motion, camera pan + moving "
                     f"occluders), full SVGF temporal+variance+{DEPTH} a-trous levels",
         "width": W, "height": H, "levels": DEPTH,
         "l2": f"inputs larger than L2: distinct frames of {IN_BYTES_PX * W * H / 1e6:.0f} MB cycle through HBM; "
               f"internal planes {BYTES_FRAME * W * H / 1e6:.0f} MB/frame of traffic"}
    if world == 1:
        c["parallelism"] = "1 GPU"
    elif mode == "banded":
        c["parallelism"] = f"{world} GPUs: row bands x{world}, per-level halo rows over NVLink peer stores, neighbour point-to-point only"
    else:
        c["parallelism"] = f"{world} GPUs, one independent sequence each, no collective"
    return c


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (nvidia-smi equivalent via NVML).

    Used at N = 1 only, where polling costs nothing measurable.  Under torchrun an NVML query inside the timed region
    stalls the band path's stream-ordered cross-GPU hand-offs (one poller per rank: 0.87-0.98 ms per 8K band frame at
    N = 8 against 0.78 ms without; rank 0 alone at 50 Hz: still 2.65 against 2.38 ms at N = 2), so N > 1 uses
    DeviceClockProbe below instead (profiles/r2_scaling.md)."""

    def __init__(self, index, enabled=True, period_s=0.005):
        self.samples, self.reasons, self.stop, self.max_mhz, self.period = [], set(), False, None, period_s
        try:
            if os.environ.get("RMD_BENCH_NO_SAMPLER") == "1" or not enabled:
                raise RuntimeError("disabled")
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
            time.sleep(self.period)

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
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


class DeviceClockProbe:
    """SM clock during the timed region WITHOUT NVML: a one-warp kernel on its own stream spins ~200 us in the middle of
    the region and reports %clock64 cycles per %globaltimer nanosecond (rmd_debug_clock_probe).  Throttle reasons and
    the maximum clock are read through NVML right AFTER the region has been synchronised.  Used for N > 1: one NVML
    query inside the timed region stalls the band path by ~2 ms (measured: 2.65 ms per 8K frame at N = 2 with three
    queries from rank 0 in the region, 2.38 ms with none; 0.87-0.98 vs 0.78 ms at N = 8; profiles/r2_scaling.md)."""

    def __init__(self, device):
        import torch
        from raymarchdenoisercuda_b200 import _lib
        self.torch, self.lib, self.device = torch, _lib.load(), device
        self.stream = torch.cuda.Stream()
        self.out = torch.zeros(2, dtype=torch.int64, device="cuda")

    def fire(self):
        self.lib.rmd_debug_clock_probe(ctypes.c_void_p(self.out.data_ptr()), 200, ctypes.c_void_p(self.stream.cuda_stream))

    def summary(self):
        self.torch.cuda.synchronize()
        cyc, ns = [int(v) for v in self.out.cpu().tolist()]
        res = {"sm_mhz": round(1000.0 * cyc / ns, 1) if ns > 0 else None, "sm_max_mhz": None, "reasons": ["unavailable"],
               "samples": 1 if ns > 0 else 0,
               "method": "device: %clock64 / %globaltimer over a 200 us one-warp spin inside the timed region (rank 0); "
                         "throttle reasons and max clock from NVML immediately after the region (an NVML query inside "
                         "it stalls the cross-GPU hand-offs)"}
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            res["sm_max_mhz"] = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            res["sm_mhz_after"] = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80}
            res["reasons"] = sorted(k for k, bit in names.items() if r & bit)
        except Exception as e:  # noqa: BLE001
            res["nvml_error"] = str(e)
        return res


# ---------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(W, H, seed, steps, warmup, budget_s):
    """Times the CPU oracle on frames of the workload with an explicit OpenMP team.  Whole frames when the run fits
    `budget_s`, else the largest top crop (multiple of 16 rows) that does.  Returns (Mpixel/s, ms/step, threads,
    sample description, whole_frame flag)."""
    from oracle import pyoracle as po
    from raymarchdenoisercuda_b200.synth import synth_frame
    threads = po.set_threads(host_threads())
    probe_rows = min(H, 128)
    planes = synth_frame(W, H, seed, 0)
    orc = po.SvgfOracle(W, probe_rows)
    crop = [x[:probe_rows] for x in planes]
    orc.frame(*crop, depth=DEPTH)  # first frame: every pixel takes the 7x7 variance pass
    t0 = time.perf_counter()
    orc.frame(*crop, depth=DEPTH)
    per_row = (time.perf_counter() - t0) / probe_rows
    orc.close()
    n = max(1, steps + warmup)
    rows = H if per_row * H * n <= budget_s else max(16, min(H, int(budget_s / (per_row * n)) // 16 * 16))
    nf = min(n, 4)
    frames = [[x[:rows] for x in (planes if f == 0 else synth_frame(W, H, seed, f))] for f in range(nf)]
    orc = po.SvgfOracle(W, rows)
    for i in range(warmup):
        orc.frame(*frames[i % nf], depth=DEPTH)
    t0 = time.perf_counter()
    for i in range(steps):
        orc.frame(*frames[(warmup + i) % nf], depth=DEPTH)
    dt = time.perf_counter() - t0
    orc.close()
    whole = rows == H
    sample = (f"each step = {'the whole' if whole else f'the top {rows} rows of a'} {W}x{H} frame ({W}x{rows} px), "
              f"oracle port (oracle/oracle_svgf.c), OpenMP {threads} threads")
    return W * rows * steps / dt / 1e6, dt / steps * 1e3, threads, sample, whole


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  Under torchrun only rank 0 works (the other ranks exit 0 without work)."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its children: undo that before libgomp initialises
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    W, H, seed, _ = WORKLOADS[args.workload]
    v, ms, threads, sample, whole = cpu_arm(W, H, seed, args.steps, args.warmup, args.cpu_budget)
    mode = "banded" if world > 1 else "single"
    print(json.dumps({
        "impl": "reference", "metric": "Mpixel/s full SVGF frame", "value": v, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world, mode),
        "note": "the reference has no SVGF code and no CPU path (SURVEY §0): the CPU arm is the oracle port; GPUs idle",
        "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample,
                         "whole_frames": whole},
        "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------
# product arm helpers
# ---------------------------------------------------------------------------------------------------------------
def _t(x):
    import torch
    return torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x)


def single_gpu_run(name, steps, warmup, device, max_frames, per_frame=False, with_e2e=True, from_reset=False, sampler=None):
    """One sequence on one GPU.  Returns a dict with the device-resident timing, the per-pass split and (optionally)
    the end-to-end numbers.  per_frame: one event per frame (worst-frame report); from_reset: no warm-up frames,
    the timed sequence starts with an empty history (configs[2])."""
    import torch
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200.synth import synth_frame
    W, H, seed, _ = WORKLOADS[name]
    px = W * H
    nframes = min(steps + (0 if from_reset else warmup), max_frames)
    n_pinned = min(nframes, 4 if px > 8e6 else 8) if with_e2e else 0
    host, dev = [], []
    for f in range(nframes):
        planes = [_t(x) for x in synth_frame(W, H, seed, f)]
        if f < n_pinned:
            host.append([p.pin_memory() for p in planes])
            dev.append([p.cuda(non_blocking=True) for p in host[-1]])
        else:
            dev.append([p.cuda() for p in planes])
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=DEPTH, radius=2)
    ctx = rmd.SvgfContext(W, H, device)
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()
    res = {"width": W, "height": H, "frames_resident": nframes}

    if not from_reset:
        for i in range(warmup):
            ctx.frame(*dev[i % nframes], out, params)
    res["launches_per_frame"] = None
    torch.cuda.synchronize()
    base = 0 if from_reset else warmup
    if per_frame:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for i in range(steps):
            ctx.frame(*dev[(base + i) % nframes], out, params)
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        frame_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        res["frame_ms"] = frame_ms
        ms = float(sum(frame_ms))
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.__enter__()      # clocks are sampled during the timed region only
        e0.record(stream)
        for i in range(steps):
            ctx.frame(*dev[(base + i) % nframes], out, params)
        e1.record(stream)
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.__exit__()
        ms = e0.elapsed_time(e1)
    res["launches_per_frame"] = ctx.last_launch_count()
    res["ms_total"] = ms
    res["ms_per_step"] = ms / steps
    res["value"] = px * steps / (ms * 1e-3) / 1e6

    # per-kernel split with events between the passes (same frames; the marks switch PDL overlap off)
    ctx.set_profiling(True)
    acc = None
    n_prof = min(steps, 20)
    for i in range(n_prof):
        ctx.frame(*dev[(base + steps + i) % nframes], out, params)
        t = np.array(ctx.pass_times_ms())
        acc = t if acc is None else acc + t
    ctx.set_profiling(False)
    pass_ms = (acc / n_prof).tolist()
    res["pass_ms"] = {"temporal": pass_ms[0], "variance": pass_ms[1], "levels": pass_ms[2:2 + DEPTH]}
    res["checksum_device"] = float(out[..., :3].double().mean())
    ctx.close()

    if with_e2e:
        for key, want_f32 in (("e2e", False), ("e2e_fp32", True)):
            ctx_h = rmd.SvgfContext(W, H, device)
            outs_f = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)] if want_f32 else [None, None]
            outs_8 = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
            n_e2e = min(steps, 20)
            for i in range(max(3, min(warmup, 6))):
                ctx_h.frame_host(*host[i % n_pinned], outs_f[i & 1], params, out_rgba8=outs_8[i & 1])
            ctx_h.host_wait()
            t0 = time.perf_counter()
            for i in range(n_e2e):
                ctx_h.frame_host(*host[(warmup + i) % n_pinned], outs_f[i & 1], params, out_rgba8=outs_8[i & 1])
            ctx_h.host_wait()
            dt = time.perf_counter() - t0
            d2h = (16 + 4) * px if want_f32 else 4 * px
            res[key] = {"value": px * n_e2e / dt / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": IN_BYTES_PX * px,
                        "d2h_bytes_per_step": d2h, "ms_per_step": dt / n_e2e * 1e3, "steps": n_e2e,
                        "api": "rmd_svgf_frame_host (pinned host planes, 3-stream copy/compute overlap), result = "
                               + ("float4 radiance + RGBA8" if want_f32 else "RGBA8 `denoised` (reference include/gbuffer.h:10)"),
                        "checksum": float(outs_8[(n_e2e - 1) & 1][..., :3].double().mean())}
            ctx_h.close()
    del dev, host
    torch.cuda.empty_cache()
    return res


def level_roofline(res, name):
    """The `roofline` object for the dominant kernel (a-trous level) of a single-GPU run."""
    W, H = res["width"], res["height"]
    px = W * H
    peak, peak_src = peaks()
    level_ms = res["pass_ms"]["levels"]
    dom_ms = float(np.mean(level_ms))
    achieved = BYTES_LEVEL * px / (dom_ms * 1e-3) / 1e9
    traffic, warp_inst, evidence = None, None, None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        ent = tj.get(name)
        if ent:
            traffic = ent.get("atrous_level_dram_bytes_per_launch")
            warp_inst = ent.get("atrous_level_warp_inst_per_launch")
            evidence = ent.get("evidence")
    out = {"bound": "hbm", "kernel": "a-trous level: atrous_kernel<S = 1,2,4,8,16> (5 launches/frame, mean)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
           "traffic_evidence": evidence, "peak_source": peak_src, "algorithmic_bytes_per_px": BYTES_LEVEL,
           "launch_ms": dom_ms,
           "frame": {"algorithmic_bytes_per_px": BYTES_FRAME,
                     "achieved": BYTES_FRAME * px / (res["ms_per_step"] * 1e-3) / 1e9,
                     "frac": BYTES_FRAME * px / (res["ms_per_step"] * 1e-3) / 1e9 / peak},
           "pass_ms": res["pass_ms"]}
    if warp_inst:
        ipc = warp_inst / (dom_ms * 1e-3 * 1.965e9 * 148 * 4)
        out["second_ceilings"] = {
            "what": "the level kernel is bound by instruction issue, not by HBM: four near-saturated limits at once "
                    "(DESIGN.md §6): issue slots, the XU pipe (2 MUFU per tap), shared-memory wavefronts, register-file "
                    "operand ports (dispatch stalls)",
            "warp_inst_per_clk_per_scheduler": ipc, "evidence": "profiles/r2_atrous_stalls.md"}
    return out


def box_vs_reference_gpu(device):
    """Legacy box path (filterKernelBaseline / filterKernelTiled, radius 2, depth 1) at 7680x4320: this library next
    to the reference's own kernels, unmodified src/filter.cu rebuilt for sm_100a (oracle/_ref/libref_gpu.so), same GPU,
    L2 flushed between launches, CUDA events, best of 10."""
    import torch
    import raymarchdenoisercuda_b200 as rmd
    from oracle import pyoracle
    W, H = 7680, 4320
    g = torch.Generator(device="cuda").manual_seed(1234)
    d_in = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.zeros_like(d_in)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    frame = rmd.GBuffer((W, H), d_in, d_out)
    p = rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=1, radius=2)
    s = torch.cuda.current_stream().cuda_stream
    peak, _ = peaks()

    def best_us(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return min(ts)

    rec = {"workload": f"{W}x{H} RGBA8, radius 2, depth 1, L2 flushed between launches, best of 10",
           "algorithmic_bytes_per_px": 8}
    rows = [("rmd_filter_tiled", lambda: rmd.filter_tiled(frame, p)), ("rmd_filter_baseline", lambda: rmd.filter_baseline(frame, p))]
    if os.path.exists(pyoracle.REF_GPU_LIB):
        ref = ctypes.CDLL(pyoracle.REF_GPU_LIB)
        ref.ref_gpu_launch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
        rows += [("reference filterKernelTiled (cacheInput=0), sm_100a rebuild",
                  lambda: ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, 1, 0, s)),
                 ("reference filterKernelBaseline, sm_100a rebuild",
                  lambda: ref.ref_gpu_launch(d_in.data_ptr(), d_out.data_ptr(), None, None, W, H, 2, 1, 0, 0, s))]
    else:
        rec["reference"] = "oracle/_ref/libref_gpu.so missing (build it where /root/reference exists: make -C oracle ref)"
    for name, fn in rows:
        us = best_us(fn)
        gbs = 8.0 * W * H / (us * 1e-6) / 1e9
        rec[name] = {"us": us, "mpixel_s": W * H / us, "gb_s": gbs, "frac_of_hbm_peak": gbs / peak}
    del d_in, d_out, flush
    torch.cuda.empty_cache()
    return rec


# ---------------------------------------------------------------------------------------------------------------
# N > 1: one frame sequence in row bands
# ---------------------------------------------------------------------------------------------------------------
def banded_run(args, name, rank, world, local_rank, steps, warmup):
    import torch
    import torch.distributed as dist
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    from raymarchdenoisercuda_b200.synth import synth_frame
    W, H, seed, _ = WORKLOADS[name]
    nframes = min(steps + warmup + 4, 36 if W * H > 8e6 else 72)   # + the 4 set-up frames: the sequence never wraps
    band = shard.row_bands(H, world, DEPTH)[rank]
    perlevel = args.scheme == "perlevel"
    if perlevel:
        b = shard.BandedSvgfV2(W, H, band, local_rank)
        if world > 1:
            b.connect_ipc()
    else:
        b = shard.BandedSvgf(W, H, band, shard.banded_halo(DEPTH), local_rank)
    n_pinned = min(nframes, 4)
    host, dev = [], []
    for f in range(nframes):
        planes = [_t(np.ascontiguousarray(b.slice_rows(x))) for x in synth_frame(W, H, seed, f)]
        if f < n_pinned:
            host.append([p.pin_memory() for p in planes])
            dev.append([p.cuda(non_blocking=True) for p in host[-1]])
        else:
            dev.append([p.cuda() for p in planes])
    out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
    out8 = torch.empty((b.ext_rows, W, 4), dtype=torch.uint8, device="cuda")
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=DEPTH, radius=2)
    stream = torch.cuda.current_stream()
    link = shard.P2PLink(b) if (world > 1 and args.exchange == "p2p" and not perlevel) else None

    def step(planes, o8=None):
        if perlevel:
            b.frame(*planes, out, params, out_rgba8=o8)
            return
        b.ctx.frame(*planes, out, params, out_rgba8=o8)
        if link is not None:
            link.exchange()
        elif world > 1:
            b.exchange_distributed()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()  # every rank has its buffers mapped and its data resident before the first flag wait
    # Set-up, not warm-up: the first frames after the CUDA-IPC mappings are created run 2-4x slower (first touch of the
    # peer-mapped receive buffers over NVLink, profiles/r1_scaling.md); they are paid once per sequence, so the
    # sequence is started before the W warm-up steps the command line asks for.
    for i in range(4):
        step(dev[i % nframes])
    barrier()
    for i in range(warmup):
        step(dev[(4 + i) % nframes])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe = DeviceClockProbe(local_rank) if rank == 0 else None
    e0.record(stream)
    for i in range(steps):
        if probe is not None and i == steps // 2:
            probe.fire()
        step(dev[(4 + warmup + i) % nframes])
    e1.record(stream)
    barrier()
    clocks = probe.summary() if probe is not None else None   # NVML only now: every rank is past its timed region
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    timeouts = b.timeouts() if perlevel else (link.timeouts() if link is not None else 0)

    # ---- end to end: this rank's band of the pinned host G-buffer -> H2D -> frame -> D2H of its owned RGBA8 rows ----
    copy_s, d2h_s = torch.cuda.Stream(), torch.cuda.Stream()
    slots = [[torch.empty_like(t, device="cuda") for t in host[0]] for _ in range(2)]
    o8_dev = [torch.empty_like(out8) for _ in range(2)]
    o8_host = [torch.empty((band.rows, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_down = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i):
        s = i & 1
        with torch.cuda.stream(copy_s):
            if i >= 2:
                copy_s.wait_event(ev_done[s])       # the slot's inputs were read by frame i-2
            for d, h in zip(slots[s], host[i % n_pinned]):
                d.copy_(h, non_blocking=True)
            ev_up[s].record(copy_s)
        stream.wait_event(ev_up[s])
        if i >= 2:
            stream.wait_event(ev_down[s])           # the slot's output was drained by download i-2
        step(slots[s], o8_dev[s])
        ev_done[s].record(stream)
        with torch.cuda.stream(d2h_s):
            d2h_s.wait_event(ev_done[s])
            o8_host[s].copy_(b.owned(o8_dev[s]), non_blocking=True)
            ev_down[s].record(d2h_s)

    n_e2e = min(steps, 20)
    for i in range(4):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(4 + i)
    torch.cuda.synchronize()
    e2e_s = shard.max_over_ranks(time.perf_counter() - t0, device="cuda")
    barrier()
    timeouts += (b.timeouts() if perlevel else 0)
    h2d = IN_BYTES_PX * b.ext_rows * W
    d2h = 4 * band.rows * W
    tot_h2d = shard.sum_over_ranks(h2d, device="cuda")
    tot_d2h = shard.sum_over_ranks(d2h, device="cuda")
    px = W * H
    res = {"value": px * steps / (ms * 1e-3) / 1e6, "ms_per_step": ms / steps, "clocks": clocks,
           "band_rows": band.rows, "ext_rows": b.ext_rows, "frames_resident": nframes,
           "timeouts": int(shard.sum_over_ranks(timeouts, device="cuda")),
           "launches_per_frame": b.launches_per_frame() if perlevel else b.ctx.last_launch_count(),
           "e2e": {"value": px * n_e2e / e2e_s / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": int(tot_h2d),
                   "d2h_bytes_per_step": int(tot_d2h), "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
                   "api": "BandedSvgfV2.frame per rank on its band (+ halo rows) of the pinned host G-buffer: H2D, band frame, "
                          "D2H of the owned RGBA8 rows; double-buffered copy streams; bytes summed over ranks"}}
    if perlevel:
        b.close()
    return res


def replicas_run(rank, world, local_rank, steps, warmup):
    """configs[4]: one independent 1080p sequence per GPU, no communication (weak scaling)."""
    import torch
    import torch.distributed as dist
    import raymarchdenoisercuda_b200 as rmd
    from raymarchdenoisercuda_b200 import shard
    from raymarchdenoisercuda_b200.synth import synth_frame
    W, H = 1920, 1080
    seed = 0x5EED0100 + rank
    nframes = min(steps + warmup, 24)
    dev = [[_t(x).cuda() for x in synth_frame(W, H, seed, f)] for f in range(nframes)]
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    params = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=DEPTH, radius=2)
    ctx = rmd.SvgfContext(W, H, local_rank)
    stream = torch.cuda.current_stream()
    for i in range(warmup):
        ctx.frame(*dev[i % nframes], out, params)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe = DeviceClockProbe(local_rank) if rank == 0 else None
    e0.record(stream)
    for i in range(steps):
        if probe is not None and i == steps // 2:
            probe.fire()
        ctx.frame(*dev[(warmup + i) % nframes], out, params)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    if world > 1:
        dist.barrier()
    clocks = probe.summary() if probe is not None else None
    ctx.close()
    return {"clocks": clocks, "workload": f"configs[4]: {world} independent 1920x1080 sequences, one context + stream per GPU, no collective",
            "scaling": "weak", "value": world * W * H * steps / (ms * 1e-3) / 1e6, "unit": "Mpixel/s",
            "ms_per_step": ms / steps, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="8k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (1080p, 4K sequence, box path, replicas)")
    ap.add_argument("--cpu-budget", type=float, default=None,
                    help="seconds of CPU-oracle work (default: 25 for cpu_baseline, 90 for --impl reference)")
    ap.add_argument("--mode", default="auto", choices=["auto", "banded", "sequences"],
                    help="N>1: 'banded' (default) = ONE frame sequence split into row bands over the ranks (strong "
                         "scaling); 'sequences' = one independent sequence per GPU (weak scaling)")
    ap.add_argument("--scheme", default="perlevel", choices=["perlevel", "halo"],
                    help="banded mode: 'perlevel' = a-trous levels produce only the band's rows and push per-level halo rows "
                         "to the neighbour over NVLink (no recompute); 'halo' = 80 recomputed halo rows + one history swap per frame")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="banded --scheme halo: history-row exchange over NVLink peer mappings or NCCL send/recv")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.cpu_budget is None:
            args.cpu_budget = 90.0   # seconds of oracle work for the whole --steps/--warmup run
        run_reference(args, rank, world)
        return
    if args.cpu_budget is None:
        args.cpu_budget = 25.0
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1; the synthetic generator (host, OpenMP) would then build the 8K sequence on
        # one core per rank.  Give every rank its share of the host before libgomp initialises.
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // int(os.environ.get("LOCAL_WORLD_SIZE", world))))

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(args.warmup, 3)
    steps = args.steps
    W, H, seed, _ = WORKLOADS[args.workload]
    px = W * H
    peak, peak_src = peaks()
    mode = "single" if world == 1 else ("sequences" if args.mode == "sequences" else "banded")

    if mode == "single":
        clk = ClockSampler(local_rank)
        res = single_gpu_run(args.workload, steps, warmup, local_rank, 32 if px > 8e6 else 72, sampler=clk)
        line = {
            "metric": "Mpixel/s full SVGF frame", "value": res["value"], "unit": "Mpixel/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload),
            "frames_resident": res["frames_resident"],
            "roofline": level_roofline(res, args.workload), "cpu_baseline": None, "e2e": res["e2e"],
            "e2e_fp32": res["e2e_fp32"], "gpu_launches": res["launches_per_frame"] * steps,
            "launches_per_frame": res["launches_per_frame"], "clocks": clk.summary(), "checksum": res["checksum_device"],
        }
        if not args.no_sub and args.workload == "8k":
            sub = {}
            r = single_gpu_run("1080p", 40, 8, local_rank, 48)
            sub["configs1_1080p"] = {"config": workload_config("1080p"), "value": r["value"], "unit": "Mpixel/s",
                                     "ms_per_step": r["ms_per_step"], "steps": 40, "roofline": level_roofline(r, "1080p"),
                                     "e2e": r["e2e"], "e2e_fp32": r["e2e_fp32"]}
            r = single_gpu_run("4k", 64, 0, local_rank, 64, per_frame=True, with_e2e=False, from_reset=True)
            fm = r["frame_ms"]
            steady = float(np.mean(fm[8:]))
            steady_med = float(np.median(fm[8:]))
            # a steady-state frame far above the median is a stall between frames (seen once: one frame of 56 at
            # 2.9 ms on an otherwise 1.135 ms sequence), not a property of the frame: listed, and still in the mean
            outliers = [[i, float(t)] for i, t in enumerate(fm) if i >= 8 and t > 1.5 * steady_med]
            sub["configs2_4k_sequence"] = {
                "config": workload_config("4k"), "frames": 64, "from": "empty history (frame 0 has no history: every "
                "pixel takes the 7x7 variance pass until its history is 4 frames long)",
                "steady_state_ms": steady, "steady_state_frames": "8..63", "steady_state_mpixel_s": 3840 * 2160 / steady / 1e3,
                "steady_state_median_ms": steady_med, "steady_state_outlier_frames": outliers,
                "worst_frame_ms": float(max(fm)), "worst_frame_index": int(np.argmax(fm)), "first_frames_ms": fm[:8],
                "worst_history_less_frame_ms": float(max(fm[:4])),
                "target_ms": 1.0, "roofline": level_roofline(dict(r, ms_per_step=steady), "4k")}
            sub["reference_gpu_box"] = box_vs_reference_gpu(local_rank)
            line["sub_records"] = sub
        if not args.no_cpu_baseline:
            os.environ.setdefault("OMP_NUM_THREADS", str(host_threads()))
            v, ms_cpu, threads, sample, whole = cpu_arm(W, H, seed, 2, 1, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample,
                                    "ms_per_step": ms_cpu, "whole_frames": whole}
        print(json.dumps(line))
        return

    if mode == "banded":
        res = banded_run(args, args.workload, rank, world, local_rank, steps, warmup)
        sub = None
        if not args.no_sub:
            sub = {"replicas": replicas_run(rank, world, local_rank, 40, 8)}
        if rank == 0:
            frame_gbs = BYTES_FRAME * px / (res["ms_per_step"] * 1e-3) / 1e9 / world
            line = {
                "metric": "Mpixel/s full SVGF frame", "value": res["value"], "unit": "Mpixel/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.workload, world, "banded"),
                "band": {"rows": res["band_rows"], "ext_rows": res["ext_rows"], "scheme": args.scheme,
                         "exchange": "NVLink peer stores + flag words (CUDA IPC), neighbour point-to-point" if args.scheme == "perlevel" else args.exchange,
                         "flag_wait_timeouts": res["timeouts"]},
                "frames_resident": res["frames_resident"],
                "roofline": {"bound": "hbm", "kernel": "whole frame, per GPU (all passes incl. halo exchange)",
                             "achieved": frame_gbs, "peak": peak, "unit": "GB/s", "frac": frame_gbs / peak, "traffic": None,
                             "peak_source": peak_src, "algorithmic_bytes_per_px": BYTES_FRAME},
                "cpu_baseline": None, "e2e": res["e2e"], "gpu_launches": res["launches_per_frame"] * steps,
                "launches_per_frame": res["launches_per_frame"], "clocks": res["clocks"],
            }
            if sub:
                line["sub_records"] = sub
            print(json.dumps(line))
    else:
        r = replicas_run(rank, world, local_rank, steps, warmup)
        if rank == 0:
            frame_gbs = BYTES_FRAME * 1920 * 1080 / (r["ms_per_step"] * 1e-3) / 1e9
            print(json.dumps({
                "metric": "Mpixel/s full SVGF frame", "value": r["value"], "unit": "Mpixel/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config("1080p", world, "sequences"),
                "roofline": {"bound": "hbm", "kernel": "whole frame, per GPU", "achieved": frame_gbs, "peak": peak, "unit": "GB/s",
                             "frac": frame_gbs / peak, "traffic": None, "peak_source": peak_src},
                "cpu_baseline": None, "e2e": None, "gpu_launches": 7 * steps, "clocks": r["clocks"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
