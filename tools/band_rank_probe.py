#!/usr/bin/env python
"""torchrun -N ranks: per-rank, per-stage GPU time of the per-level band path on the 8K frame (CUDA events between the
stages, mean over the timed frames), gathered and printed by rank 0.  Shows where a band frame spends its time and which
rank is the slowest — the number bench.py reports is the max over ranks of the whole frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
import raymarchdenoisercuda_b200 as rmd  # noqa: E402
from raymarchdenoisercuda_b200 import shard  # noqa: E402
from raymarchdenoisercuda_b200.synth import synth_frame  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, H, NF, STEPS = 7680, 4320, 8, 24
b = shard.BandedSvgfV2(W, H, shard.row_bands(H, world)[rank], lr)
b.connect_ipc()
dev = [[torch.from_numpy(np.ascontiguousarray(b.slice_rows(x)).view(np.int32) if x.dtype == np.uint32
                         else np.ascontiguousarray(b.slice_rows(x))).cuda() for x in synth_frame(W, H, 0x5EED0003, f)] for f in range(NF)]
out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
stream = torch.cuda.current_stream()
torch.cuda.synchronize(); dist.barrier()
for i in range(6):
    b.frame(*dev[i % NF], out, p)
torch.cuda.synchronize(); dist.barrier()
# whole frames, no marks
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(STEPS):
    b.frame(*dev[i % NF], out, p)
e1.record(stream)
torch.cuda.synchronize(); dist.barrier()
frame_us = e0.elapsed_time(e1) / STEPS * 1e3
# per stage
acc = np.zeros(6)
for i in range(STEPS):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record(stream)
    for s in range(6):
        b.stage(s, *dev[i % NF], out, p)
        ev[s + 1].record(stream)
    torch.cuda.synchronize()
    acc += [ev[s].elapsed_time(ev[s + 1]) * 1e3 for s in range(6)]
t = torch.tensor([frame_us] + (acc / STEPS).tolist() + [b.timeouts()], device="cuda", dtype=torch.float64)
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
if rank == 0:
    print("rank  frame_us | temporal+push  var+L0  L1  L2  L3  L4+unpack | timeouts")
    for r, x in enumerate(allt):
        x = x.tolist()
        print(f"{r:4d}  {x[0]:8.1f} | " + "  ".join(f"{v:7.1f}" for v in x[1:7]) + f" | {int(x[7])}")
b.close()
dist.destroy_process_group()
