"""torchrun: per-stage GPU time and host enqueue time of the per-level banded frame (v2)."""
import os, sys, json, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raymarchdenoisercuda_b200 as rmd
from raymarchdenoisercuda_b200 import shard
from raymarchdenoisercuda_b200.synth import synth_frame
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, H = 7680, 4320
rows_div = int(os.environ.get("EMULATE_RANKS", world))   # band height as if there were this many ranks
band_all = shard.row_bands(H, rows_div)
# take `world` adjacent bands from the middle so that every band has the size of an 8-rank run
start = (rows_div - world) // 2
mine = band_all[start + rank]
band = shard.Band(rank, mine.row0, mine.rows, 0, 0)
b = shard.BandedSvgfV2(W, H, band, lr)
# neighbours only inside the group
planes = synth_frame(W, H, 3, 0)
dev = [torch.from_numpy(np.ascontiguousarray(b.slice_rows(x)).view(np.int32) if x.dtype == np.uint32 else np.ascontiguousarray(b.slice_rows(x))).cuda() for x in planes]
# neighbours only inside the group: hide the outer sides while linking, the context keeps its halo rows
top, bot = b.top, b.bot
if rank == 0: b.top = 0
if rank == world - 1: b.bot = 0
b.connect_ipc()
b.top, b.bot = top, bot
out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(6); host = 0.0; n = 0
for i in range(16):
    e = [ev() for _ in range(7)]
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(6):
        e[s].record(); b.stage(s, *dev, out, p)
    e[6].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if i >= 4:
        acc += [e[k].elapsed_time(e[k + 1]) for k in range(6)]; host += (t1 - t0) * 1e3; n += 1
t = torch.tensor(list(acc / n) + [host / n], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"ranks": world, "band_rows": band.rows, "ext_rows": b.ext_rows, "gpu_ms_per_stage_max": [round(float(v), 3) for v in t[:6]], "gpu_ms_frame": round(float(t[:6].sum()), 3), "host_enqueue_ms": round(float(t[6]), 3), "timeouts": b.lib.rmd_p2p_timeouts()}))
dist.destroy_process_group()
