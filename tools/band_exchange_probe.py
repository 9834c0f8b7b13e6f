"""Times the pieces of the banded mode's per-frame exchange (torchrun, N ranks): frame, pack, send/recv, unpack."""
import os, sys, json
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
halo = shard.banded_halo(5)
b = shard.BandedSvgf(W, H, shard.row_bands(H, world)[rank], halo, lr)
planes = synth_frame(W, H, 3, 0)
dev = [torch.from_numpy(np.ascontiguousarray(b.slice_rows(x)).view(np.int32) if x.dtype == np.uint32 else np.ascontiguousarray(b.slice_rows(x))).cuda() for x in planes]
out = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
def ops():
    o, r = [], rank
    if b.top: o += [dist.P2POp(dist.isend, b.send_up, r - 1), dist.P2POp(dist.irecv, b.recv_up, r - 1)]
    if b.bot: o += [dist.P2POp(dist.isend, b.send_dn, r + 1), dist.P2POp(dist.irecv, b.recv_dn, r + 1)]
    return o
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(4); n = 0
for i in range(14):
    e = [ev() for _ in range(5)]
    dist.barrier(); torch.cuda.synchronize()
    e[0].record(); b.ctx.frame(*dev, out, p)
    e[1].record(); b.pack()
    e[2].record()
    for req in dist.batch_isend_irecv(ops()): req.wait()
    e[3].record(); b.unpack()
    e[4].record(); torch.cuda.synchronize()
    if i >= 4:
        acc += [e[k].elapsed_time(e[k + 1]) for k in range(4)]; n += 1
t = torch.tensor(acc / n, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"ranks": world, "ext_rows": b.ext_rows, "ms_max_over_ranks": dict(zip(["frame", "pack", "sendrecv", "unpack"], [round(float(v), 3) for v in t])), "bytes_per_boundary": b.ctx.history_bytes(halo)}))
dist.destroy_process_group()
