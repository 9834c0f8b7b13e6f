"""torchrun -N ranks: row-banded frames with the NVLink P2P exchange vs the full frame computed on every rank's own
GPU; owned rows must be bit-identical for every frame."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raymarchdenoisercuda_b200 as rmd
from raymarchdenoisercuda_b200 import shard
from raymarchdenoisercuda_b200.synth import synth_frame
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, H = 640, 200 * world + 56
V2 = "--v2" in sys.argv
if V2:
    b = shard.BandedSvgfV2(W, H, shard.row_bands(H, world)[rank], lr)
    b.connect_ipc()
    link = None
else:
    halo = shard.banded_halo(5)
    b = shard.BandedSvgf(W, H, shard.row_bands(H, world)[rank], halo, lr)
    link = shard.P2PLink(b)
full = rmd.SvgfContext(W, H, lr)
out_b = torch.empty((b.ext_rows, W, 4), dtype=torch.float32, device="cuda")
out_f = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
bad = 0
for f in range(8):
    planes = synth_frame(W, H, 0x5EED0051, f)
    tens = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in planes]
    full.frame(*tens, out_f, p)
    if V2:
        b.frame(*[b.slice_rows(t).contiguous() for t in tens], out_b, p)
    else:
        b.ctx.frame(*[b.slice_rows(t).contiguous() for t in tens], out_b, p)
        link.exchange()
    torch.cuda.synchronize()
    same = torch.equal(b.owned(out_b), out_f[b.band.row0:b.band.row0 + b.band.rows])
    bad += 0 if same else 1
    print(f"rank {rank} frame {f}: owned rows bit-identical = {same}", flush=True)
t = torch.tensor([bad, b.timeouts() if V2 else link.timeouts()], device="cuda"); dist.all_reduce(t)
if rank == 0:
    print("P2P BANDED CHECK", "v2 (per-level)" if V2 else "v1 (halo recompute)", "OK" if int(t[0]) == 0 and int(t[1]) == 0 else f"FAILED mismatches={int(t[0])} timeouts={int(t[1])}")
dist.destroy_process_group()
