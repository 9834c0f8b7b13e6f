import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raymarchdenoisercuda_b200 as rmd
from oracle import pyoracle as po
from util import cornell_svgf_inputs
npz = np.load(os.path.join(ROOT, "tests", "golden", "cornell_gbuffer.npz"))
c, a, g, m = cornell_svgf_inputs(npz)
H, W, _ = c.shape
dev = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in (c, a, g, m)]
def P(d): return rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=d, radius=2)
for stop, depth in [(1, 5), (2, 5), (0, 0), (0, 1), (0, 2), (0, 5)]:
    ctx, orc = rmd.SvgfContext(W, H), po.SvgfOracle(W, H)
    ctx.set_stop_after(stop)
    out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.frame(*dev, out, P(depth)); torch.cuda.synchronize()
    ref = orc.frame(c, a, g, m, depth=depth)
    if stop == 1:
        gg, rr = ctx.read_plane(0), orc.plane(po.PLANE_TEMPORAL_COLOR_PRE)
    elif stop == 2:
        gg, rr = ctx.read_plane(0), orc.plane(po.PLANE_TEMPORAL_COLOR)
    else:
        gg, rr = out.cpu().numpy(), ref
    err = np.abs(gg[..., :3] - rr[..., :3])
    y, x, ch = np.unravel_index(np.argmax(err), err.shape)
    print(f"stop={stop} depth={depth}: max-abs {err.max():.3e} at (y={y},x={x},ch={ch}) gpu={gg[y,x,:3]} ref={rr[y,x,:3]} albedo={a[y,x,:3]} render={c[y,x,:3]} rel={err.max()/max(abs(rr[y,x,ch]),1e-9):.2e}  #px>1e-3: {(err.max(-1)>1e-3).sum()}")
    if stop == 2:
        vg, vr = ctx.read_plane(1)[..., 0], orc.plane(po.PLANE_TEMPORAL_VAR)[..., 0]
        print("   variance plane: max", vr.max(), "max-abs err", np.abs(vg - vr).max(), "rel", (np.abs(vg-vr)/(np.abs(vr)+1e-6)).max())
    ctx.close()
