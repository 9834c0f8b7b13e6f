"""Small end-to-end run for compute-sanitizer: 3 frames 203x117 (ragged tiles), both a-trous kernels, box path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raymarchdenoisercuda_b200 as rmd
from raymarchdenoisercuda_b200.synth import synth_frame
W, H = 203, 117
for tile in ("0", "1"):
    os.environ["RMD_ATROUS_RING"] = tile
    ctx = rmd.SvgfContext(W, H)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    out8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    p = rmd.FilterParams(type=rmd.FilterType.WAVELET, depth=5, radius=2)
    for f in range(3):
        c, a, g, m = synth_frame(W, H, 3, f)
        d = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in (c, a, g, m)]
        ctx.frame(*d, out, p, out_rgba8=out8)
    torch.cuda.synchronize()
    ctx.close()
img = torch.randint(0, 256, (97, 131, 4), dtype=torch.uint8, device="cuda")
o = torch.zeros_like(img); b0 = torch.zeros_like(img); b1 = torch.zeros_like(img)
rmd.filter_tiled(rmd.GBuffer((131, 97), img, o, buffer=(b0, b1)), rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=3, radius=3))
torch.cuda.synchronize()
print("sanitize_small done", float(out.mean()))
