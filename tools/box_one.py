"""One box-filter launch per size (for ncu): python tools/box_one.py W H [radius] [reps]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raymarchdenoisercuda_b200 as rmd  # noqa: E402

W, H = int(sys.argv[1]), int(sys.argv[2])
radius = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d_in = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda")
d_out = torch.zeros_like(d_in)
frame = rmd.GBuffer((W, H), d_in, d_out)
p = rmd.FilterParams(type=rmd.FilterType.AVERAGE, depth=1, radius=radius)
for _ in range(reps):
    rmd.filter_tiled(frame, p)
torch.cuda.synchronize()
print("ok")
