"""Generates the golden vectors of the legacy box path FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref/libref_cpu.so, the
reference's unmodified src/filter.cu compiled for the host by oracle/Makefile):

    python tests/golden/make_box_golden.py

Writes
  tests/golden/cornell_render_rgba.npz   the reference's only fixture, render/cornell/1/render.png,
                                         decoded as the reference decodes it (Image(path, 4):
                                         stbi_load with req_comp=4 => RGBA, A=255; src/image.cpp:33-40)
  tests/golden/box_golden.json           sha256[:16] + byte sums of the reference kernels' outputs on
                                         that image, levels 1..5 host-iterated, both kernels
  tests/golden/box_crop_golden.npz       a 96x64 crop with the full reference outputs (depth 1 and 5)
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    img = np.array(Image.open("/root/reference/render/cornell/1/render.png").convert("RGB"))
    H, W, _ = img.shape
    rgba = np.concatenate([img, np.full((H, W, 1), 255, np.uint8)], axis=2).copy()
    np.savez_compressed(os.path.join(HERE, "cornell_render_rgba.npz"), render=rgba)
    gold = {"input_sha": sha(rgba), "shape": [W, H], "radius": 2, "tiled": [], "baseline": []}
    for variant in ("tiled", "baseline"):
        x = rgba
        for lvl in range(5):
            x = pyoracle.ref_cpu_level(x, 2, variant)
            if variant == "baseline":
                x[..., 3] = 0
            hashed = x if variant == "tiled" else x[..., :3]
            gold[variant].append({"level": lvl + 1, "sha": sha(hashed), "sum": int(hashed.astype(np.int64).sum()),
                                  "px_250_250": [int(v) for v in x[250, 250]]})
    with open(os.path.join(HERE, "box_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    crop = rgba[200:264, 150:246].copy()  # 96 wide x 64 high, noisy interior
    out = {"render": crop}
    for variant in ("tiled", "baseline"):
        for depth in (1, 5):
            for radius in (1, 2, 3):
                out[f"{variant}_d{depth}_r{radius}"] = pyoracle.ref_cpu_filter(crop, radius, depth, variant)
    np.savez_compressed(os.path.join(HERE, "box_crop_golden.npz"), **out)
    print(json.dumps(gold)[:400])


if __name__ == "__main__":
    main()
