"""Decodes the reference's cornell fixture (render/cornell/1/{render,albedo,normal}.png, 500x500 RGB8) into
tests/golden/cornell_gbuffer.npz so that BASELINE configs[0] can run where /root/reference is absent.
Run in the build container:  python tests/golden/make_cornell_gbuffer.py
depth.png is saturated (every byte 255) and there is no motion fixture, so neither is stored."""
import os
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/render/cornell/1"
planes = {n: np.array(Image.open(os.path.join(SRC, n + ".png")).convert("RGB")) for n in ("render", "albedo", "normal")}
depth = np.array(Image.open(os.path.join(SRC, "depth.png")).convert("RGB"))
assert depth.min() == 255, "depth fixture is expected to be saturated"
np.savez_compressed(os.path.join(HERE, "cornell_gbuffer.npz"), **planes)
print({k: (v.shape, int(v.mean())) for k, v in planes.items()})
