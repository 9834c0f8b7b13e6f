"""Golden SVGF vectors: the oracle (oracle/oracle_svgf.c, double accumulations) on
  * a 4-frame 96x64 synthetic sequence (seed 0x5EED00A1: camera motion, disocclusions, sky) — every output frame,
  * the reference's cornell fixture (tests/golden/cornell_gbuffer.npz) as one frame — a 4x-decimated output,
stored as fp32 in tests/golden/svgf_golden.npz.  They pin the oracle against silent drift (CPU test) and give the
CUDA path a committed target that does not execute the oracle (GPU test).  The SVGF path has no reference
implementation to generate vectors from (the reference ships only the box filter), so these vectors are
the specification's own — "parity unpinned" stays in force for this path.
Run in the build container:  python tests/golden/make_svgf_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyoracle as po  # noqa: E402
from raymarchdenoisercuda_b200.synth import synth_frame  # noqa: E402
from util import cornell_svgf_inputs  # noqa: E402

SEQ_W, SEQ_H, SEQ_SEED, SEQ_FRAMES = 96, 64, 0x5EED00A1, 4

if __name__ == "__main__":
    out = {}
    orc = po.SvgfOracle(SEQ_W, SEQ_H)
    for f in range(SEQ_FRAMES):
        out[f"seq_{f}"] = orc.frame(*synth_frame(SEQ_W, SEQ_H, SEQ_SEED, f), depth=5).astype(np.float32)
        out[f"seq_histlen_{f}"] = orc.plane(po.PLANE_HISTLEN).copy()
    c, a, g, m = cornell_svgf_inputs(np.load(os.path.join(HERE, "cornell_gbuffer.npz")))
    H, W, _ = c.shape
    out["cornell_dec4"] = po.SvgfOracle(W, H).frame(c, a, g, m, depth=5)[::4, ::4].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "svgf_golden.npz"), **out)
    print({k: (v.shape, float(np.asarray(v, np.float64).mean())) for k, v in out.items()})
