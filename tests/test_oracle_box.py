"""Pins oracle/oracle_box.c: against the REFERENCE's own src/filter.cu compiled for the host
(oracle/_ref/libref_cpu.so, built by oracle/Makefile when /root/reference is present), against the
golden vectors that build produced (tests/golden/, generator committed beside them) and against
SURVEY.md Appendix C."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle
from util import sha16

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# SURVEY.md Appendix C (derived by the survey from reading the code; confirmed here)
APPENDIX_C = {
    "input": "e6bc2e8029fe4d0e",
    "tiled": ["5e68cbc4fd223f24", "5efd36a1dd1ce5bc", "3af0017ff1152bcc", "9b51f17bdc8eabe4", "5d2929d456d49d13"],
    "baseline": ["b42c68daf74304b4", "5d15545a28bb55d1", "cd2cb4b0142b1294", "f8f4ffe9910def79", "aff262d5d385be3b"],
    "tiled_sum": (59170040, 58077667), "baseline_sum": (66749289, 65565741),
}


def _cornell():
    return np.load(os.path.join(GOLD, "cornell_render_rgba.npz"))["render"]


def test_golden_file_matches_survey_appendix_c():
    gold = json.load(open(os.path.join(GOLD, "box_golden.json")))
    assert gold["input_sha"] == APPENDIX_C["input"]
    for v in ("tiled", "baseline"):
        assert [e["sha"] for e in gold[v]] == APPENDIX_C[v]
        assert (gold[v][0]["sum"], gold[v][4]["sum"]) == APPENDIX_C[v + "_sum"]
    assert gold["tiled"][0]["px_250_250"] == [127, 135, 126, 0]
    assert gold["baseline"][0]["px_250_250"][:3] == [127, 127, 127]


@pytest.mark.parametrize("variant", ["tiled", "baseline"])
def test_oracle_reproduces_reference_on_cornell(variant):
    gold = json.load(open(os.path.join(GOLD, "box_golden.json")))
    img = _cornell()
    assert sha16(img) == gold["input_sha"]
    for depth in range(1, 6):
        out = pyoracle.box_filter(img, 2, depth, variant)
        hashed = out if variant == "tiled" else out[..., :3]
        assert sha16(hashed) == gold[variant][depth - 1]["sha"], (variant, depth)
        assert int(hashed.astype(np.int64).sum()) == gold[variant][depth - 1]["sum"]


@pytest.mark.parametrize("variant", ["tiled", "baseline"])
def test_oracle_reproduces_reference_crop_vectors(variant):
    g = np.load(os.path.join(GOLD, "box_crop_golden.npz"))
    for depth in (1, 5):
        for radius in (1, 2, 3):
            ref = g[f"{variant}_d{depth}_r{radius}"]
            out = pyoracle.box_filter(g["render"], radius, depth, variant)
            sel = slice(None) if variant == "tiled" else slice(0, 3)
            assert np.array_equal(out[..., sel], ref[..., sel]), (variant, depth, radius)


@pytest.mark.skipif(not pyoracle.ref_cpu_available(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (16, 16), (37, 21), (130, 65)])
def test_oracle_vs_reference_source_random(shape):
    """Ragged / tiny frames, radius larger than the frame: border rule skip + renormalise (src/filter.cu:38-39)."""
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    for variant in ("tiled", "baseline"):
        for radius in (1, 2, 5):
            ref = pyoracle.ref_cpu_filter(img, radius, 2, variant)
            out = pyoracle.box_filter(img, radius, 2, variant)
            sel = slice(None) if variant == "tiled" else slice(0, 3)
            assert np.array_equal(out[..., sel], ref[..., sel]), (variant, radius)


def test_float_divide_equals_integer_floor_division():
    """the fp32 divide + truncate of the reference equals integer floor division."""
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (40, 50, 4), dtype=np.uint8)
    out = pyoracle.box_filter(img, 2, 1, "tiled")
    H, W, _ = img.shape
    pad = np.zeros((H + 4, W + 4, 4), np.int64)
    pad[2:-2, 2:-2] = img
    cnt = np.zeros((H + 4, W + 4), np.int64)
    cnt[2:-2, 2:-2] = 1
    s = sum(pad[dy:dy + H, dx:dx + W] for dy in range(5) for dx in range(5))
    c = sum(cnt[dy:dy + H, dx:dx + W] for dy in range(5) for dx in range(5))
    assert np.array_equal(out[..., :3], (s[..., :3] // c[..., None]).astype(np.uint8))
    assert np.all(out[..., 3] == 0)


def test_integer_division_by_multiplication_is_exact_for_every_reachable_sum():
    """csrc/box_filter.cu replaces the reference's `(uchar)(float(sum) / float(count))` (src/filter.cu:48-53) by
    umulhi(sum, ceil(2^32 / count)).  Every count the strip kernel can meet (2 .. 81) and every sum 0 .. 255*count:
    fp32 division + truncation == integer division == the multiplication."""
    for count in range(2, 82):
        n = np.arange(0, 255 * count + 1, dtype=np.uint64)
        magic = np.uint64(((1 << 32) + count - 1) // count)
        by_mul = (n * magic) >> np.uint64(32)
        by_f32 = (n.astype(np.float32) / np.float32(count)).astype(np.uint8)   # IEEE fp32 division, truncation
        assert np.array_equal(by_mul, n // np.uint64(count)), count
        assert np.array_equal(by_f32.astype(np.uint64), by_mul), count
