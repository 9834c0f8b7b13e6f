"""Shared helpers of the test-suite: hand-made G-buffers in the storage formats of
include/rmd_b200.h, error metrics with the tolerances BASELINE.json states."""
import hashlib

import numpy as np

MAX_ABS_TOL = 1e-3   # BASELINE.json north_star: "max-abs error 1e-3 on linear radiance"
PSNR_MIN_DB = 60.0   # BASELINE.json north_star: "PSNR >= 60 dB"


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def oct_encode(n):
    """unit normals (..., 3) -> uint32 word (snorm16 x | snorm16 y << 16), octahedral."""
    n = np.asarray(n, np.float64)
    n = n / np.abs(n).sum(-1, keepdims=True)
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    fx = np.where(z < 0, (1 - np.abs(y)) * np.where(x >= 0, 1, -1), x)
    fy = np.where(z < 0, (1 - np.abs(x)) * np.where(y >= 0, 1, -1), y)
    sx = np.round(fx * 32767).astype(np.int32) & 0xFFFF
    sy = np.round(fy * 32767).astype(np.int32) & 0xFFFF
    return (sx | (sy << 16)).astype(np.uint32)


def make_guide(normal, z):
    """normal (H,W,3) float, z (H,W) float32 -> guide (H,W,2) uint32."""
    g = np.empty(z.shape + (2,), np.uint32)
    g[..., 0] = oct_encode(normal)
    g[..., 1] = np.asarray(z, np.float32).view(np.uint32)
    return g


def flat_gbuffer(H, W, radiance, albedo_u8=255, z=4.0, normal=(0.0, 0.0, 1.0)):
    """Constant-geometry G-buffer with the given radiance (H,W,3) or scalar."""
    rad = np.broadcast_to(np.asarray(radiance, np.float32), (H, W, 3)) if np.ndim(radiance) != 3 else radiance
    color = np.ones((H, W, 4), np.float16)
    color[..., :3] = rad
    albedo = np.full((H, W, 4), 255, np.uint8)
    albedo[..., :3] = albedo_u8
    nrm = np.broadcast_to(np.asarray(normal, np.float64), (H, W, 3))
    zz = np.broadcast_to(np.asarray(z, np.float32), (H, W)).copy()
    guide = make_guide(nrm, zz)
    motion = np.zeros((H, W, 2), np.float16)
    return color, albedo, guide, motion


def lum(c):
    return 0.2126 * c[..., 0] + 0.7152 * c[..., 1] + 0.0722 * c[..., 2]


def cornell_svgf_inputs(npz):
    """BASELINE configs[0] / SURVEY §8d config 1b: the reference's 8-bit cornell planes as SVGF inputs.
    radiance = render/255 (linear, as stored), albedo as is, normal n = rgb/255 renormalised with zero-length
    normals mapped to (0,0,1) (the fixture lost all negative components), depth = 1.0 everywhere (depth.png is
    saturated), no motion."""
    render, albedo, normal = npz["render"], npz["albedo"], npz["normal"]
    H, W, _ = render.shape
    color = np.ones((H, W, 4), np.float16)
    color[..., :3] = render.astype(np.float32) / 255.0
    alb = np.full((H, W, 4), 255, np.uint8)
    alb[..., :3] = albedo
    n = normal.astype(np.float64) / 255.0
    ln = np.linalg.norm(n, axis=-1, keepdims=True)
    n = np.where(ln > 1e-6, n / np.maximum(ln, 1e-6), np.array([0.0, 0.0, 1.0]))
    guide = make_guide(n, np.ones((H, W), np.float32))
    motion = np.zeros((H, W, 2), np.float16)
    return color, alb, guide, motion


def stress_gbuffer(W, H, seed, frame):
    """Adversarial hand-made G-buffer (numpy): piecewise-constant patches with random full-sphere normals
    (exercises the octahedral fold), depth steps and slopes, sky holes, zero albedo channels, HDR fireflies and
    motion up to +-20 px on the 1/16-px grid (so that floor() and thresholds stay exact in fp32)."""
    rng = np.random.default_rng(seed)               # scene layout: same for every frame
    P = 16                                          # patch size
    ph, pw = (H + P - 1) // P, (W + P - 1) // P
    nrm = rng.normal(size=(ph, pw, 3)); nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    zb = rng.choice([1.0, 2.5, 2.75, 8.0, 64.0], size=(ph, pw)).astype(np.float32)
    slope = rng.choice([0.0, 1 / 1024, 1 / 64], size=(ph, pw)).astype(np.float32)
    alb = rng.integers(0, 256, size=(ph, pw, 3)).astype(np.uint8)
    alb[rng.random((ph, pw)) < 0.15] = 0             # black albedo patches (demodulation floor)
    alb[rng.random((ph, pw, 3)) < 0.1] = 0           # single zero channels
    sky = rng.random((ph, pw)) < 0.08
    mvp = (rng.integers(-320, 321, size=(ph, pw, 2)) / 16.0).astype(np.float32)
    up = lambda a: np.repeat(np.repeat(a, P, axis=0), P, axis=1)[:H, :W]
    xs = np.arange(W, dtype=np.float32)[None, :]
    z = up(zb) + up(slope) * (xs % P)
    z = np.where(up(sky), 0.0, z).astype(np.float32)
    fr = np.random.default_rng(seed * 1000 + frame)  # per-frame noise
    rad = fr.exponential(1.0, size=(H, W, 3)).astype(np.float32) * fr.uniform(0.05, 2.0, size=(H, W, 1)).astype(np.float32)
    fire = fr.random((H, W)) < 0.002
    rad[fire] *= 500.0
    color = np.ones((H, W, 4), np.float16)
    color[..., :3] = np.minimum(rad, 6.0e4)
    albedo = np.full((H, W, 4), 255, np.uint8)
    albedo[..., :3] = up(alb)
    guide = make_guide(up(nrm), z)
    motion = up(mvp).astype(np.float16)
    return color, albedo, guide, motion


def gbuffer_to_svgf_inputs_f32(render, albedo, normal):
    """numpy mirror (IEEE fp32, operation for operation) of csrc/svgf_ctx.cu:gbuffer_convert_kernel, the device-side
    conversion behind rmd_svgf_frame_gbuffer: RGBA8 (H,W,4) planes -> (color f16, albedo u8, guide u32x2, motion f16)."""
    f = np.float32
    k = f(1.0) / f(255.0)
    H, W, _ = render.shape
    color = np.ones((H, W, 4), np.float16)
    color[..., :3] = (render[..., :3].astype(f) * k).astype(np.float16)
    n = normal[..., :3].astype(f) * k
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    ln = np.sqrt(((x * x + y * y) + z * z).astype(f)).astype(f)
    zero = ln < f(1e-6)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f(1.0) / ln).astype(f)
        x, y, z = np.where(zero, f(0), x * inv).astype(f), np.where(zero, f(0), y * inv).astype(f), np.where(zero, f(1), z * inv).astype(f)
        s = ((np.abs(x) + np.abs(y)) + np.abs(z)).astype(f)
        px, py = (x / s).astype(f), (y / s).astype(f)
    ox = ((f(1) - np.abs(py)) * np.where(px >= 0, f(1), f(-1))).astype(f)
    oy = ((f(1) - np.abs(px)) * np.where(py >= 0, f(1), f(-1))).astype(f)
    px, py = np.where(z < 0, ox, px), np.where(z < 0, oy, py)
    sx = np.rint(px * f(32767.0)).astype(np.int32) & 0xFFFF
    sy = np.rint(py * f(32767.0)).astype(np.int32) & 0xFFFF
    guide = np.empty((H, W, 2), np.uint32)
    guide[..., 0] = (sx | (sy << 16)).astype(np.uint32)
    guide[..., 1] = np.float32(1.0).view(np.uint32)
    return color, np.ascontiguousarray(albedo), guide, np.zeros((H, W, 2), np.float16)
