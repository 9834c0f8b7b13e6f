"""Python front-end of the host-only synthetic G-buffer generator (synth/synth_scene.c)."""
import numpy as np

from . import _lib


def synth_frame(width, height, seed, frame, layers=24):
    """Returns (color f16 (H,W,4), albedo u8 (H,W,4), guide u32 (H,W,2), motion f16 (H,W,2)) numpy arrays."""
    color = np.empty((height, width, 4), np.float16)
    albedo = np.empty((height, width, 4), np.uint8)
    guide = np.empty((height, width, 2), np.uint32)
    motion = np.empty((height, width, 2), np.float16)
    _lib.load_synth().rmd_synth_frame(width, height, seed, layers, frame, color.ctypes.data, albedo.ctypes.data,
                                      guide.ctypes.data, motion.ctypes.data)
    return color, albedo, guide, motion
