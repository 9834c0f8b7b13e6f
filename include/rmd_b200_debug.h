/*
 * rmd_b200_debug.h — test and inspection hooks of librmd_b200.so.  Not part of the drop-in boundary: nothing the
 * reference's denoise path would bind lives here (the reference has no such introspection, src/test.cu:68-90 only
 * launches and synchronises).  Used by tests/, bench.py (launch counts) and tools/.
 */
#ifndef RMD_B200_DEBUG_H
#define RMD_B200_DEBUG_H

#include "rmd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of kernels the last rmd_svgf_frame enqueued (bench.py's gpu_launches). */
int rmd_svgf_last_launch_count(const rmd_svgf_ctx* ctx);

/* Kernels enqueued by the stages of the last band frame (rmd_svgf_band_stage 0..depth). */
int rmd_svgf_band_launch_count(const rmd_svgf_ctx* ctx);

/* Test/inspection hook: copies an internal plane of the LAST frame to host memory
 * (synchronises `stream`).  Planes are tightly packed W*H.  */
enum {
    RMD_PLANE_TEMPORAL_COLOR = 0, /* float4: accumulated demodulated rgb, .w = luminance  (after temporal [+ variance]) */
    RMD_PLANE_TEMPORAL_VAR = 1,   /* float : variance                                      (after temporal [+ variance]) */
    RMD_PLANE_MOMENTS = 2,        /* float2: accumulated luminance moments                 */
    RMD_PLANE_HISTLEN = 3,        /* uint8 : history length N'                             */
    RMD_PLANE_HISTORY_COLOR = 4,  /* float4: level-0 output (next frame's colour history)  */
    RMD_PLANE_GUIDE = 5,          /* float4: decoded normal xyz, z                         */
    RMD_PLANE_SLOPE = 6           /* float : depth slope dz                                */
};
int rmd_svgf_read_plane(rmd_svgf_ctx* ctx, int plane, void* host_dst, size_t host_bytes, void* stream);
/* Debug knob for per-pass parity: 0 = full frame, 1 = stop after the temporal pass,
 * 2 = stop after the variance pass (planes above then hold that stage's output). */
int rmd_svgf_set_stop_after(rmd_svgf_ctx* ctx, int stage);

/* SM clock from the device: one warp spins `spin_us` microseconds on `stream` and writes {elapsed %clock64 cycles,
 * elapsed %globaltimer nanoseconds} to dev_out2 (device memory, 16 bytes).  MHz = 1000 * cycles / ns.  Used by bench.py
 * under torchrun instead of NVML polling, which stalls stream-ordered cross-GPU hand-offs. */
int rmd_debug_clock_probe(unsigned long long* dev_out2, unsigned int spin_us, void* stream);

/* Host-side enumeration of the tiles ONE a-trous level launch would run — no device work, no GPU needed.  The level
 * kernel maps blockIdx to (column block, row phase, lattice tile) with the same inline function the CPU walks here,
 * so tests can check on the CPU that a launch (and a band's boundary + interior pair, split = 1 / 2) stores every
 * row of [row0, row0 + rows) exactly once per column block and nothing else.
 *   level 0..4 (step 2^level); split / edge ranges as rmd_svgf_band_stage uses them (split 0: every tile of the rows;
 *   1: only tiles holding a row of [edge0_a, +edge_n_a) or [edge0_b, +edge_n_b); 2: all other tiles);
 *   variant < 0 = the shipped kernel variant of that level; reverse = walk last tile first (no effect on coverage).
 *   cover: int[height * column_blocks], incremented; column_blocks / tiles (CTAs that do not return early) are outputs.
 * Returns the grid size (>= 0) or RMD_E_*. */
int rmd_debug_level_cover(int width, int height, int level, int row0, int rows, int split, int edge0_a, int edge_n_a,
                          int edge0_b, int edge_n_b, int reverse, int variant, int* cover, int* column_blocks, int* tiles);

#ifdef __cplusplus
}
#endif
#endif /* RMD_B200_DEBUG_H */
