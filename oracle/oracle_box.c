/*
 * oracle_box.c — CPU restatement of the reference's box-filter kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * Parity status: PINNED.  oracle/_ref/libref_cpu.so is the reference's own
 * src/filter.cu compiled for the host under a SIMT shim (oracle/ref_build/);
 * tests/test_oracle_box.py checks this restatement bit-for-bit against it on
 * render/cornell/1/render.png-derived fixtures (tests/golden/) and on random
 * frames, and against SURVEY.md Appendix C's hashes.
 *
 * Follows, line by line:
 *   filterKernelBaseline   reference src/filter.cu:13-58
 *   filterKernelTiled      reference src/filter.cu:87-158 (cacheInput=false branch, :131-132)
 *   inRange / flattenIndex reference include/extended_math.h:62-68
 * Levels are iterated as separate passes (the in-kernel depth loop of the
 * reference synchronises only one block, src/filter.cu:56, so multi-level
 * results of a single launch are a race).
 */
#include <stdint.h>
#include <stddef.h>

static inline int in_range(int x, int y, int W, int H) { /* extended_math.h:62-64 */
    return (x >= 0) && (x < W) && (y >= 0) && (y < H);
}
static inline int flatten(int x, int y, int W) { return y * W + x; } /* extended_math.h:66-68 */

/* One level of filterKernelBaseline (src/filter.cu:30-54).  in/out: uchar4[W*H]. */
void oracle_box_baseline_level(const uint8_t* in, uint8_t* out, int W, int H, int radius) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float acc_x = 0.f, acc_y = 0.f, acc_z = 0.f, norm = 0.f; /* :30-31 */
            for (int dx = -radius; dx <= radius; ++dx)              /* :34  x outer */
                for (int dy = -radius; dy <= radius; ++dy) {        /* :35  y inner */
                    int nx = x + dx, ny = y + dy;                   /* :36 */
                    if (!in_range(nx, ny, W, H)) continue;          /* :38-39 */
                    const float w = 1.f;                            /* :41 */
                    const uint8_t* m = in + 4 * (size_t)flatten(nx, ny, W);
                    acc_x += w * m[0];
                    acc_y += w * m[1];
                    acc_z += w * m[2];
                    norm += w;
                }
            acc_x /= norm; acc_y /= norm; acc_z /= norm;            /* :48 */
            (void)acc_y; (void)acc_z;
            uint8_t* o = out + 4 * (size_t)flatten(x, y, W);
            o[0] = (uint8_t)acc_x;                                  /* :51 */
            o[1] = (uint8_t)acc_x;                                  /* :52 (sic: .x) */
            o[2] = (uint8_t)acc_x;                                  /* :53 (sic: .x) */
            o[3] = 0; /* :50 leaves .w uninitialised; the oracle and the product define it as 0 */
        }
}

/* One level of filterKernelTiled with cacheInput=false (src/filter.cu:115-155). */
void oracle_box_tiled_level(const uint8_t* in, uint8_t* out, int W, int H, int radius) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float acc_x = 0.f, acc_y = 0.f, acc_z = 0.f, norm = 0.f;
            for (int dx = -radius; dx <= radius; ++dx)
                for (int dy = -radius; dy <= radius; ++dy) {
                    int nx = x + dx, ny = y + dy;
                    if (!in_range(nx, ny, W, H)) continue;          /* :124-125 */
                    const float w = 1.f;                            /* :127 */
                    const uint8_t* m = in + 4 * (size_t)flatten(nx, ny, W); /* :132 */
                    acc_x += w * m[0];
                    acc_y += w * m[1];
                    acc_z += w * m[2];
                    norm += w;
                }
            acc_x /= norm; acc_y /= norm; acc_z /= norm;            /* :149 */
            uint8_t* o = out + 4 * (size_t)flatten(x, y, W);
            o[0] = (uint8_t)acc_x; o[1] = (uint8_t)acc_y; o[2] = (uint8_t)acc_z; /* :151-154 */
            o[3] = 0;                                               /* aggregate init */
        }
}

/* depth levels with the reference's ping-pong (src/filter.cu:24-25):
 *   in  = level == 0 ? render : buffer[level % 2]
 *   out = level == depth-1 ? denoised : buffer[(level + 1) % 2]
 * variant: 0 = baseline, 1 = tiled(cacheInput=false).  buf0/buf1 may be NULL when depth == 1. */
int oracle_box_filter(const uint8_t* render, uint8_t* denoised, uint8_t* buf0, uint8_t* buf1, int W, int H,
                      int radius, int depth, int variant) {
    uint8_t* buffer[2] = {buf0, buf1};
    if (depth > 1 && (!buf0 || !buf1)) return -1;
    for (int level = 0; level < depth; ++level) {
        const uint8_t* in = (level == 0) ? render : buffer[level % 2];
        uint8_t* out = (level == depth - 1) ? denoised : buffer[(level + 1) % 2];
        if (variant == 0) oracle_box_baseline_level(in, out, W, H, radius);
        else oracle_box_tiled_level(in, out, W, H, radius);
    }
    return 0;
}
