/*
 * ref_cpu_driver.cpp — runs the reference's filterKernelBaseline / filterKernelTiled
 * bodies (compiled from /root/reference/src/filter.cu under simt_shim.h) over an
 * emulated launch grid: block 16x16, grid ceil(W/16) x ceil(H/16), exactly the
 * geometry of the reference's call sites (src/test.cu:70-73, 82-85).
 * One level per call (depth is forced to 1, cacheInput to false: the only
 * configuration of the reference that is a deterministic function of its input,
 * SURVEY.md §8c).  TEST INFRASTRUCTURE ONLY.
 */
#include "filter.cuh"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
uchar4 tile[64 * 1024]; /* the kernel's `extern __shared__ uchar4 tile[]` (src/filter.cu:100); unused with cacheInput=false */
void printGPUProperties() {}

extern "C" int ref_cpu_filter_level(const unsigned char* in, unsigned char* out, int W, int H, int radius, int variant) {
    GBuffer frame{};
    frame.shape = {W, H};
    frame.render = (uchar4*)in;
    frame.denoised = (uchar4*)out;
    FilterParams params{};
    params.type = FilterParams::AVERAGE;
    params.depth = 1;
    params.radius = radius;
    params.cacheInput = false;
    params.cacheBuffer = false;
    const int bx = 16, by = 16;
    const int gx = (W + bx - 1) / bx, gy = (H + by - 1) / by;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < gx * gy; ++b) {
        blockDim = dim3(bx, by, 1);
        gridDim = dim3(gx, gy, 1);
        blockIdx = {(unsigned)(b % gx), (unsigned)(b / gx), 0};
        for (int ty = 0; ty < by; ++ty)
            for (int tx = 0; tx < bx; ++tx) {
                threadIdx = {(unsigned)tx, (unsigned)ty, 0};
                if (variant == 0) filterKernelBaseline(frame, params);
                else filterKernelTiled(frame, params);
            }
    }
    return 0;
}
