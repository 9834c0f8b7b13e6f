/*
 * ref_gpu_driver.cu — launches the reference's UNMODIFIED kernels (compiled from
 * /root/reference/src/filter.cu for sm_100a; the shipped Makefile targets sm_75
 * with -G, reference Makefile:8) with the reference's own launch geometry
 * (src/test.cu:70-75, 82-87), but checks errors and lets the caller read results
 * back, which the reference's tests never do.  Used on the GPU box (a) to confirm
 * that the CPU build of the same source (libref_cpu.so) and the restatement
 * (oracle_box.c) reproduce the real kernels bit for bit, and (b) to time the
 * reference's kernels beside the new path.  TEST INFRASTRUCTURE ONLY.
 */
#include "filter.cuh"
#include <cuda_runtime.h>

/* device pointers in, one launch, no sync */
extern "C" int ref_gpu_launch(void* in, void* out, void* buf0, void* buf1, int W, int H, int radius, int depth,
                              int variant, int cacheInput, void* stream) {
    GBuffer frame{};
    frame.shape = {W, H};
    frame.render = (uchar4*)in;
    frame.denoised = (uchar4*)out;
    frame.buffer[0] = (uchar4*)buf0;
    frame.buffer[1] = (uchar4*)buf1;
    FilterParams params{};
    params.type = FilterParams::AVERAGE;
    params.depth = depth;
    params.radius = radius;
    params.cacheInput = cacheInput != 0;
    dim3 blockSize(16, 16);
    dim3 gridSize((W + blockSize.x - 1) / blockSize.x, (H + blockSize.y - 1) / blockSize.y);
    if (variant == 0) filterKernelBaseline<<<gridSize, blockSize, 49152, (cudaStream_t)stream>>>(frame, params);
    else filterKernelTiled<<<gridSize, blockSize, 30 * 1024, (cudaStream_t)stream>>>(frame, params);
    return (int)cudaGetLastError();
}
