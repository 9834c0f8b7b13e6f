// include/compat/filter.cuh — source-level drop-in for the reference's include/filter.cuh: the same `FilterParams`
// (11-23; 36 bytes) and the same two `__global__` entry points (25-26) that the caller launches itself with
// `<<<grid, block, smem>>>` (src/test.cu:73-75, 85-87).  They are defined in librmd_compat.a (relocatable device
// code: compile the caller with -rdc=true / -dc and link the archive in place of the reference's src/filter.cu).
//
// Contract of the two kernels here (csrc/compat/filter_compat.cu):
//   * any 2-D block shape, any grid that covers the image with one thread per pixel, any dynamic shared-memory size
//     (a tile is staged in whatever the launch provides; with none, taps are read from global memory);
//   * params.depth > 1 is evaluated INSIDE the one launch by recomputing the halo per block, so the result equals
//     `depth` host-iterated levels; the reference's in-kernel level loop synchronises one block only and races
//     (src/filter.cu:56);
//   * AVERAGE: filterKernelBaseline replicates the red channel (src/filter.cu:51-53), filterKernelTiled averages
//     R, G, B (:142-155) and always returns what the reference's cacheInput=false path returns (its cached path
//     reads the tile with the wrong stride, :66-67 vs :97, 130); GAUSSIAN / CROSS: DESIGN.md §3b; WAVELET needs the
//     SVGF context (rmd_svgf_frame_gbuffer) and is a no-op here that raises rmdCompatLastError();
//   * .w of the output is 0.
// The fast path for the same arithmetic with library-chosen geometry is rmd_filter_baseline / rmd_filter_tiled.
#pragma once
#ifndef RMD_COMPAT_FILTER_CUH
#define RMD_COMPAT_FILTER_CUH

#include "utils.h"
#include "extended_math.h"
#include "vector.h"
#include "image.h"
#include "gbuffer.h"

struct FilterParams {
    enum FilterType { AVERAGE, GAUSSIAN, CROSS, WAVELET } type;
    int depth;
    int level;
    int radius;
    float sigmaSpace;
    float sigmaColor;
    float sigmaAlbedo;
    float sigmaNormal;

    bool cacheInput = true;
    bool cacheBuffer = true;
};
static_assert(sizeof(FilterParams) == 36, "FilterParams must keep the reference layout (include/filter.cuh:11-23)");

KERNEL void filterKernelBaseline(GBuffer frame, const FilterParams params);
KERNEL void filterKernelTiled(GBuffer frame, const FilterParams params);

// Block-cooperative halo-tile copy (reference src/filter.cu:60-85), kept for sources that call it: `tile` receives the
// (blockDim.x + 2*radius) x (blockDim.y + 2*radius) window of `in` around the block, row stride blockDim.x + 2*radius
// (the stride the reference's consumer expects, :97, 130), zero outside the image; ends with __syncthreads().
CUDA_FUNC void cacheTile(uchar4* tile, uchar4* in, int2 shape, int radius) {
    const int tw = blockDim.x + 2 * radius, th = blockDim.y + 2 * radius;
    const int x0 = blockIdx.x * blockDim.x - radius, y0 = blockIdx.y * blockDim.y - radius;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < tw * th; i += blockDim.x * blockDim.y) {
        const int2 p = make_int2(x0 + i % tw, y0 + i / tw);
        tile[i] = inRange(p, shape) ? in[flattenIndex(p, shape)] : make_uchar4(0, 0, 0, 0);
    }
    __syncthreads();
}

// 0 when every compat kernel launched so far could serve its request, else the last RMD_E_* code (synchronises)
int rmdCompatLastError();

#endif
