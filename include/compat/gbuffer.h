// include/compat/gbuffer.h — `GBuffer` with the reference's exact layout (include/gbuffer.h:6-14: int2 shape + six
// uchar4 device pointers, 56 bytes) and a WORKING `CudaGBuffer` (include/gbuffer.h:20-33 declares ctor, dtor,
// allocate and openImages(filepath, stream) and defines none of them).  Same member names; extended with
// the planes the SVGF path needs and the reference lacks (depth, motion) and with the readback its TODO hints at
// (`denoisedCPU`).
#pragma once
#ifndef RMD_COMPAT_GBUFFER_H
#define RMD_COMPAT_GBUFFER_H

#include "image.h"

struct GBuffer {
    int2 shape;

    uchar4* render;
    uchar4* denoised;
    uchar4* normal;
    uchar4* albedo;
    uchar4* buffer[2];
};
static_assert(sizeof(GBuffer) == 56, "GBuffer must keep the reference layout (include/gbuffer.h:6-14)");

struct CPUGBuffer : GBuffer {
    Image render, albedo, normal;
};

struct CudaGBuffer : GBuffer {
    CudaVector<uchar4> renderVec, albedoVec, normalVec, denoisedVec;
    CudaVector<uchar4> bufferVec;  // both ping-pong planes, back to back
    uchar4* denoisedCPU = nullptr; // pinned host copy of `denoised`, filled by download()
    // beyond the reference: optional linear depth (fp32) and motion (2 x fp32, pixels, prev = p + motion) planes
    CudaVector<float> depthVec, motionVec;
    float* depth = nullptr;
    float* motion = nullptr;

    CudaGBuffer() : GBuffer{} {}
    explicit CudaGBuffer(int2 shape);
    CudaGBuffer(const CudaGBuffer&) = delete;
    CudaGBuffer& operator=(const CudaGBuffer&) = delete;
    ~CudaGBuffer();

    void allocate(int2 shape);
    // Loads <filepath>/render.png, albedo.png, normal.png (the layout of the reference's render/cornell/1/) as RGBA8
    // and uploads them on `stream` (allocating for the images' size when needed).  Optional extras in the same
    // directory: depth.npy (H,W float32) and motion.npy (H,W,2 float32).  Throws std::runtime_error when a mandatory
    // file is missing or the sizes disagree.
    void openImages(std::string filepath, cudaStream_t stream = 0);
    // denoised -> denoisedCPU on `stream`; synchronises the stream and returns denoisedCPU
    uchar4* download(cudaStream_t stream = 0);

   private:
    std::vector<uchar4> staging[3];  // host copies must outlive the asynchronous uploads
};

#endif
