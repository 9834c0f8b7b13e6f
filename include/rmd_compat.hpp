// rmd_compat.hpp — C++ host mirror of the reference's interface for the denoise path, over the
// C ABI of rmd_b200.h.  Header-only; needs only the CUDA runtime headers.
//
// Mirrors (same names, argument meaning and error behaviour):
//   struct FilterParams         reference include/filter.cuh:11-23
//   struct GBuffer              reference include/gbuffer.h:6-14
//   struct CudaGBuffer          reference include/gbuffer.h:20-33 — declared there (ctor, dtor, allocate,
//                               openImages(stream)) and defined nowhere; implemented here on raw planes
//   filterBaseline/filterTiled  host launchers replacing the caller-side <<<grid, block, smem>>> launches
//                               of filterKernelBaseline / filterKernelTiled (reference src/test.cu:73-75, 85-87)
// Errors become std::runtime_error, which is what the reference's harness catches (src/test.cu:40-42).
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "rmd_b200.h"

namespace rmd_compat {

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + rmd_error_string(rc));
}

struct FilterParams {
    enum FilterType { AVERAGE, GAUSSIAN, CROSS, WAVELET } type = AVERAGE;
    int depth = 0;
    int level = 0;
    int radius = 0;
    float sigmaSpace = 0, sigmaColor = 0, sigmaAlbedo = 0, sigmaNormal = 0;
    bool cacheInput = true;
    bool cacheBuffer = true;
    RmdFilterParams c() const {
        return RmdFilterParams{(int32_t)type, depth, level, radius, sigmaSpace, sigmaColor, sigmaAlbedo, sigmaNormal,
                               (uint8_t)cacheInput, (uint8_t)cacheBuffer};
    }
};
static_assert(sizeof(FilterParams) == sizeof(RmdFilterParams), "FilterParams must stay bit-compatible");

struct GBuffer {  // non-owning view, reference include/gbuffer.h:6-14
    int2 shape{0, 0};
    uchar4* render = nullptr;
    uchar4* denoised = nullptr;
    uchar4* normal = nullptr;
    uchar4* albedo = nullptr;
    uchar4* buffer[2] = {nullptr, nullptr};
    RmdGBuffer c() const { return RmdGBuffer{shape.x, shape.y, render, denoised, normal, albedo, {buffer[0], buffer[1]}}; }
};
static_assert(sizeof(GBuffer) == sizeof(RmdGBuffer), "GBuffer must stay bit-compatible");

inline void filterBaseline(const GBuffer& frame, const FilterParams& params, cudaStream_t stream = 0) {
    RmdGBuffer g = frame.c();
    RmdFilterParams p = params.c();
    check(rmd_filter_baseline(&g, &p, stream), "filterBaseline");
}
inline void filterTiled(const GBuffer& frame, const FilterParams& params, cudaStream_t stream = 0) {
    RmdGBuffer g = frame.c();
    RmdFilterParams p = params.c();
    check(rmd_filter_tiled(&g, &p, stream), "filterTiled");
}

// Owner of the device planes of the legacy path (what the reference sketched as CudaGBuffer).
struct CudaGBuffer : GBuffer {
    uchar4* denoisedCPU = nullptr;  // pinned host copy of `denoised` (reference include/gbuffer.h:24)
    CudaGBuffer() = default;
    explicit CudaGBuffer(int2 s) { allocate(s); }
    CudaGBuffer(const CudaGBuffer&) = delete;
    CudaGBuffer& operator=(const CudaGBuffer&) = delete;
    ~CudaGBuffer() { release(); }

    void allocate(int2 s) {
        release();
        shape = s;
        const size_t bytes = (size_t)s.x * s.y * sizeof(uchar4);
        uchar4** planes[] = {&render, &denoised, &normal, &albedo, &buffer[0], &buffer[1]};
        for (uchar4** p : planes)
            if (cudaMalloc((void**)p, bytes) != cudaSuccess) throw std::runtime_error("CudaGBuffer::allocate: cudaMalloc failed");
        if (cudaMallocHost((void**)&denoisedCPU, bytes) != cudaSuccess) throw std::runtime_error("CudaGBuffer::allocate: cudaMallocHost failed");
    }
    // Uploads RGBA8 host planes (what Image(path, 4) yields, reference src/image.cpp:33-40); null = leave as is.
    void upload(const void* renderRGBA, const void* albedoRGBA, const void* normalRGBA, cudaStream_t stream = 0) {
        const size_t bytes = (size_t)shape.x * shape.y * sizeof(uchar4);
        if (renderRGBA) cudaMemcpyAsync(render, renderRGBA, bytes, cudaMemcpyHostToDevice, stream);
        if (albedoRGBA) cudaMemcpyAsync(albedo, albedoRGBA, bytes, cudaMemcpyHostToDevice, stream);
        if (normalRGBA) cudaMemcpyAsync(normal, normalRGBA, bytes, cudaMemcpyHostToDevice, stream);
    }
    void download(cudaStream_t stream = 0) {
        cudaMemcpyAsync(denoisedCPU, denoised, (size_t)shape.x * shape.y * sizeof(uchar4), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
    }

  private:
    void release() {
        uchar4* planes[] = {render, denoised, normal, albedo, buffer[0], buffer[1]};
        for (uchar4* p : planes)
            if (p) cudaFree(p);
        if (denoisedCPU) cudaFreeHost(denoisedCPU);
        render = denoised = normal = albedo = buffer[0] = buffer[1] = denoisedCPU = nullptr;
    }
};

// Per-sequence SVGF context (RAII over rmd_svgf_create / rmd_svgf_destroy).
class SvgfContext {
  public:
    SvgfContext(int width, int height, int device = 0) { check(rmd_svgf_create(&ctx_, width, height, device), "rmd_svgf_create"); }
    ~SvgfContext() { if (ctx_) rmd_svgf_destroy(ctx_); }
    SvgfContext(const SvgfContext&) = delete;
    SvgfContext& operator=(const SvgfContext&) = delete;
    void reset() { check(rmd_svgf_reset(ctx_), "rmd_svgf_reset"); }
    void frame(const RmdSvgfFrame& f, const FilterParams& p, const RmdSvgfParams* sp = nullptr, cudaStream_t stream = 0) {
        RmdFilterParams c = p.c();
        check(rmd_svgf_frame(ctx_, &f, &c, sp, stream), "rmd_svgf_frame");
    }
    void frameHost(const RmdSvgfFrame& f, const FilterParams& p, const RmdSvgfParams* sp = nullptr) {
        RmdFilterParams c = p.c();
        check(rmd_svgf_frame_host(ctx_, &f, &c, sp), "rmd_svgf_frame_host");
    }
    // The reference's own GBuffer (RGBA8 planes) through SVGF: what filterKernel*(frame, {.type = WAVELET}) was reserved for.
    void frame(const GBuffer& g, const FilterParams& p, const RmdSvgfParams* sp = nullptr, cudaStream_t stream = 0) {
        RmdGBuffer cg = g.c();
        RmdFilterParams c = p.c();
        check(rmd_svgf_frame_gbuffer(ctx_, &cg, &c, sp, nullptr, stream), "rmd_svgf_frame_gbuffer");
    }
    void hostWait() { check(rmd_svgf_host_wait(ctx_), "rmd_svgf_host_wait"); }
    rmd_svgf_ctx* raw() { return ctx_; }

  private:
    rmd_svgf_ctx* ctx_ = nullptr;
};

}  // namespace rmd_compat
