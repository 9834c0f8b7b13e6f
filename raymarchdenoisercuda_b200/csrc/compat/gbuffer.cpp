// gbuffer.cpp — CudaGBuffer (include/compat/gbuffer.h): the host-side owner of the device planes that the reference
// declares and never defines (reference include/gbuffer.h:20-33: ctor, dtor, allocate, openImages(filepath, stream)).
// Cold path: allocation and fixture upload, once per sequence.
#include "gbuffer.h"
#include "extended_math.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace {
bool file_exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (f) fclose(f);
    return f != nullptr;
}
}  // namespace

CudaGBuffer::CudaGBuffer(int2 shape) : GBuffer{} { allocate(shape); }

CudaGBuffer::~CudaGBuffer() {
    if (denoisedCPU) cudaFreeHost(denoisedCPU);
}

void CudaGBuffer::allocate(int2 s) {
    if (s.x <= 0 || s.y <= 0) throw std::runtime_error("CudaGBuffer::allocate: empty shape");
    const size_t n = (size_t)totalSize(s);
    shape = s;
    renderVec = CudaVector<uchar4>(n);
    albedoVec = CudaVector<uchar4>(n);
    normalVec = CudaVector<uchar4>(n);
    denoisedVec = CudaVector<uchar4>(n);
    bufferVec = CudaVector<uchar4>(2 * n);
    for (const CudaVector<uchar4>* v : {&renderVec, &albedoVec, &normalVec, &denoisedVec, &bufferVec})
        if (v->error() != cudaSuccess)
            throw std::runtime_error(std::string("CudaGBuffer::allocate: ") + cudaGetErrorString(v->error()));
    render = renderVec.data(); albedo = albedoVec.data(); normal = normalVec.data(); denoised = denoisedVec.data();
    buffer[0] = bufferVec.data(); buffer[1] = bufferVec.data() + n;
    if (denoisedCPU) { cudaFreeHost(denoisedCPU); denoisedCPU = nullptr; }
    RMD_CHECK_CUDA(cudaMallocHost((void**)&denoisedCPU, n * sizeof(uchar4)));
    depthVec = CudaVector<float>(); motionVec = CudaVector<float>();
    depth = motion = nullptr;
}

void CudaGBuffer::openImages(std::string filepath, cudaStream_t stream) {
    if (!filepath.empty() && filepath.back() != '/') filepath += '/';
    const char* names[3] = {"render.png", "albedo.png", "normal.png"};
    Image img[3];
    for (int i = 0; i < 3; ++i) img[i] = Image(filepath + names[i], 4);   // RGBA8, A = 255 for RGB files
    for (int i = 1; i < 3; ++i)
        if (img[i].shape.x != img[0].shape.x || img[i].shape.y != img[0].shape.y)
            throw std::runtime_error("CudaGBuffer::openImages: " + std::string(names[i]) + " differs in size from render.png");
    const int2 s = make_int2(img[0].shape.x, img[0].shape.y);
    if (s.x != shape.x || s.y != shape.y || !render) allocate(s);
    const size_t n = (size_t)totalSize(s);
    CudaVector<uchar4>* dst[3] = {&renderVec, &albedoVec, &normalVec};
    for (int i = 0; i < 3; ++i) {
        staging[i].resize(n);   // must outlive the asynchronous copy
        memcpy(staging[i].data(), img[i].data, n * sizeof(uchar4));
        dst[i]->copyFromAsync(staging[i].data(), n, stream);
    }
    std::vector<float> f;
    if (file_exists(filepath + "depth.npy")) {
        const int3 d = rmdLoadNpyFloat(filepath + "depth.npy", f);
        if (d.x != s.x || d.y != s.y || d.z != 1) throw std::runtime_error("CudaGBuffer::openImages: depth.npy must be (H,W) float32");
        depthVec = CudaVector<float>(f.data(), n);
        depth = depthVec.data();
    }
    if (file_exists(filepath + "motion.npy")) {
        const int3 d = rmdLoadNpyFloat(filepath + "motion.npy", f);
        if (d.x != s.x || d.y != s.y || d.z != 2) throw std::runtime_error("CudaGBuffer::openImages: motion.npy must be (H,W,2) float32");
        motionVec = CudaVector<float>(f.data(), 2 * n);
        motion = motionVec.data();
    }
}

uchar4* CudaGBuffer::download(cudaStream_t stream) {
    if (!denoised || !denoisedCPU) throw std::runtime_error("CudaGBuffer::download: not allocated");
    denoisedVec.copyToAsync(denoisedCPU, stream);
    RMD_CHECK_CUDA(cudaStreamSynchronize(stream));
    return denoisedCPU;
}
