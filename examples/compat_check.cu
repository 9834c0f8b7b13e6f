// examples/compat_check.cu — exercises the source-level drop-in exactly the way the reference's own sources use it:
// `#include "filter.cuh"` and `filterKernelBaseline<<<grid, block, smem>>>(GBuffer{...}, FilterParams{...})`
// (reference src/test.cu:70-75, 82-87), linked against librmd_compat.a instead of the reference's src/filter.cu.
// Unlike the reference's harness it reads results back: every case's output plane is written to <out>/<name>.bin
// for tests/test_compat_dropin.py to compare with the oracle.
//
//   compat_check kernels <in.npy (H,W,4) uint8> <albedo.npy> <normal.npy> <outdir>
//   compat_check gbuffer <dir with render.png albedo.png normal.png> <outdir>      CudaGBuffer::openImages + SVGF
#include "filter.cuh"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "rmd_b200.h"

namespace {

void write_bin(const std::string& path, const void* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)p, (std::streamsize)n);
    if (!f) throw std::runtime_error("cannot write " + path);
}

struct Case {
    const char* name;
    bool baseline;
    dim3 block;
    unsigned smem;
    FilterParams params;
};

int run_kernels(const std::string& in_path, const std::string& al_path, const std::string& no_path, const std::string& out) {
    Image render(in_path, 4), albedo(al_path, 4), normal(no_path, 4);
    const int2 shape = make_int2(render.shape.x, render.shape.y);
    const size_t n = (size_t)totalSize(shape);
    CudaVector<uchar4> d_in((uchar4*)render.data, n), d_al((uchar4*)albedo.data, n), d_no((uchar4*)normal.data, n);
    CudaVector<uchar4> d_out(n), b0(n), b1(n);
    std::vector<uchar4> host(n);
    const Case cases[] = {
        // the reference's own two launches (src/test.cu:70-75, 82-87)
        {"baseline_ref_call", true, dim3(16, 16), 49152, {.type = FilterParams::AVERAGE, .depth = 1, .radius = 2}},
        {"tiled_ref_call", false, dim3(16, 16), 30 * 1024, {.type = FilterParams::AVERAGE, .depth = 1, .radius = 2}},
        // no shared memory at all: taps from global memory
        {"tiled_no_smem", false, dim3(32, 8), 0, {.type = FilterParams::AVERAGE, .depth = 1, .radius = 2}},
        // five levels inside one launch (the reference's in-kernel loop races, src/filter.cu:56)
        {"tiled_depth5", false, dim3(16, 16), 30 * 1024, {.type = FilterParams::AVERAGE, .depth = 5, .radius = 2}},
        {"baseline_depth3", true, dim3(16, 16), 49152, {.type = FilterParams::AVERAGE, .depth = 3, .radius = 1}},
        // too little shared memory for the whole block: sub-tiles
        {"baseline_subtiles", true, dim3(8, 8), 1024, {.type = FilterParams::AVERAGE, .depth = 3, .radius = 1}},
        {"tiled_odd_block", false, dim3(13, 7), 4096, {.type = FilterParams::AVERAGE, .depth = 2, .radius = 2}},
        {"tiled_radius7", false, dim3(32, 4), 40000, {.type = FilterParams::AVERAGE, .depth = 1, .radius = 7}},
        // the two types the reference enumerates and never implements
        {"gaussian_depth2", false, dim3(16, 16), 30 * 1024, {.type = FilterParams::GAUSSIAN, .depth = 2, .radius = 2}},
        {"cross_depth2", false, dim3(16, 16), 40000,
         {.type = FilterParams::CROSS, .depth = 2, .radius = 3, .sigmaSpace = 1.5f, .sigmaColor = 0.25f, .sigmaAlbedo = 0.05f, .sigmaNormal = 0.3f}},
    };
    for (const Case& c : cases) {
        cudaMemset(d_out.data(), 0xAB, n * sizeof(uchar4));
        cudaMemset(b0.data(), 0xAB, n * sizeof(uchar4));
        cudaMemset(b1.data(), 0xAB, n * sizeof(uchar4));
        const dim3 grid((shape.x + c.block.x - 1) / c.block.x, (shape.y + c.block.y - 1) / c.block.y);
        GBuffer g{.shape = shape, .render = d_in.data(), .denoised = d_out.data(), .normal = d_no.data(), .albedo = d_al.data()};
        g.buffer[0] = b0.data();
        g.buffer[1] = b1.data();
        if (c.baseline) filterKernelBaseline<<<grid, c.block, c.smem>>>(g, c.params);
        else filterKernelTiled<<<grid, c.block, c.smem>>>(g, c.params);
        RMD_CHECK_CUDA(cudaGetLastError());
        RMD_CHECK_CUDA(cudaDeviceSynchronize());
        if (rmdCompatLastError() != 0) throw std::runtime_error(std::string(c.name) + ": rmdCompatLastError() != 0");
        d_out.copyTo(host.data());
        write_bin(out + "/" + c.name + ".bin", host.data(), n * sizeof(uchar4));
        printf("case %s ok\n", c.name);
    }
    // WAVELET cannot be served by a stateless launch: no-op + error code
    GBuffer g{.shape = shape, .render = d_in.data(), .denoised = d_out.data()};
    filterKernelTiled<<<dim3(1, 1), dim3(16, 16), 0>>>(g, {.type = FilterParams::WAVELET, .depth = 5, .radius = 2});
    printf("wavelet error code %d\n", rmdCompatLastError());
    return 0;
}

int run_gbuffer(const std::string& dir, const std::string& out) {
    CudaGBuffer gb;
    cudaStream_t stream;
    RMD_CHECK_CUDA(cudaStreamCreate(&stream));
    gb.openImages(dir, stream);   // render.png / albedo.png / normal.png -> device planes (async on `stream`)
    printf("gbuffer %d x %d depth %s motion %s\n", gb.shape.x, gb.shape.y, gb.depth ? "yes" : "no", gb.motion ? "yes" : "no");
    rmd_svgf_ctx* ctx = nullptr;
    if (rmd_svgf_create(&ctx, gb.shape.x, gb.shape.y, 0)) throw std::runtime_error("rmd_svgf_create failed");
    RmdGBuffer view{gb.shape.x, gb.shape.y, gb.render, gb.denoised, gb.normal, gb.albedo, {gb.buffer[0], gb.buffer[1]}};
    RmdFilterParams p{RMD_FILTER_WAVELET, 5, 0, 2, 0.f, 0.f, 0.f, 0.f, 1, 1};
    const int rc = rmd_svgf_frame_gbuffer(ctx, &view, &p, nullptr, nullptr, stream);
    if (rc) throw std::runtime_error(std::string("rmd_svgf_frame_gbuffer: ") + rmd_error_string(rc));
    const uchar4* host = gb.download(stream);   // denoisedCPU
    const size_t n = (size_t)totalSize(gb.shape);
    write_bin(out + "/svgf_denoised.bin", host, n * sizeof(uchar4));
    Image::save(out + "/svgf_denoised.png", (byte*)host, make_int3(gb.shape.x, gb.shape.y, 4));
    // the legacy AVERAGE path on the same buffer, through the kernels the reference's harness launches
    filterKernelTiled<<<dim3((gb.shape.x + 15) / 16, (gb.shape.y + 15) / 16), dim3(16, 16), 30 * 1024, stream>>>(
        gb, {.type = FilterParams::AVERAGE, .depth = 1, .radius = 2});
    RMD_CHECK_CUDA(cudaGetLastError());
    write_bin(out + "/box_denoised.bin", gb.download(stream), n * sizeof(uchar4));
    rmd_svgf_destroy(ctx);
    cudaStreamDestroy(stream);
    printf("gbuffer ok\n");
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    try {
        if (argc == 6 && !strcmp(argv[1], "kernels")) return run_kernels(argv[2], argv[3], argv[4], argv[5]);
        if (argc == 4 && !strcmp(argv[1], "gbuffer")) return run_gbuffer(argv[2], argv[3]);
        if (argc == 3 && !strcmp(argv[1], "png")) {   // decode + re-encode (no GPU): tests the PNG codec alone
            Image img(argv[2], 4);
            img.save(std::string(argv[2]) + ".roundtrip.png");
            printf("png %d x %d\n", img.shape.x, img.shape.y);
            return 0;
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "compat_check: %s\n", e.what());
        return 2;
    }
    fprintf(stderr, "usage: compat_check kernels <render.npy> <albedo.npy> <normal.npy> <outdir> | gbuffer <dir> <outdir> | png <file>\n");
    return 1;
}
