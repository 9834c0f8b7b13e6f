// examples/harness.cpp — the reference's test harness flow (reference src/test.cu:16-48: list the registered
// cases, run those whose name matches a regular expression, time each with std::chrono, report
// "Passed with … ms" / "Fail with …") re-hosted on librmd_b200.so through include/rmd_compat.hpp.
//
// Cases: FILTER_BASELINE and FILTER_TILED are the two cases the reference ships (src/test.cu:69-90: 1920x1080,
// AVERAGE, depth 1, radius 2), with the caller-side <<<grid, block, smem>>> launches replaced by
// rmd_compat::filterBaseline / filterTiled.  SVGF_WAVELET is the case the reference's FilterParams::WAVELET was
// reserved for: the same GBuffer planes through the SVGF path (rmd_svgf_frame_gbuffer), eight frames so that the
// temporal history is exercised.
//
// Build (host compiler only):
//   g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/harness.cpp -o harness \
//       -Lraymarchdenoisercuda_b200 -lrmd_b200 -L/usr/local/cuda/lib64 -lcudart \
//       -Wl,-rpath,$PWD/raymarchdenoisercuda_b200 -Wl,-rpath,/usr/local/cuda/lib64
// Run:  ./harness [regex]      (default ".*"),   ./harness --list  (no GPU needed)
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <regex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rmd_compat.hpp"

namespace {

using Case = std::pair<std::string, std::function<void()>>;
std::vector<Case>& cases() {
    static std::vector<Case> v;
    return v;
}
struct Register {
    Register(const char* name, std::function<void()> fn) { cases().emplace_back(name, std::move(fn)); }
};

void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

constexpr int kW = 1920, kH = 1080;  // the reference harness's shape (src/test.cu:64)

// device planes of one frame, filled with a deterministic pattern (the reference leaves them uninitialised)
struct Planes {
    uchar4 *render = nullptr, *denoised = nullptr, *normal = nullptr, *albedo = nullptr;
    Planes() {
        const size_t n = (size_t)kW * kH;
        std::vector<uchar4> h(n);
        for (uchar4** p : {&render, &denoised, &normal, &albedo}) cuda_ok(cudaMalloc((void**)p, n * 4), "cudaMalloc");
        auto fill = [&](uchar4* d, auto f) {
            for (int y = 0; y < kH; ++y)
                for (int x = 0; x < kW; ++x) h[(size_t)y * kW + x] = f(x, y);
            cuda_ok(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
        };
        fill(render, [](int x, int y) {
            const unsigned s = (unsigned)(x * 73856093) ^ (unsigned)(y * 19349663);  // checker + hash noise
            const unsigned char base = ((x / 64 + y / 64) & 1) ? 200 : 60;
            return make_uchar4((unsigned char)(base + (s >> 8) % 48), (unsigned char)(base + (s >> 16) % 48),
                               (unsigned char)(base + (s >> 24) % 48), 255);
        });
        fill(albedo, [](int x, int y) { return ((x / 64 + y / 64) & 1) ? make_uchar4(230, 230, 230, 255) : make_uchar4(120, 90, 60, 255); });
        fill(normal, [](int x, int) { return x < kW / 2 ? make_uchar4(0, 0, 255, 255) : make_uchar4(255, 0, 0, 255); });
    }
    ~Planes() { for (uchar4* p : {render, denoised, normal, albedo}) cudaFree(p); }
    rmd_compat::GBuffer gbuffer() const {
        rmd_compat::GBuffer g;
        g.shape = {kW, kH};
        g.render = render; g.denoised = denoised; g.normal = normal; g.albedo = albedo;
        return g;
    }
};
Planes& planes() {
    static Planes p;
    return p;
}

Register r1("FILTER_BASELINE", [] {
    rmd_compat::FilterParams p;
    p.type = rmd_compat::FilterParams::AVERAGE; p.depth = 1; p.radius = 2;
    rmd_compat::filterBaseline(planes().gbuffer(), p);
    cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
});

Register r2("FILTER_TILED", [] {
    rmd_compat::FilterParams p;
    p.type = rmd_compat::FilterParams::AVERAGE; p.depth = 1; p.radius = 2;
    rmd_compat::filterTiled(planes().gbuffer(), p);
    cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
});

Register r3("SVGF_WAVELET", [] {
    rmd_compat::SvgfContext ctx(kW, kH);
    rmd_compat::FilterParams p;
    p.type = rmd_compat::FilterParams::WAVELET; p.depth = 5; p.radius = 2;
    for (int f = 0; f < 8; ++f) ctx.frame(planes().gbuffer(), p);
    cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    uchar4 px;
    cuda_ok(cudaMemcpy(&px, planes().denoised + (size_t)(kH / 2) * kW + kW / 4, 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    if (px.w != 255) throw std::runtime_error("denoised plane was not written");
});

}  // namespace

int main(int argc, char** argv) {
    const std::string pattern = argc > 1 ? argv[1] : ".*";
    const char* rule = "----------------------------------------------------------\n";
    std::printf("%s%zu available tests: ", rule, cases().size());
    for (auto& c : cases()) std::printf("%s ", c.first.c_str());
    std::printf("\n%s", rule);
    if (pattern == "--list") return 0;
    const std::regex re(pattern);
    int failed = 0;
    for (auto& [name, fn] : cases()) {
        if (!std::regex_match(name, re)) continue;
        try {
            std::printf("TEST %s:\n", name.c_str());
            const auto t0 = std::chrono::high_resolution_clock::now();
            fn();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
            std::printf("Passed with %.3f ms\n", ms);
        } catch (const std::runtime_error& e) {
            std::printf("Fail with %s\n", e.what());
            ++failed;
        } catch (...) {
            std::printf("Failed\n");
            ++failed;
        }
        std::printf("%s", rule);
    }
    return failed ? 1 : 0;
}
