// weighted_filter.cu — FilterParams::GAUSSIAN and FilterParams::CROSS for the reference's RGBA8 G-buffer, behind the
// same two entry points as the box path (rmd_filter_baseline / rmd_filter_tiled dispatch on params->type).
//
// Reference hooks: enum and sigma fields include/filter.cuh:12-19 (enumerated, never read by a reference kernel);
// plane formats include/gbuffer.h:6-14; ping-pong, tap order, border rule, division and truncation
// src/filter.cu:24-25, 34-35, 38-39, 48-53.  Arithmetic: csrc/weighted.cuh == oracle/oracle_weighted.c, bit for bit.
//
// Kernel: a CTA of 32x8 threads produces a 32x8 tile; the (32+2r)x(8+2r) neighbourhood of the colour plane (and of
// the albedo / normal planes when their term is on) is staged in shared memory with coalesced loads, texels outside
// the image marked invalid; each thread walks its (2r+1)^2 taps.  Roofline: HBM, 8 B/px (+4 B/px per guide plane);
// the tap loop (about 25 instructions per tap) keeps it compute-bound for r >= 2.
#include "common.cuh"
#include "weighted.cuh"

namespace rmd {
namespace {

constexpr int kWfBx = 32, kWfBy = 8;

template <bool USE_A, bool USE_N>
__global__ void __launch_bounds__(kWfBx* kWfBy) weighted_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                 const uint32_t* __restrict__ albedo,
                                                                 const uint32_t* __restrict__ normal, int W, int H, int r,
                                                                 WeightScales k) {
    extern __shared__ uint32_t sm[];
    const int tw = kWfBx + 2 * r, th = kWfBy + 2 * r, n = tw * th;
    uint32_t* sC = sm;
    uint32_t* sA = sm + n;       // only touched when USE_A
    uint32_t* sN = sm + 2 * n;   // only touched when USE_N
    const int x0 = blockIdx.x * kWfBx - r, y0 = blockIdx.y * kWfBy - r;
    const int tid = threadIdx.y * kWfBx + threadIdx.x;
    for (int i = tid; i < n; i += kWfBx * kWfBy) {
        const int ty = i / tw, tx = i - ty * tw;
        const int gx = x0 + tx, gy = y0 + ty;
        const bool inside = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t q = (size_t)gy * W + gx;
        // bit 24..31 (the unused .w byte) carries "inside the image" while the texel sits in shared memory
        sC[i] = inside ? ((__ldg(in + q) & 0x00FFFFFFu) | 0x01000000u) : 0u;
        if (USE_A) sA[i] = inside ? __ldg(albedo + q) : 0u;
        if (USE_N) sN[i] = inside ? __ldg(normal + q) : 0u;
    }
    __syncthreads();
    const int x = blockIdx.x * kWfBx + threadIdx.x, y = blockIdx.y * kWfBy + threadIdx.y;
    if (x >= W || y >= H) return;
    const int ci = (threadIdx.y + r) * tw + threadIdx.x + r;
    const uint32_t cp = sC[ci], ap = USE_A ? sA[ci] : 0u, np = USE_N ? sN[ci] : 0u;
    WeightedAcc acc{0.f, 0.f, 0.f, 0.f};
    for (int dx = -r; dx <= r; ++dx)        // x outer, y inner (reference src/filter.cu:34-35)
        for (int dy = -r; dy <= r; ++dy) {
            const int qi = ci + dy * tw + dx;
            const uint32_t cq = sC[qi];
            if (!(cq >> 24)) continue;      // outside the image: skipped, not counted (src/filter.cu:38-39)
            weighted_tap(acc, k, dx * dx + dy * dy, cp, cq, ap, USE_A ? sA[qi] : 0u, np, USE_N ? sN[qi] : 0u);
        }
    out[(size_t)y * W + x] = weighted_finish(acc);
}

}  // namespace

int weighted_filter(const RmdGBuffer* f, const RmdFilterParams* p, cudaStream_t s) {
    if (!f || !p) return RMD_E_NULL;
    if (f->width <= 0 || f->height <= 0 || (long long)f->width * f->height > 0x7FFFFFFFLL) return RMD_E_SHAPE;
    if (p->radius < 0 || p->radius > RMD_BOX_MAX_RADIUS || p->depth < 1) return RMD_E_PARAM;
    if (p->sigmaSpace < 0 || p->sigmaColor < 0 || p->sigmaAlbedo < 0 || p->sigmaNormal < 0) return RMD_E_PARAM;
    WeightScales k;
    if (weight_scales_from_params(p->type, p->radius, p->sigmaSpace, p->sigmaColor, p->sigmaAlbedo, p->sigmaNormal, &k))
        return RMD_E_PARAM;
    if (!f->render || !f->denoised) return RMD_E_NULL;
    if (p->depth > 1 && (!f->buffer[0] || !f->buffer[1])) return RMD_E_NULL;
    const bool use_a = k.ka > 0.0f, use_n = k.kn > 0.0f;
    if ((use_a && !f->albedo) || (use_n && !f->normal)) return RMD_E_NULL;
    if (((uintptr_t)f->render | (uintptr_t)f->denoised | (uintptr_t)f->buffer[0] | (uintptr_t)f->buffer[1] |
         (uintptr_t)f->albedo | (uintptr_t)f->normal) & 3u)
        return RMD_E_ALIGN;
    const int tw = kWfBx + 2 * p->radius, th = kWfBy + 2 * p->radius;
    const size_t smem = (size_t)tw * th * 4 * 3;
    auto kern = use_a ? (use_n ? weighted_kernel<true, true> : weighted_kernel<true, false>)
                      : (use_n ? weighted_kernel<false, true> : weighted_kernel<false, false>);
    if (smem > 48 * 1024) RMD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((f->width + kWfBx - 1) / kWfBx, (f->height + kWfBy - 1) / kWfBy), block(kWfBx, kWfBy);
    for (int level = 0; level < p->depth; ++level) {
        const void* in = level == 0 ? f->render : f->buffer[level % 2];                 // src/filter.cu:24
        void* out = level == p->depth - 1 ? f->denoised : f->buffer[(level + 1) % 2];   // src/filter.cu:25
        kern<<<grid, block, smem, s>>>((const uint32_t*)in, (uint32_t*)out, (const uint32_t*)f->albedo,
                                       (const uint32_t*)f->normal, f->width, f->height, p->radius, k);
    }
    return (int)cudaGetLastError();
}

}  // namespace rmd
