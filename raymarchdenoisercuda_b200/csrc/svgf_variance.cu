// svgf_variance.cu — pass 2 of the SVGF path: 7x7 cross-bilateral estimate of colour,
// moments and variance for pixels whose history is shorter than `short_history`
// (DESIGN.md spec S3; oracle/oracle_svgf.c:pass_variance).  No reference counterpart.
//
// The pass is sparse in steady state (only disoccluded regions qualify), so it is
// driven by the per-tile flags the temporal kernel wrote: an un-flagged CTA exits
// after one 4-byte load.  Results go to side planes and a second tiny kernel
// patches them into the temporal output, which keeps every 7x7 read on the
// un-modified temporal output (Jacobi semantics, as the oracle).
//
// Roofline: HBM.  Worst case (every pixel qualifies, e.g. the first frames of a
// sequence): estimate reads colour 16 + moments 8 + guide 16 + slope 4 + histlen 1
// and writes 16 + 4; patch reads 16 + 4 + 1 and writes 16 + 4 => 106 B/px.
// Steady state: 4 B per 32x8 tile.
#include "svgf.cuh"

namespace rmd {
namespace {

__global__ void __launch_bounds__(kTemporalBx* kTemporalBy) variance_kernel(const VarianceArgs a) {
    if (a.tile_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0u) return;
    const int x = blockIdx.x * kTemporalBx + threadIdx.x;
    const int y = blockIdx.y * kTemporalBy + threadIdx.y;
    const int W = a.W, H = a.H, Wp = a.Wp;
    if (x >= W || y >= H) return;
    const size_t p = (size_t)y * Wp + x;
    const float4 gp = a.g4[p];
    if (gp.w == 0.0f) return;
    const int Nn = a.n[p];
    if (Nn >= a.k.short_hist) return;
    const float4 cp = a.c4[p];
    const float2 mp = a.m[p];
    const float kLog2e = 1.4426950408889634f;
    const float zs = a.k.sigma_z * fmaxf(a.dz[p], 1e-8f);
    const float il = kLog2e / a.k.lscale;
    float sw = 1.0f, sr = cp.x, sg = cp.y, sb = cp.z, s0 = mp.x, s1 = mp.y;
    for (int dx = -3; dx <= 3; ++dx) {
        const int qx = x + dx;
        if (qx < 0 || qx >= W) continue;
#pragma unroll
        for (int dy = -3; dy <= 3; ++dy) {
            const int qy = y + dy;
            if ((dx == 0 && dy == 0) || qy < 0 || qy >= H) continue;
            const size_t q = (size_t)qy * Wp + qx;
            const float4 gq = __ldg(a.g4 + q);
            const float d = fmaxf(fmaf(gp.z, gq.z, fmaf(gp.y, gq.y, gp.x * gq.x)), 0.0f);
            const float dist = sqrtf((float)(dx * dx + dy * dy));
            const float iz = kLog2e / fmaf(zs, dist, 1e-6f);
            const float4 cq = __ldg(a.c4 + q);
            float e = a.k.sigma_n * fast_lg2(d);
            e = fmaf(-fabsf(gp.w - gq.w), iz, e);
            e = fmaf(-fabsf(cp.w - cq.w), il, e);
            const float w = fast_ex2(e);  // sky / back-facing taps: d = 0 -> lg2 = -inf -> w = 0
            const float2 mq = __ldg(a.m + q);
            sw += w;
            sr = fmaf(w, cq.x, sr); sg = fmaf(w, cq.y, sg); sb = fmaf(w, cq.z, sb);
            s0 = fmaf(w, mq.x, s0); s1 = fmaf(w, mq.y, s1);
        }
    }
    const float inv = 1.0f / fmaxf(sw, 1e-6f);
    const float r = sr * inv, g = sg * inv, b = sb * inv, m0 = s0 * inv, m1 = s1 * inv;
    const float var = fmaxf(0.0f, m1 - m0 * m0) * (4.0f / (float)Nn);
    a.side_c4[p] = make_float4(r, g, b, luminance(r, g, b));
    a.side_v[p] = var;
}

__global__ void __launch_bounds__(kTemporalBx* kTemporalBy) variance_patch_kernel(const VarianceArgs a) {
    if (a.tile_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0u) return;
    const int x = blockIdx.x * kTemporalBx + threadIdx.x;
    const int y = blockIdx.y * kTemporalBy + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const size_t p = (size_t)y * a.Wp + x;
    if (a.g4[p].w == 0.0f || a.n[p] >= a.k.short_hist) return;
    a.patch_c4[p] = a.side_c4[p];
    a.patch_v[p] = a.side_v[p];
}

}  // namespace

int launch_variance(const VarianceArgs& a, cudaStream_t s) {
    dim3 block(kTemporalBx, kTemporalBy);
    dim3 grid((a.W + kTemporalBx - 1) / kTemporalBx, (a.H + kTemporalBy - 1) / kTemporalBy);
    variance_kernel<<<grid, block, 0, s>>>(a);
    variance_patch_kernel<<<grid, block, 0, s>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace rmd
