// svgf_variance.cu — pass 2 of the SVGF path: 7x7 cross-bilateral estimate of colour,
// moments and variance for pixels whose history is shorter than `short_history`
// (DESIGN.md spec S3; oracle/oracle_svgf.c:pass_variance).  No reference counterpart.
//
// The pass is sparse in steady state (only disoccluded regions qualify), so it is
// driven by the per-tile flags the temporal kernel wrote: an un-flagged CTA exits
// after one 4-byte load.  A flagged CTA compacts its qualifying pixels into a
// shared-memory list and gives ONE pixel to each of the first `count` threads, so
// the 49-tap loop runs on dense warps instead of on the scattered lanes that
// happen to qualify.
// Jacobi semantics without a second pass: the temporal kernel also stores the
// colour of every short-history pixel in a side plane that nobody modifies; a tap
// reads the side plane when the neighbour is itself a short-history pixel (which
// this pass may already have overwritten) and the temporal plane otherwise, and
// the result is written in place.
//
// Roofline: HBM.  Worst case (every pixel qualifies, e.g. the first frames of a
// sequence): reads side colour 16 + moments 8 + guide 16 + slope 4 + histlen 1,
// writes colour 16 + variance 4 => 65 B/px (+16 B/px the temporal pass spent on
// the side plane).  Steady state: 4 B per 32x8 tile plus the disoccluded pixels.
#include "svgf.cuh"

namespace rmd {
namespace {

__global__ void __launch_bounds__(kTemporalBx* kTemporalBy) variance_kernel(const VarianceArgs a) {
    if (a.tile_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0u) return;
    __shared__ int s_count;
    __shared__ unsigned short s_list[kTemporalBx * kTemporalBy];
    const int W = a.W, H = a.H, Wp = a.Wp;
    const int tid = threadIdx.y * kTemporalBx + threadIdx.x;
    if (tid == 0) s_count = 0;
    __syncthreads();
    {   // compaction: which pixels of the tile take the spatial estimate?
        const int x = blockIdx.x * kTemporalBx + threadIdx.x, y = blockIdx.y * kTemporalBy + threadIdx.y;
        bool need = false;
        if (x < W && y < H) {
            const size_t p = (size_t)y * Wp + x;
            need = a.n[p] < a.k.short_hist && a.g4[p].w != 0.0f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        int base = 0;
        if (threadIdx.x == 0 && m) base = atomicAdd(&s_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (need) s_list[base + __popc(m & ((1u << threadIdx.x) - 1u))] = (unsigned short)tid;
    }
    __syncthreads();
    if (tid >= s_count) return;
    const int id = s_list[tid];
    const int x = blockIdx.x * kTemporalBx + (id & (kTemporalBx - 1)), y = blockIdx.y * kTemporalBy + id / kTemporalBx;
    const size_t p = (size_t)y * Wp + x;
    const float4 gp = a.g4[p];
    const int Nn = a.n[p];
    const float4 cp = a.side_c4[p];  // untouched temporal output of this (short-history) pixel
    const float2 mp = a.m[p];
    const float kLog2e = 1.4426950408889634f;
    const float zs = a.k.sigma_z * fmaxf(a.dz[p], 1e-8f);
    const float il = kLog2e / a.k.lscale;
    const int short_hist = a.k.short_hist;
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
            // a short-history neighbour may already have been overwritten in place: read its side copy
            const bool q_short = a.n[q] < short_hist && gq.w != 0.0f;
            const float4 cq = q_short ? a.side_c4[q] : a.c4[q];
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
    a.patch_c4[p] = make_float4(r, g, b, luminance(r, g, b));
    a.patch_v[p] = var;
}

}  // namespace

int launch_variance(const VarianceArgs& a, cudaStream_t s) {
    dim3 block(kTemporalBx, kTemporalBy);
    dim3 grid((a.W + kTemporalBx - 1) / kTemporalBx, (a.H + kTemporalBy - 1) / kTemporalBy);
    variance_kernel<<<grid, block, 0, s>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace rmd
