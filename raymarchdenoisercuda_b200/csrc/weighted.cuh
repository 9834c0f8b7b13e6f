// weighted.cuh — arithmetic of FilterParams::GAUSSIAN / FilterParams::CROSS on the reference's RGBA8 planes
// (reference include/filter.cuh:12, 16-19: enumerated and parameterised, read by none of its kernels).
// Normative definition: DESIGN.md §3b; CPU restatement: oracle/oracle_weighted.c.  Every operation below is an
// individually rounded IEEE operation (or an integer one), in the oracle's order, so that the CUDA result is
// bit-identical to the oracle's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmd {

struct WeightScales {
    float ks, kc, ka, kn;  // log2(e) / (2 sigma^2 [* 255^2]); 0 = term off
};

// 2^x for x <= 0: degree-5 polynomial on [-0.5, 0.5] (Cephes exp2f), fmaf chain, exact power-of-two scaling
__device__ __forceinline__ float exp2_neg(float x) {
    if (!(x >= -125.0f)) return 0.0f;
    float i = floorf(x);
    float f = __fsub_rn(x, i);
    if (f > 0.5f) { i = __fadd_rn(i, 1.0f); f = __fsub_rn(f, 1.0f); }
    float p = 1.535336188319500e-4f;
    p = __fmaf_rn(p, f, 1.339887440266574e-3f);
    p = __fmaf_rn(p, f, 9.618437357674640e-3f);
    p = __fmaf_rn(p, f, 5.550332471162809e-2f);
    p = __fmaf_rn(p, f, 2.402264791363012e-1f);
    p = __fmaf_rn(p, f, 6.931472028550421e-1f);
    p = __fmaf_rn(p, f, 1.0f);
    return __fmul_rn(p, __uint_as_float((uint32_t)((int)i + 127) << 23));
}

// sum of squared byte differences of the r,g,b bytes of two packed uchar4 texels
__device__ __forceinline__ int dist2_rgb(uint32_t a, uint32_t b) {
    const uint32_t d = __vabsdiffu4(a & 0x00FFFFFFu, b & 0x00FFFFFFu);
    return (int)__dp4a(d, d, 0u);
}

struct WeightedAcc {
    float r, g, b, w;
};

// one tap: centre texels (cp, ap, np), tap texels (cq, aq, nq), squared pixel distance d2s
__device__ __forceinline__ void weighted_tap(WeightedAcc& acc, const WeightScales& k, int d2s, uint32_t cp, uint32_t cq,
                                             uint32_t ap, uint32_t aq, uint32_t np, uint32_t nq) {
    float e = __fmul_rn(k.ks, (float)d2s);
    e = __fmaf_rn(k.kc, (float)dist2_rgb(cp, cq), e);
    e = __fmaf_rn(k.ka, (float)dist2_rgb(ap, aq), e);
    e = __fmaf_rn(k.kn, (float)dist2_rgb(np, nq), e);
    const float w = exp2_neg(-e);
    acc.r = __fmaf_rn(w, (float)(cq & 0xFFu), acc.r);
    acc.g = __fmaf_rn(w, (float)((cq >> 8) & 0xFFu), acc.g);
    acc.b = __fmaf_rn(w, (float)((cq >> 16) & 0xFFu), acc.b);
    acc.w = __fadd_rn(acc.w, w);
}

__device__ __forceinline__ uint32_t weighted_finish(const WeightedAcc& acc) {
    // IEEE division, rounded to the nearest code (see oracle_weighted.c: truncation would bias a weighted mean)
    const uint32_t r = (uint32_t)(unsigned char)__fadd_rn(__fdiv_rn(acc.r, acc.w), 0.5f);
    const uint32_t g = (uint32_t)(unsigned char)__fadd_rn(__fdiv_rn(acc.g, acc.w), 0.5f);
    const uint32_t b = (uint32_t)(unsigned char)__fadd_rn(__fdiv_rn(acc.b, acc.w), 0.5f);
    return r | (g << 8) | (b << 16);  // .w = 0 (reference src/filter.cu:151-155)
}

// host: the four exponent scales from the reference's knobs, computed in double and rounded once
inline int weight_scales_from_params(int type, int radius, float sigmaSpace, float sigmaColor, float sigmaAlbedo,
                                     float sigmaNormal, WeightScales* out) {
    const double log2e = 1.4426950408889634;
    if (type != 1 && type != 2) return -1;
    const double ss = sigmaSpace > 0 ? sigmaSpace : 0.5 * (radius > 1 ? radius : 1);
    out->ks = (float)(log2e / (2.0 * ss * ss));
    out->kc = out->ka = out->kn = 0.0f;
    if (type == 2) {
        if (sigmaColor > 0) out->kc = (float)(log2e / (2.0 * (double)sigmaColor * sigmaColor * 65025.0));
        if (sigmaAlbedo > 0) out->ka = (float)(log2e / (2.0 * (double)sigmaAlbedo * sigmaAlbedo * 65025.0));
        if (sigmaNormal > 0) out->kn = (float)(log2e / (2.0 * (double)sigmaNormal * sigmaNormal * 65025.0));
    }
    return 0;
}

}  // namespace rmd
