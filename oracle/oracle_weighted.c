/*
 * oracle_weighted.c — CPU restatement of FilterParams::GAUSSIAN and FilterParams::CROSS on the reference's RGBA8
 * planes.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as oracle_box.c).
 *
 * Parity status: UNPINNED by the reference.  The reference enumerates the two types (include/filter.cuh:12) and
 * carries their knobs (sigmaSpace / sigmaColor / sigmaAlbedo / sigmaNormal, include/filter.cuh:16-19), but no kernel
 * reads either: there is no reference output to compare with.  What IS taken from the reference:
 * plane formats and ping-pong (src/filter.cu:24-25), tap order x outer / y inner (:34-35), "skip the tap outside the
 * image and renormalise" (:38-39), float accumulation and IEEE division (:48), .w = 0 (:151-155).  The quotient is
 * ROUNDED to the nearest code (the reference truncates its box sums, which are exact integers; a truncated weighted
 * mean would lose up to one code per level on a constant image).  The weights are the textbook Gaussian / cross-bilateral ones:
 *
 *   w(p,q) = 2^-( ks*|q-p|^2 + kc*|c_p-c_q|^2 + ka*|a_p-a_q|^2 + kn*|n_p-n_q|^2 )
 *   ks = log2(e) / (2 sigmaSpace^2)            sigmaSpace  = 0 => max(radius, 1) / 2
 *   kc = log2(e) / (2 sigmaColor^2  * 255^2)   sigmaColor  = 0 => term off   (CROSS only; colour of the level's input)
 *   ka = log2(e) / (2 sigmaAlbedo^2 * 255^2)   sigmaAlbedo = 0 => term off   (CROSS only; frame.albedo)
 *   kn = log2(e) / (2 sigmaNormal^2 * 255^2)   sigmaNormal = 0 => term off   (CROSS only; frame.normal)
 *
 * Every squared distance is an integer (8-bit planes), the four k's are rounded to float once, the exponent is an
 * explicit fmaf chain, and 2^x is the polynomial below — evaluated with the same fmaf sequence by the CUDA kernels
 * (csrc/weighted.cuh), so GPU and oracle agree BIT FOR BIT (fmaf and IEEE division are correctly rounded on both).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../include/rmd_b200.h"

/* 2^x for x <= 0: Cephes exp2f's degree-5 polynomial on [-0.5, 0.5], fmaf chain, exact power-of-two scaling */
static inline float exp2_neg(float x) {
    if (!(x >= -125.0f)) return 0.0f;
    float i = floorf(x);
    float f = x - i;
    if (f > 0.5f) { i += 1.0f; f -= 1.0f; }
    float p = 1.535336188319500e-4f;
    p = fmaf(p, f, 1.339887440266574e-3f);
    p = fmaf(p, f, 9.618437357674640e-3f);
    p = fmaf(p, f, 5.550332471162809e-2f);
    p = fmaf(p, f, 2.402264791363012e-1f);
    p = fmaf(p, f, 6.931472028550421e-1f);
    p = fmaf(p, f, 1.0f);
    union { uint32_t u; float v; } s;
    s.u = (uint32_t)((int)i + 127) << 23;
    return p * s.v;
}

typedef struct {
    float ks, kc, ka, kn;
} weights_k;

/* the four exponent scales, computed in double and rounded to float once (the product does the same on the host) */
int oracle_weighted_scales(const RmdFilterParams* p, float out[4]) {
    const double log2e = 1.4426950408889634;
    if (!p || (p->type != RMD_FILTER_GAUSSIAN && p->type != RMD_FILTER_CROSS)) return -1;
    const double ss = p->sigmaSpace > 0 ? p->sigmaSpace : 0.5 * (p->radius > 1 ? p->radius : 1);
    out[0] = (float)(log2e / (2.0 * ss * ss));
    out[1] = out[2] = out[3] = 0.0f;
    if (p->type == RMD_FILTER_CROSS) {
        if (p->sigmaColor > 0) out[1] = (float)(log2e / (2.0 * (double)p->sigmaColor * p->sigmaColor * 65025.0));
        if (p->sigmaAlbedo > 0) out[2] = (float)(log2e / (2.0 * (double)p->sigmaAlbedo * p->sigmaAlbedo * 65025.0));
        if (p->sigmaNormal > 0) out[3] = (float)(log2e / (2.0 * (double)p->sigmaNormal * p->sigmaNormal * 65025.0));
    }
    return 0;
}

static inline int dist2_rgb(const uint8_t* a, const uint8_t* b) {
    const int d0 = (int)a[0] - b[0], d1 = (int)a[1] - b[1], d2 = (int)a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

static void weighted_level(const uint8_t* in, uint8_t* out, const uint8_t* albedo, const uint8_t* normal, int W, int H,
                           int radius, weights_k k) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t p = 4 * ((size_t)y * W + x);
            float ar = 0.f, ag = 0.f, ab = 0.f, ws = 0.f;
            for (int dx = -radius; dx <= radius; ++dx)      /* x outer (src/filter.cu:34) */
                for (int dy = -radius; dy <= radius; ++dy) { /* y inner (:35) */
                    const int nx = x + dx, ny = y + dy;
                    if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue; /* :38-39 */
                    const size_t q = 4 * ((size_t)ny * W + nx);
                    float e = k.ks * (float)(dx * dx + dy * dy);
                    e = fmaf(k.kc, (float)dist2_rgb(in + p, in + q), e);
                    e = fmaf(k.ka, albedo ? (float)dist2_rgb(albedo + p, albedo + q) : 0.0f, e);
                    e = fmaf(k.kn, normal ? (float)dist2_rgb(normal + p, normal + q) : 0.0f, e);
                    const float w = exp2_neg(-e);
                    ar = fmaf(w, (float)in[q + 0], ar);
                    ag = fmaf(w, (float)in[q + 1], ag);
                    ab = fmaf(w, (float)in[q + 2], ab);
                    ws = ws + w;
                }
            /* IEEE division (:48); rounded to nearest, not truncated like the box sums (:51-53): truncating a
             * weighted mean would darken a constant image by one code per level */
            out[p + 0] = (uint8_t)(ar / ws + 0.5f);
            out[p + 1] = (uint8_t)(ag / ws + 0.5f);
            out[p + 2] = (uint8_t)(ab / ws + 0.5f);
            out[p + 3] = 0;
        }
}

/* depth levels with the reference's ping-pong (src/filter.cu:24-25); albedo / normal may be NULL when their term is off */
int oracle_weighted_filter(const uint8_t* render, uint8_t* denoised, uint8_t* buf0, uint8_t* buf1, const uint8_t* albedo,
                           const uint8_t* normal, int W, int H, const RmdFilterParams* p) {
    float s[4];
    if (oracle_weighted_scales(p, s)) return -1;
    if (p->depth < 1 || p->radius < 0) return -1;
    if (p->depth > 1 && (!buf0 || !buf1)) return -1;
    weights_k k = {s[0], s[1], s[2], s[3]};
    if ((k.ka > 0 && !albedo) || (k.kn > 0 && !normal)) return -1;
    uint8_t* buffer[2] = {buf0, buf1};
    for (int level = 0; level < p->depth; ++level) {
        const uint8_t* in = level == 0 ? render : buffer[level % 2];
        uint8_t* out = level == p->depth - 1 ? denoised : buffer[(level + 1) % 2];
        weighted_level(in, out, k.ka > 0 ? albedo : NULL, k.kn > 0 ? normal : NULL, W, H, p->radius, k);
    }
    return 0;
}

/* exposed for tests/test_oracle_weighted.py: accuracy of the shared 2^x polynomial */
float oracle_exp2_neg(float x) { return exp2_neg(x); }
