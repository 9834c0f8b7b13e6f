/*
 * oracle_svgf.c — CPU oracle of the SVGF denoise path (temporal accumulation,
 * variance estimation, edge-avoiding a-trous wavelet levels).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * Parity status: PARITY UNPINNED by the reference.  The reference announces SVGF
 * (README.md:3-10) but ships only an unweighted box filter (`float w = 1;`
 * src/filter.cu:41,127); it has no temporal, variance or a-trous code, no tests
 * with golden vectors (src/test.cu:68-90 never reads a result back) and no
 * third-party dependency containing the algorithm.  This file
 * therefore restates the PUBLISHED algorithm — Schied et al., "Spatiotemporal
 * Variance-Guided Filtering", HPG 2017; Dammertz et al., "Edge-Avoiding A-Trous
 * Wavelet Transform", HPG 2010 — as frozen in SURVEY.md Appendix A / DESIGN.md
 * "SVGF specification", anchored on the only artefacts the reference has for
 * it: the B3-spline taps `waveletSpline = {3/8, 1/4, 1/16}` (src/filter.cu:10),
 * the FilterParams knobs depth/radius/sigma* (include/filter.cuh:11-23), the
 * border rule "skip the tap and renormalise" (src/filter.cu:38-39, 46, 49), the
 * tap order x-outer/y-inner (src/filter.cu:34-35) and the row-major unpadded
 * plane layout (include/extended_math.h:66-68).  It is additionally pinned by
 * analytic known answers in tests/test_oracle_svgf.py (constant image is a fixed
 * point, an impulse reproduces the B3 footprint, orthogonal normals do not
 * bleed, static scene history converges to the running mean).
 *
 * Arithmetic contract (SURVEY §7 "Decision flips"): every PREDICATE (floor,
 * reprojection validity, thresholds, sky test) is evaluated in fp32 with the
 * exact operation order the CUDA kernels use (this file is compiled with
 * -ffp-contract=off; the kernels use __fmul_rn/__fadd_rn there), every
 * ACCUMULATION is done in double, and every plane is rounded to fp32 exactly
 * where the CUDA path stores it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rmd_b200_debug.h"
#include "oracle.h"

struct oracle_svgf {
    int W, H;
    int have_history;
    int parity;
    float* hist_c4;    /* float4: previous frame's level-0 output (rgb, L)       */
    float* mom[2];     /* float2 ping-pong                                        */
    uint8_t* hl[2];    /* history length ping-pong                                */
    float* g4[2];      /* float4 decoded guide ping-pong (nx, ny, nz, z)          */
    float* dz;         /* float depth slope of the current frame                  */
    float* t_c4_pre;   /* float4 temporal output before the variance pass         */
    float* t_v_pre;
    float* t_c4;       /* float4 after the variance pass                          */
    float* t_v;
    float* lv_c4[2];   /* level ping-pong                                         */
    float* lv_v[2];
};

static const float kLumR = 0.2126f, kLumG = 0.7152f, kLumB = 0.0722f;
/* reference src/filter.cu:10 */
static const double kSpline[3] = {3.0 / 8.0, 1.0 / 4.0, 1.0 / 16.0};

static inline double lum(double r, double g, double b) { return (double)kLumR * r + (double)kLumG * g + (double)kLumB * b; }

static inline float half_to_float(uint16_t h) {
    _Float16 v;
    memcpy(&v, &h, 2);
    return (float)v;
}

/* Guide decode, fp32, fixed operation order (mirrored by csrc/svgf_common.cuh:decode_guide). */
static inline void decode_guide(uint32_t w0, uint32_t w1, float out[4]) {
    float z;
    memcpy(&z, &w1, 4);
    if (!(z > 0.0f) || !isfinite(z)) { /* sky */
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    int sx = (int16_t)(w0 & 0xFFFFu), sy = (int16_t)(w0 >> 16);
    if (sx < -32767) sx = -32767;
    if (sy < -32767) sy = -32767;
    const float c = 1.0f / 32767.0f;
    float fx = (float)sx * c, fy = (float)sy * c;
    float fz = (1.0f - fabsf(fx)) - fabsf(fy);
    if (fz < 0.0f) {
        float ox = (1.0f - fabsf(fy)) * (fx >= 0.0f ? 1.0f : -1.0f);
        float oy = (1.0f - fabsf(fx)) * (fy >= 0.0f ? 1.0f : -1.0f);
        fx = ox; fy = oy;
    }
    float len2 = (fx * fx + fy * fy) + fz * fz;
    float inv = 1.0f / sqrtf(len2);
    out[0] = fx * inv; out[1] = fy * inv; out[2] = fz * inv; out[3] = z;
}

static inline float dot3f(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

oracle_svgf* oracle_svgf_create(int W, int H) {
    oracle_svgf* s = (oracle_svgf*)calloc(1, sizeof(*s));
    if (!s) return NULL;
    size_t n = (size_t)W * H;
    s->W = W; s->H = H;
    s->hist_c4 = (float*)calloc(n * 4, 4);
    for (int i = 0; i < 2; ++i) {
        s->mom[i] = (float*)calloc(n * 2, 4);
        s->hl[i] = (uint8_t*)calloc(n, 1);
        s->g4[i] = (float*)calloc(n * 4, 4);
        s->lv_c4[i] = (float*)calloc(n * 4, 4);
        s->lv_v[i] = (float*)calloc(n, 4);
    }
    s->dz = (float*)calloc(n, 4);
    s->t_c4_pre = (float*)calloc(n * 4, 4);
    s->t_v_pre = (float*)calloc(n, 4);
    s->t_c4 = (float*)calloc(n * 4, 4);
    s->t_v = (float*)calloc(n, 4);
    return s;
}

void oracle_svgf_destroy(oracle_svgf* s) {
    if (!s) return;
    free(s->hist_c4);
    for (int i = 0; i < 2; ++i) { free(s->mom[i]); free(s->hl[i]); free(s->g4[i]); free(s->lv_c4[i]); free(s->lv_v[i]); }
    free(s->dz); free(s->t_c4_pre); free(s->t_v_pre); free(s->t_c4); free(s->t_v);
    free(s);
}

void oracle_svgf_reset(oracle_svgf* s) { s->have_history = 0; }

typedef struct {
    float sigma_z, sigma_l, sigma_n;
    float alpha_c, alpha_m;
    int cap, short_hist;
    float dtol, nthr, afloor, lscale;
    int depth;
} resolved_params;

static int resolve(const RmdFilterParams* fp, const RmdSvgfParams* sp, resolved_params* r) {
    if (!fp) return RMD_E_NULL;
    if (fp->type != RMD_FILTER_WAVELET) return RMD_E_PARAM;
    if (fp->radius != 2) return RMD_E_PARAM;
    if (fp->depth < 0 || fp->depth > RMD_SVGF_MAX_LEVELS) return RMD_E_PARAM;
    r->depth = fp->depth;
    r->sigma_z = fp->sigmaSpace > 0 ? fp->sigmaSpace : 1.0f;
    r->sigma_l = fp->sigmaColor > 0 ? fp->sigmaColor : 4.0f;
    r->sigma_n = fp->sigmaNormal > 0 ? fp->sigmaNormal : 128.0f;
    r->alpha_c = sp && sp->alpha_color > 0 ? sp->alpha_color : 0.05f;
    r->alpha_m = sp && sp->alpha_moments > 0 ? sp->alpha_moments : 0.2f;
    r->cap = sp && sp->history_cap > 0 ? sp->history_cap : 32;
    if (r->cap > 255) r->cap = 255;
    r->short_hist = sp && sp->short_history > 0 ? sp->short_history : 4;
    r->dtol = sp && sp->depth_tolerance > 0 ? sp->depth_tolerance : 0.1f;
    r->nthr = sp && sp->normal_threshold > 0 ? sp->normal_threshold : 0.9f;
    r->afloor = sp && sp->albedo_floor > 0 ? sp->albedo_floor : 1e-3f;
    r->lscale = sp && sp->variance_lum_scale > 0 ? sp->variance_lum_scale : 10.0f;
    return 0;
}

/* ---- pass 0: guide decode + depth slope (DESIGN.md spec S1) --------------------------- */
static void pass_guide(oracle_svgf* s, const uint32_t* guide, float* g4) {
    const int W = s->W, H = s->H;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            decode_guide(guide[2 * p], guide[2 * p + 1], g4 + 4 * p);
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            float z = g4[4 * p + 3];
            if (z == 0.0f) { s->dz[p] = 0.0f; continue; }
            int xr = x + 1 < W ? x + 1 : x, yd = y + 1 < H ? y + 1 : y;
            float zx = g4[4 * ((size_t)y * W + xr) + 3], zy = g4[4 * ((size_t)yd * W + x) + 3];
            float a = fabsf(zx - z), b = fabsf(zy - z);
            s->dz[p] = a > b ? a : b;
        }
}

/* reprojection tap validity — fp32, same operation order as the kernel */
static inline int tap_valid(const float* g4prev, int W, int H, int tx, int ty, const float* gp, float rhs, float nthr) {
    if (tx < 0 || ty < 0 || tx >= W || ty >= H) return 0;
    const float* gq = g4prev + 4 * ((size_t)ty * W + tx);
    float lhs = fabsf(gq[3] - gp[3]);
    if (!(lhs <= rhs)) return 0;
    if (!(dot3f(gq, gp) >= nthr)) return 0;
    return 1;
}

/* ---- pass 1: temporal reprojection + accumulation (spec S2) ------------------------- */
static void pass_temporal(oracle_svgf* s, const RmdSvgfFrame* f, const resolved_params* r) {
    const int W = s->W, H = s->H;
    const uint16_t* color = (const uint16_t*)f->color;
    const uint8_t* albedo = (const uint8_t*)f->albedo;
    const uint16_t* motion = (const uint16_t*)f->motion;
    const int cur = s->parity, prv = s->parity ^ 1;
    const float* g4 = s->g4[cur];
    const float* g4p = s->g4[prv];
    const float* momp = s->mom[prv];
    const uint8_t* hlp = s->hl[prv];
    float* mom = s->mom[cur];
    uint8_t* hl = s->hl[cur];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            const float* gp = g4 + 4 * p;
            float c[3] = {half_to_float(color[4 * p]), half_to_float(color[4 * p + 1]), half_to_float(color[4 * p + 2])};
            float* oc = s->t_c4_pre + 4 * p;
            if (gp[3] == 0.0f) { /* sky: pass through */
                oc[0] = c[0]; oc[1] = c[1]; oc[2] = c[2];
                oc[3] = (float)lum(c[0], c[1], c[2]);
                s->t_v_pre[p] = 0.0f; mom[2 * p] = mom[2 * p + 1] = 0.0f; hl[p] = 0;
                continue;
            }
            float il[3];
            for (int k = 0; k < 3; ++k) {
                float a = (float)albedo[4 * p + k] * (1.0f / 255.0f);
                if (a < r->afloor) a = r->afloor;
                il[k] = c[k] / a;
            }
            double Lc = lum(il[0], il[1], il[2]);
            double mu[2] = {Lc, Lc * Lc};
            double Cp[3] = {il[0], il[1], il[2]}, Mp[2] = {mu[0], mu[1]};
            int N = 0;
            if (s->have_history) {
                float qx = (float)x + half_to_float(motion[2 * p]);
                float qy = (float)y + half_to_float(motion[2 * p + 1]);
                float q0x = floorf(qx), q0y = floorf(qy);
                float fx = qx - q0x, fy = qy - q0y;
                int ix = (int)q0x, iy = (int)q0y;
                float rhs = r->dtol * gp[3] + 2.0f * s->dz[p];
                float wt[4] = {(1.0f - fx) * (1.0f - fy), fx * (1.0f - fy), (1.0f - fx) * fy, fx * fy};
                const int ox[4] = {0, 1, 0, 1}, oy[4] = {0, 0, 1, 1};
                float sumw = 0.0f;
                int ok[4];
                for (int t = 0; t < 4; ++t) {
                    ok[t] = tap_valid(g4p, W, H, ix + ox[t], iy + oy[t], gp, rhs, r->nthr);
                    if (ok[t]) sumw += wt[t];
                }
                int rx = (int)floorf(qx + 0.5f), ry = (int)floorf(qy + 0.5f);
                int found = 0;
                if (sumw >= 0.01f) {
                    double ac[3] = {0, 0, 0}, am[2] = {0, 0};
                    for (int t = 0; t < 4; ++t)
                        if (ok[t]) {
                            size_t q = (size_t)(iy + oy[t]) * W + (ix + ox[t]);
                            for (int k = 0; k < 3; ++k) ac[k] += (double)wt[t] * s->hist_c4[4 * q + k];
                            am[0] += (double)wt[t] * momp[2 * q];
                            am[1] += (double)wt[t] * momp[2 * q + 1];
                        }
                    for (int k = 0; k < 3; ++k) Cp[k] = ac[k] / (double)sumw;
                    Mp[0] = am[0] / (double)sumw; Mp[1] = am[1] / (double)sumw;
                    found = 1;
                } else {
                    double ac[3] = {0, 0, 0}, am[2] = {0, 0};
                    int cnt = 0;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx)
                            if (tap_valid(g4p, W, H, rx + dx, ry + dy, gp, rhs, r->nthr)) {
                                size_t q = (size_t)(ry + dy) * W + (rx + dx);
                                for (int k = 0; k < 3; ++k) ac[k] += s->hist_c4[4 * q + k];
                                am[0] += momp[2 * q]; am[1] += momp[2 * q + 1];
                                ++cnt;
                            }
                    if (cnt > 0) {
                        for (int k = 0; k < 3; ++k) Cp[k] = ac[k] / cnt;
                        Mp[0] = am[0] / cnt; Mp[1] = am[1] / cnt;
                        found = 1;
                    }
                }
                if (found) N = hlp[(size_t)clampi(ry, 0, H - 1) * W + clampi(rx, 0, W - 1)];
            }
            int Nn = N + 1 < r->cap ? N + 1 : r->cap;
            float invN = 1.0f / (float)Nn;
            double a_c = invN > r->alpha_c ? invN : r->alpha_c;
            double a_m = invN > r->alpha_m ? invN : r->alpha_m;
            double Cn[3], Mn[2];
            for (int k = 0; k < 3; ++k) Cn[k] = Cp[k] + a_c * ((double)il[k] - Cp[k]);
            for (int k = 0; k < 2; ++k) Mn[k] = Mp[k] + a_m * (mu[k] - Mp[k]);
            double var = Mn[1] - Mn[0] * Mn[0];
            if (var < 0) var = 0;
            oc[0] = (float)Cn[0]; oc[1] = (float)Cn[1]; oc[2] = (float)Cn[2];
            oc[3] = (float)lum(Cn[0], Cn[1], Cn[2]);
            s->t_v_pre[p] = (float)var;
            mom[2 * p] = (float)Mn[0]; mom[2 * p + 1] = (float)Mn[1];
            hl[p] = (uint8_t)Nn;
        }
}

/* ---- pass 2: 7x7 spatial variance for short histories (spec S3) -------------------- */
static void pass_variance(oracle_svgf* s, const resolved_params* r) {
    const int W = s->W, H = s->H;
    const float* g4 = s->g4[s->parity];
    const float* mom = s->mom[s->parity];
    const uint8_t* hl = s->hl[s->parity];
    size_t n = (size_t)W * H;
    memcpy(s->t_c4, s->t_c4_pre, n * 16);
    memcpy(s->t_v, s->t_v_pre, n * 4);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            const float* gp = g4 + 4 * p;
            if (gp[3] == 0.0f) continue;
            int Nn = hl[p];
            if (Nn >= r->short_hist) continue;
            const float* cp = s->t_c4_pre + 4 * p;
            double sw = 1.0, sc[3] = {cp[0], cp[1], cp[2]}, sm[2] = {mom[2 * p], mom[2 * p + 1]};
            double zs = (double)r->sigma_z * fmax((double)s->dz[p], 1e-8);
            for (int dx = -3; dx <= 3; ++dx)
                for (int dy = -3; dy <= 3; ++dy) {
                    if (!dx && !dy) continue;
                    int qx = x + dx, qy = y + dy;
                    if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
                    size_t q = (size_t)qy * W + qx;
                    const float* gq = g4 + 4 * q;
                    if (gq[3] == 0.0f) continue;
                    double d = (double)gp[0] * gq[0] + (double)gp[1] * gq[1] + (double)gp[2] * gq[2];
                    if (d <= 0) continue;
                    double wn = pow(d, (double)r->sigma_n);
                    double tz = fabs((double)gp[3] - gq[3]) / (zs * sqrt((double)(dx * dx + dy * dy)) + 1e-6);
                    double tl = fabs((double)cp[3] - s->t_c4_pre[4 * q + 3]) / (double)r->lscale;
                    double w = wn * exp(-tz - tl);
                    sw += w;
                    for (int k = 0; k < 3; ++k) sc[k] += w * s->t_c4_pre[4 * q + k];
                    sm[0] += w * mom[2 * q]; sm[1] += w * mom[2 * q + 1];
                }
            if (sw < 1e-6) sw = 1e-6;
            double c[3] = {sc[0] / sw, sc[1] / sw, sc[2] / sw};
            double m0 = sm[0] / sw, m1 = sm[1] / sw;
            double var = m1 - m0 * m0;
            if (var < 0) var = 0;
            var *= 4.0 / (double)Nn;
            float* oc = s->t_c4 + 4 * p;
            oc[0] = (float)c[0]; oc[1] = (float)c[1]; oc[2] = (float)c[2];
            oc[3] = (float)lum(c[0], c[1], c[2]);
            s->t_v[p] = (float)var;
        }
}

/* ---- pass 3..: one a-trous level, step = 1 << level (spec S4) ----------------------- */
static void pass_atrous(const oracle_svgf* s, const resolved_params* r, int step, const float* in_c4,
                        const float* in_v, float* out_c4, float* out_v) {
    const int W = s->W, H = s->H;
    const float* g4 = s->g4[s->parity];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            const float* gp = g4 + 4 * p;
            const float* cp = in_c4 + 4 * p;
            if (gp[3] == 0.0f) { /* sky passes through every level */
                memcpy(out_c4 + 4 * p, cp, 16);
                out_v[p] = in_v[p];
                continue;
            }
            /* 3x3 Gaussian prefilter of the variance, edge taps clamped */
            double vbar = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    double g = (dx == 0 ? 0.5 : 0.25) * (dy == 0 ? 0.5 : 0.25);
                    vbar += g * in_v[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
                }
            double phi_l = (double)r->sigma_l * sqrt(fmax(0.0, vbar)) + 1e-4;
            double zs = (double)r->sigma_z * fmax((double)s->dz[p], 1e-8) * (double)step;
            double h0 = kSpline[0] * kSpline[0];
            double sw = h0, sc[3] = {h0 * cp[0], h0 * cp[1], h0 * cp[2]}, sv = h0 * h0 * in_v[p];
            for (int dx = -2; dx <= 2; ++dx)     /* x outer, reference src/filter.cu:34 */
                for (int dy = -2; dy <= 2; ++dy) { /* y inner, reference src/filter.cu:35 */
                    if (!dx && !dy) continue;
                    int qx = x + step * dx, qy = y + step * dy;
                    if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue; /* skip + renormalise, src/filter.cu:38-39 */
                    size_t q = (size_t)qy * W + qx;
                    const float* gq = g4 + 4 * q;
                    if (gq[3] == 0.0f) continue; /* sky taps carry no weight */
                    double d = (double)gp[0] * gq[0] + (double)gp[1] * gq[1] + (double)gp[2] * gq[2];
                    if (d <= 0) continue;
                    double wn = pow(d, (double)r->sigma_n);
                    double tz = fabs((double)gp[3] - gq[3]) / (zs * sqrt((double)(dx * dx + dy * dy)) + 1e-6);
                    double tl = fabs((double)cp[3] - in_c4[4 * q + 3]) / phi_l;
                    double w = wn * exp(-tz - tl);
                    double hw = kSpline[abs(dx)] * kSpline[abs(dy)] * w;
                    sw += hw;
                    for (int k = 0; k < 3; ++k) sc[k] += hw * in_c4[4 * q + k];
                    sv += hw * hw * in_v[q];
                }
            double c[3] = {sc[0] / sw, sc[1] / sw, sc[2] / sw};
            float* oc = out_c4 + 4 * p;
            oc[0] = (float)c[0]; oc[1] = (float)c[1]; oc[2] = (float)c[2];
            oc[3] = (float)lum(c[0], c[1], c[2]);
            out_v[p] = (float)(sv / (sw * sw));
        }
}

int oracle_svgf_frame(oracle_svgf* s, const RmdSvgfFrame* f, const RmdFilterParams* fp, const RmdSvgfParams* sp) {
    if (!s || !f) return RMD_E_NULL;
    if (f->width != s->W || f->height != s->H) return RMD_E_SHAPE;
    if (!f->color || !f->albedo || !f->guide || !f->motion || !f->out) return RMD_E_NULL;
    resolved_params r;
    int rc = resolve(fp, sp, &r);
    if (rc) return rc;
    const int W = s->W, H = s->H;
    const size_t n = (size_t)W * H;
    s->parity ^= 1;
    pass_guide(s, (const uint32_t*)f->guide, s->g4[s->parity]);
    pass_temporal(s, f, &r);
    pass_variance(s, &r);
    const float* in_c4 = s->t_c4;
    const float* in_v = s->t_v;
    for (int l = 0; l < r.depth; ++l) {
        /* level 0 writes the colour history of the next frame (spec S4) */
        float* oc = l == 0 ? s->hist_c4 : s->lv_c4[l & 1];
        float* ov = s->lv_v[l & 1];
        pass_atrous(s, &r, 1 << l, in_c4, in_v, oc, ov);
        in_c4 = oc; in_v = ov;
    }
    if (r.depth == 0) memcpy(s->hist_c4, s->t_c4, n * 16);
    s->have_history = 1;
    /* re-modulate (spec S5) */
    const uint8_t* albedo = (const uint8_t*)f->albedo;
    const float* g4 = s->g4[s->parity];
    float* out = (float*)f->out;
    uint8_t* out8 = (uint8_t*)f->out_rgba8;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; ++p) {
        int sky = g4[4 * p + 3] == 0.0f;
        for (int k = 0; k < 3; ++k) {
            float a = (float)albedo[4 * p + k] * (1.0f / 255.0f);
            if (a < r.afloor) a = r.afloor;
            out[4 * p + k] = sky ? in_c4[4 * p + k] : (float)((double)in_c4[4 * p + k] * (double)a);
        }
        out[4 * p + 3] = in_v[p];
        if (out8) {
            for (int k = 0; k < 3; ++k) {
                float v = out[4 * p + k];
                v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
                out8[4 * p + k] = (uint8_t)(v * 255.0f);
            }
            out8[4 * p + 3] = 255;
        }
    }
    return 0;
}

const void* oracle_svgf_plane(const oracle_svgf* s, int plane) {
    switch (plane) {
        case RMD_PLANE_TEMPORAL_COLOR: return s->t_c4;
        case RMD_PLANE_TEMPORAL_VAR: return s->t_v;
        case RMD_PLANE_MOMENTS: return s->mom[s->parity];
        case RMD_PLANE_HISTLEN: return s->hl[s->parity];
        case RMD_PLANE_HISTORY_COLOR: return s->hist_c4;
        case RMD_PLANE_GUIDE: return s->g4[s->parity];
        case RMD_PLANE_SLOPE: return s->dz;
        case ORACLE_PLANE_TEMPORAL_COLOR_PRE: return s->t_c4_pre;
        case ORACLE_PLANE_TEMPORAL_VAR_PRE: return s->t_v_pre;
        default: return NULL;
    }
}

/* OpenMP team size of the oracle (bench.py's CPU arm: torchrun exports OMP_NUM_THREADS=1 to its children, which
 * silently made the round-1 N>1 CPU arm single-threaded; the bench now sets and reports the count explicitly). */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_num_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_num_threads(void) { return 1; }
#endif
