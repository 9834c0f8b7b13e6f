/*
 * synth_scene.c — deterministic synthetic 1-spp G-buffer sequences (host only).
 *
 * Workload definition for BASELINE.json configs[1..4] (SURVEY.md §8d "Synthetic
 * generator"): the reference ships one 500x500 still with a saturated depth
 * plane and no motion vectors, so every temporal
 * configuration needs generated input.  This file is input data, not the
 * algorithm: it is used by tests/ and bench.py to feed BOTH the CUDA path and
 * the CPU oracle with the same bits.
 *
 * Scene: a background plane (z = 10, 8x8 checker albedo, a horizontal sky band)
 * plus K moving rectangles at depths 2..8 with per-rectangle unit normal,
 * albedo, irradiance and screen velocity; the whole scene pans by (+3,+1)
 * px/frame.  All positions and velocities are integers in 1/16 px, depths are
 * integers in 2^-10, so floor(), bilinear fractions and every threshold the
 * denoiser evaluates are exact in fp32 (no CPU/GPU decision flips, SURVEY §7
 * "Decision flips").  Noise is white per (pixel, frame): radiance =
 * albedo * E * g, g = 5 u^4 (mean 1, heavy tail), with 1/1024 of the pixels
 * multiplied by 8 (fireflies).
 *
 * Output planes are written in the storage formats of include/rmd_b200.h
 * (RmdSvgfFrame): RGBA16F colour, RGBA8 albedo, {oct-snorm16 normal, fp32 z}
 * guide, RG16F motion.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SYNTH_MAX_LAYERS 64

static inline uint32_t pcg_hash(uint32_t v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
static inline uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
    return pcg_hash(a ^ pcg_hash(b ^ pcg_hash(c)));
}
static inline uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return pcg_hash(a ^ pcg_hash(b ^ pcg_hash(c ^ pcg_hash(d))));
}

static inline uint16_t f32_to_f16(float f) {
    _Float16 h = (_Float16)f; /* IEEE round-to-nearest-even */
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}

typedef struct {
    int32_t w16, h16;   /* size in 1/16 px                         */
    int32_t px0, py0;   /* position at frame 0, 1/16 px            */
    int32_t vx, vy;     /* velocity incl. camera pan, 1/16 px/frame */
    int32_t zb;         /* depth at local x = 0, units of 2^-10    */
    int32_t slope;      /* depth increase per pixel of local x, 2^-10 */
    int16_t nx, ny;     /* octahedral snorm16 normal               */
    uint8_t alb[3];
    float irradiance;
} layer_t;

static inline int32_t posmod(int64_t a, int32_t m) {
    int64_t r = a % m;
    return (int32_t)(r < 0 ? r + m : r);
}

static void build_layers(int W, int H, uint32_t seed, int K, layer_t* L) {
    const int panx = 48, pany = 16; /* +3, +1 px per frame */
    for (int k = 0; k < K; ++k) {
        layer_t* l = &L[k];
        uint32_t b = seed * 0x9E3779B1u + (uint32_t)k * 0x85EBCA77u;
        int wmin = W / 16 > 4 ? W / 16 : 4, wmax = W / 4 > wmin ? W / 4 : wmin + 1;
        int hmin = H / 16 > 4 ? H / 16 : 4, hmax = H / 4 > hmin ? H / 4 : hmin + 1;
        l->w16 = 16 * (wmin + (int)(hash3(b, 1, 0) % (uint32_t)(wmax - wmin)));
        l->h16 = 16 * (hmin + (int)(hash3(b, 2, 0) % (uint32_t)(hmax - hmin)));
        l->px0 = (int32_t)(hash3(b, 3, 0) % (uint32_t)(16 * W));
        l->py0 = (int32_t)(hash3(b, 4, 0) % (uint32_t)(16 * H));
        l->vx = (int32_t)(hash3(b, 5, 0) % 257u) - 128 + panx;
        l->vy = (int32_t)(hash3(b, 6, 0) % 257u) - 128 + pany;
        l->zb = 2048 + (int32_t)(hash3(b, 7, 0) % 6145u); /* z in [2, 8] */
        l->slope = (int32_t)(hash3(b, 8, 0) % 5u);        /* 0 .. 4 * 2^-10 per px */
        /* unit normal on the z > 0 hemisphere: |fx| + |fy| <= 0.9 in octahedral space */
        int32_t ax = (int32_t)(hash3(b, 9, 0) % 29491u);
        int32_t ay = (int32_t)(hash3(b, 10, 0) % (uint32_t)(29491 - ax));
        l->nx = (int16_t)((hash3(b, 11, 0) & 1u) ? -ax : ax);
        l->ny = (int16_t)((hash3(b, 12, 0) & 1u) ? -ay : ay);
        for (int c = 0; c < 3; ++c) l->alb[c] = (uint8_t)(51 + hash3(b, 13, (uint32_t)c) % 180u);
        l->irradiance = (float)(51 + hash3(b, 14, 0) % 461u) / 256.0f; /* [0.2, 2.0] */
    }
}

/* Generates frame `frame` of the sequence identified by (W, H, seed, K). */
void rmd_synth_frame(int W, int H, uint32_t seed, int K, int frame, void* color_rgba16f, void* albedo_rgba8,
                     void* guide_u32x2, void* motion_rg16f) {
    if (K > SYNTH_MAX_LAYERS) K = SYNTH_MAX_LAYERS;
    if (K < 0) K = 0;
    layer_t L[SYNTH_MAX_LAYERS];
    build_layers(W, H, seed, K, L);
    const int panx = 48, pany = 16;
    const int W16 = 16 * W, H16 = 16 * H;
    int32_t lpx[SYNTH_MAX_LAYERS], lpy[SYNTH_MAX_LAYERS];
    for (int k = 0; k < K; ++k) {
        lpx[k] = posmod((int64_t)L[k].px0 + (int64_t)frame * L[k].vx, W16);
        lpy[k] = posmod((int64_t)L[k].py0 + (int64_t)frame * L[k].vy, H16);
    }
    const int32_t bgx = posmod((int64_t)frame * panx, W16), bgy = posmod((int64_t)frame * pany, H16);
    uint16_t* color = (uint16_t*)color_rgba16f;
    uint8_t* albedo = (uint8_t*)albedo_rgba8;
    uint32_t* guide = (uint32_t*)guide_u32x2;
    uint16_t* motion = (uint16_t*)motion_rg16f;
    const int sky_lo = H / 8, sky_hi = H / 8 + H / 16;

#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            /* background */
            int32_t blx = posmod((int64_t)16 * x - bgx, W16), bly = posmod((int64_t)16 * y - bgy, H16);
            int32_t best_z = 10240 + ((2 * blx) >> 4);
            int best = -1;
            int32_t best_lx = blx;
            for (int k = 0; k < K; ++k) {
                int32_t lx = posmod((int64_t)16 * x - lpx[k], W16), ly = posmod((int64_t)16 * y - lpy[k], H16);
                if (lx < L[k].w16 && ly < L[k].h16) {
                    int32_t z = L[k].zb + ((L[k].slope * lx) >> 4);
                    if (z < best_z) { best_z = z; best = k; best_lx = lx; }
                }
            }
            uint8_t a[3];
            int16_t nx = 0, ny = 0;
            int32_t vx = panx, vy = pany;
            float E = 1.0f;
            int sky = 0;
            if (best < 0) {
                int cy = bly >> 4, cx = blx >> 4;
                if (cy >= sky_lo && cy < sky_hi) sky = 1;
                uint8_t g = (((cx >> 3) + (cy >> 3)) & 1) ? 192 : 96;
                a[0] = a[1] = a[2] = g;
            } else {
                const layer_t* l = &L[best];
                int stripe = ((best_lx >> 4) >> 4) & 1; /* 16-px albedo stripes */
                for (int c = 0; c < 3; ++c) a[c] = stripe ? (uint8_t)((l->alb[c] * 3) >> 2) : l->alb[c];
                nx = l->nx; ny = l->ny; vx = l->vx; vy = l->vy; E = l->irradiance;
            }
            uint32_t h = hash4(seed, (uint32_t)x, (uint32_t)y, (uint32_t)frame);
            float u = (float)(h & 0xFFFFu) * (1.0f / 65536.0f);
            float u2 = u * u;
            float g = 5.0f * (u2 * u2);
            if (((h >> 16) & 1023u) == 0u) g *= 8.0f;
            float rad[3];
            if (sky) {
                rad[0] = 0.5f; rad[1] = 0.75f; rad[2] = 1.5f; /* noise-free sky radiance */
                a[0] = a[1] = a[2] = 255;
            } else {
                for (int c = 0; c < 3; ++c) rad[c] = ((float)a[c] * (1.0f / 255.0f)) * (E * g);
            }
            color[4 * p + 0] = f32_to_f16(rad[0]);
            color[4 * p + 1] = f32_to_f16(rad[1]);
            color[4 * p + 2] = f32_to_f16(rad[2]);
            color[4 * p + 3] = f32_to_f16(1.0f);
            albedo[4 * p + 0] = a[0]; albedo[4 * p + 1] = a[1]; albedo[4 * p + 2] = a[2]; albedo[4 * p + 3] = 255;
            float zf = sky ? 0.0f : (float)best_z * (1.0f / 1024.0f);
            uint32_t zbits;
            memcpy(&zbits, &zf, 4);
            guide[2 * p + 0] = (uint32_t)(uint16_t)nx | ((uint32_t)(uint16_t)ny << 16);
            guide[2 * p + 1] = zbits;
            /* the visible surface point was at p - v one frame ago: prev = p + motion */
            motion[2 * p + 0] = f32_to_f16((float)(-vx) * (1.0f / 16.0f));
            motion[2 * p + 1] = f32_to_f16((float)(-vy) * (1.0f / 16.0f));
        }
    }
}

/* A still frame with zero motion (config 1 style: no history), same scene. */
int rmd_synth_num_layers_default(void) { return 24; }
