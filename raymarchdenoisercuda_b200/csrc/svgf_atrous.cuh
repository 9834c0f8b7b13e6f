// svgf_atrous.cuh — device code shared by the a-trous level kernels (independent-tile kernel,
// svgf_atrous_tile.cu; persistent ring kernel, svgf_atrous.cu): the per-tap arithmetic, the
// per-output centre terms and the epilogue.  DESIGN.md spec S4-S5, checked against
// oracle/oracle_svgf.c:pass_atrous.
//
// Reference hooks: the taps are the reference's `waveletSpline = {3/8, 1/4, 1/16}`
// (src/filter.cu:10); the border rule is its "skip the tap and renormalise"
// (src/filter.cu:38-39, 46, 49).
#pragma once
#include "svgf.cuh"

namespace rmd {
namespace {

constexpr int align128(int v) { return (v + 127) & ~127; }

// lg2 of the B3-spline taps {3/8, 1/4, 1/16} (reference src/filter.cu:10)
__device__ __forceinline__ constexpr float lg2_spline(int a) {
    return a == 0 ? -1.4150374992788437f : (a == 1 ? -2.0f : -4.0f);
}
// distance class of a tap: |d|^2 in {1,2,4,5,8} -> 0..4
__device__ __forceinline__ constexpr int dist_class(int adx, int ady) {
    const int d2 = adx * adx + ady * ady;
    return d2 == 1 ? 0 : d2 == 2 ? 1 : d2 == 4 ? 2 : d2 == 5 ? 3 : 4;
}

template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ float4 lds128_dyn(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32_dyn(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// 3x3 Gaussian {1/4,1/8,1/16} prefilter of the variance; fixed operation order so that every
// kernel variant agrees bit for bit
__device__ __forceinline__ float vbar3x3(float tm, float tc, float tp, float mm, float mc, float mp, float bm, float bc,
                                         float bp) {
    const float top = __fadd_rn(fmaf(2.0f, tc, tm), tp);
    const float mid = __fadd_rn(fmaf(2.0f, mc, mm), mp);
    const float bot = __fadd_rn(fmaf(2.0f, bc, bm), bp);
    return __fmul_rn(__fadd_rn(fmaf(2.0f, mid, top), bot), 1.0f / 16.0f);
}

struct Centre {
    float nx, ny, nz, z, L;
    float il;     // log2(e) / phi_l
    float iz[5];  // log2(e) / (phi_z * |d| + 1e-6) per distance class
};
struct Acc {
    float r, g, b, w, v;
};

template <int ADX, int ADY>
__device__ __forceinline__ void tap(Acc& acc, const Centre& c, const float4 q, const float4 g, const float v,
                                    const float sigma_n) {
    // max(0, n.n') as the saturate modifier of the last FMA (unit normals: the upper clamp at 1 only trims rounding)
    const float d = __saturatef(fmaf(c.nz, g.z, fmaf(c.ny, g.y, c.nx * g.x)));
    float e = fmaf(fast_lg2(d), sigma_n, lg2_spline(ADX) + lg2_spline(ADY));
    e = fmaf(fabsf(c.z - g.w), -c.iz[dist_class(ADX, ADY)], e);
    e = fmaf(fabsf(c.L - q.w), -c.il, e);
    const float hw = fast_ex2(e);
    acc.w += hw;
    acc.r = fmaf(hw, q.x, acc.r);
    acc.g = fmaf(hw, q.y, acc.g);
    acc.b = fmaf(hw, q.z, acc.b);
    acc.v = fmaf(hw * hw, v, acc.v);
}

// A tap between the centres of two outputs a, b of the same thread occurs twice (a reads b's texel, b reads a's) with
// the same normal dot product, |dz| and |dL|: IEEE multiplication and |x - y| are symmetric, so evaluating them once
// gives both taps the bits of tap<>().  hw_a / hw_b = weight of a's tap on b's texel / of b's tap on a's texel.
template <int ADY>
__device__ __forceinline__ void pair_weights(const Centre& ca, const Centre& cb, const float sigma_n, float& hw_a, float& hw_b) {
    const float d = __saturatef(fmaf(ca.nz, cb.nz, fmaf(ca.ny, cb.ny, ca.nx * cb.nx)));
    const float e0 = fmaf(fast_lg2(d), sigma_n, lg2_spline(0) + lg2_spline(ADY));
    const float az = fabsf(ca.z - cb.z), al = fabsf(ca.L - cb.L);
    hw_a = fast_ex2(fmaf(al, -ca.il, fmaf(az, -ca.iz[dist_class(0, ADY)], e0)));
    hw_b = fast_ex2(fmaf(al, -cb.il, fmaf(az, -cb.iz[dist_class(0, ADY)], e0)));
}
// the accumulation half of tap<>()
__device__ __forceinline__ void tap_accumulate(Acc& acc, const float hw, const float4 q, const float v) {
    acc.w += hw;
    acc.r = fmaf(hw, q.x, acc.r);
    acc.g = fmaf(hw, q.y, acc.g);
    acc.b = fmaf(hw, q.z, acc.b);
    acc.v = fmaf(hw * hw, v, acc.v);
}

template <int S>
__device__ __forceinline__ void centre_setup(Centre& ctr, Acc& acc, const float4 c, const float4 g, const float v,
                                             const float vbar, const float dz, const AtrousArgs& a) {
    // il = log2(e) / phi_l and iz_k = log2(e) / (phi_z * d_k + 1e-6), with 1/log2(e) folded into the
    // operands so that each is one FFMA + one MUFU.RCP
    const float kLn2 = 0.6931471805599453f;
    ctr.nx = g.x; ctr.ny = g.y; ctr.nz = g.z; ctr.z = g.w; ctr.L = c.w;
    ctr.il = fast_rcp(fmaf(a.sigma_l * kLn2, fast_sqrt(fmaxf(vbar, 0.0f)), 1e-4f * kLn2));
    const float zs = a.sigma_z * fmaxf(dz, 1e-8f) * ((float)S * kLn2);
    ctr.iz[0] = fast_rcp(fmaf(zs, 1.0f, 1e-6f * kLn2));
    ctr.iz[1] = fast_rcp(fmaf(zs, 1.4142135623730951f, 1e-6f * kLn2));
    ctr.iz[2] = fast_rcp(fmaf(zs, 2.0f, 1e-6f * kLn2));
    ctr.iz[3] = fast_rcp(fmaf(zs, 2.23606797749979f, 1e-6f * kLn2));
    ctr.iz[4] = fast_rcp(fmaf(zs, 2.8284271247461903f, 1e-6f * kLn2));
    const float h0 = 0.140625f;  // (3/8)^2
    acc.w = h0;
    acc.r = h0 * c.x; acc.g = h0 * c.y; acc.b = h0 * c.z;
    acc.v = h0 * h0 * v;
}

// taps of one staged texel in a column with |dx| = ADX (centre column when ADX == 0), tile row JR,
// for the kAtrousOPT outputs of a thread (tile rows 2 .. 2 + kAtrousOPT - 1)
template <int ADX, int JR>
__device__ __forceinline__ void taps_of_texel_adx(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const float4 q,
                                                  const float4 g, const float v, const float sigma_n) {
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int dy = JR - 2 - j;
        if (dy < -2 || dy > 2) continue;
        if (dy == 0 && ADX == 0) continue;  // centre tap, already accumulated with w = 1
        const int ady = dy < 0 ? -dy : dy;
        if (ady == 0) tap<ADX, 0>(acc[j], ctr[j], q, g, v, sigma_n);
        else if (ady == 1) tap<ADX, 1>(acc[j], ctr[j], q, g, v, sigma_n);
        else tap<ADX, 2>(acc[j], ctr[j], q, g, v, sigma_n);
    }
}

// Stores of one finished output (planes for the next level and/or the caller's output with the albedo
// re-modulated, spec S5).
__device__ __forceinline__ void write_output(const AtrousArgs& a, float r, float g, float b, float lum, float v, bool sky,
                                             int x, int y) {
    if (a.out_c4) {
        const size_t p = (size_t)y * a.Wp + x;
        a.out_c4[p] = make_float4(r, g, b, lum);
        a.out_v[p] = v;
    }
    if (a.final_out) {
        const size_t p = (size_t)y * a.W + x;  // caller planes: pitch W
        if (!sky) {
            const uchar4 al = __ldg(a.albedo + p);
            r *= fmaxf(__fmul_rn((float)al.x, 1.0f / 255.0f), a.afloor);
            g *= fmaxf(__fmul_rn((float)al.y, 1.0f / 255.0f), a.afloor);
            b *= fmaxf(__fmul_rn((float)al.z, 1.0f / 255.0f), a.afloor);
        }
        st_cs_f4(a.final_out + p, make_float4(r, g, b, v));
        if (a.final_rgba8) {
            a.final_rgba8[p] = make_uchar4((unsigned char)(__saturatef(r) * 255.0f),
                                           (unsigned char)(__saturatef(g) * 255.0f),
                                           (unsigned char)(__saturatef(b) * 255.0f), 255);
        }
    }
}

// Epilogue of one output.  `c_addr` / `v_addr` are the shared-window addresses of the centre texel's
// colour and variance: a sky pixel passes its input through, and re-reading it here (rare) is cheaper
// than keeping 5 registers per output alive through the tap loop.
__device__ __forceinline__ void store_output(const AtrousArgs& a, const Acc& acc, const Centre& ctr, uint32_t c_addr,
                                             uint32_t v_addr, int x, int y) {
    const float inv = fast_rcp(acc.w);
    float r = acc.r * inv, g = acc.g * inv, b = acc.b * inv, v = acc.v * inv * inv;
    const bool sky = ctr.z == 0.0f;
    float lum;
    if (sky) {
        const float4 cC = lds128_dyn(c_addr);
        r = cC.x; g = cC.y; b = cC.z; lum = cC.w;
        v = lds32_dyn(v_addr);
    } else {
        lum = luminance(r, g, b);
    }
    write_output(a, r, g, b, lum, v, sky, x, y);
}
// persistent tile kernel: a sky output has already replaced its accumulators by the pass-through values
// (acc.w = 1, ctr.L = input luminance), so the epilogue needs registers only
__device__ __forceinline__ void store_output_regs(const AtrousArgs& a, const Acc& acc, const Centre& ctr, int x, int y) {
    const float inv = fast_rcp(acc.w);
    const float r = acc.r * inv, g = acc.g * inv, b = acc.b * inv, v = acc.v * inv * inv;
    const bool sky = ctr.z == 0.0f;
    write_output(a, r, g, b, sky ? ctr.L : luminance(r, g, b), v, sky, x, y);
}
// same with the centre texel already in registers (ring kernel: its slot may have been refilled)
__device__ __forceinline__ void store_output_vals(const AtrousArgs& a, const Acc& acc, const Centre& ctr, const float4 cC,
                                                  const float cV, int x, int y) {
    const float inv = fast_rcp(acc.w);
    float r = acc.r * inv, g = acc.g * inv, b = acc.b * inv, v = acc.v * inv * inv;
    const bool sky = ctr.z == 0.0f;
    if (sky) { r = cC.x; g = cC.y; b = cC.z; v = cV; }
    write_output(a, r, g, b, sky ? cC.w : luminance(r, g, b), v, sky, x, y);
}

}  // namespace
}  // namespace rmd
