// svgf_temporal.cu — pass 1 of the SVGF path: guide decode, depth slope, albedo
// demodulation, motion-vector reprojection with depth/normal disocclusion tests,
// colour + luminance-moment accumulation and history-length tracking.
//
// No reference counterpart (the reference has no temporal code);
// arithmetic follows DESIGN.md "SVGF specification" S1-S2 and is checked against
// oracle/oracle_svgf.c:pass_guide/pass_temporal.
//
// Roofline: HBM.  Algorithmic bytes per pixel: reads colour 8 + albedo 4 + guide 8
// + motion 4 + previous guide 16 + colour history 16 + moment history 8 + history
// length 1 = 65; writes colour 16 + variance 4 + decoded guide 16 + slope 4 +
// moments 8 + history length 1 = 49; total 114 B/px.  The four bilinear taps of a
// warp overlap (smooth motion) and are served by L1/L2, so they are counted once.
#include "svgf.cuh"

namespace rmd {

namespace {

__device__ __forceinline__ float4 half4_to_float4(uint2 h) {
    const __half2 a = *reinterpret_cast<const __half2*>(&h.x);
    const __half2 b = *reinterpret_cast<const __half2*>(&h.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// reprojection tap validity (spec S2): inside the image, depth within tolerance,
// normals agree.  Un-fused fp32, same order as oracle tap_valid().
__device__ __forceinline__ bool tap_test(bool inside, float4 gq, float4 gp, float rhs, float nthr) {
    const float lhs = fabsf(__fsub_rn(gq.w, gp.w));
    return inside && (lhs <= rhs) && (dot3_rn(gq, gp) >= nthr);
}
__device__ __forceinline__ bool tap_valid(const float4* __restrict__ prev_g4, int W, int ylo, int yhi, int Wp, int tx, int ty,
                                          float4 gp, float rhs, float nthr) {
    if (tx < 0 || ty < ylo || tx >= W || ty >= yhi) return false;
    return tap_test(true, __ldg(prev_g4 + (size_t)ty * Wp + tx), gp, rhs, nthr);
}

// Fallback of the reprojection (spec S2): 3x3 search around round(q), unweighted means of the colour and moment
// history over the valid taps.  Rare (disocclusion borders) but divergent: kept out of line so that its registers and
// its nine dependent gathers stay off the common path, and done in ONE walk (round 1 walked the window twice: once for
// the colour, once more for the moments).
struct Search3x3 {
    float r, g, b;
    double m0, m1;
    int cnt;
};
__device__ __noinline__ Search3x3 search_3x3(const float4* __restrict__ prev_g4, const float4* __restrict__ hist_c4,
                                             const float2* __restrict__ hist_m, int W, int ylo, int yhi, int Wp, int rx, int ry,
                                             float4 gp, float rhs, float nthr) {
    Search3x3 s{0.f, 0.f, 0.f, 0.0, 0.0, 0};
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
            if (tap_valid(prev_g4, W, ylo, yhi, Wp, rx + dx, ry + dy, gp, rhs, nthr)) {
                const size_t q = (size_t)(ry + dy) * Wp + (rx + dx);
                const float4 c3 = __ldg(hist_c4 + q);
                const float2 m3 = __ldg(hist_m + q);
                s.r += c3.x; s.g += c3.y; s.b += c3.z;
                s.m0 += (double)m3.x; s.m1 += (double)m3.y;
                ++s.cnt;
            }
    return s;
}

// Moment update M' = M + a (mu - M), Var = max(0, M'_2 - M'_1^2) in precision T.
template <typename T>
__device__ __forceinline__ void moments_finish(T M0, T M1, T mu0, T mu1, float am, float2& m_out, float& var_out) {
    M0 = M0 + (T)am * (mu0 - M0);
    M1 = M1 + (T)am * (mu1 - M1);
    const T v = M1 - M0 * M0;
    var_out = (float)(v > (T)0 ? v : (T)0);
    m_out = make_float2((float)M0, (float)M1);
}

// fp32 evaluates M_2 - M_1^2 with an absolute error of ~1e-7 L^2: harmless for ordinary luminance, catastrophic
// once the demodulated luminance is large (albedo at the floor: L ~ 1e3).  Pixels whose current or history
// luminance exceeds this take the FP64 path (same operation order as the oracle); they are rare.
constexpr float kBrightLuminance = 8.0f;

#ifndef RMD_TEMPORAL_MINB
#define RMD_TEMPORAL_MINB 4
#endif
// BAND = the context is one row band of a larger frame: rows outside [full_begin, full_end) only get their guide
// decoded, and history rows outside [hist_row_lo, hist_row_hi) count as outside the image.  The single-frame
// instantiation folds both away (they cost 4 us per 1080p frame when left as run-time tests).
template <bool BAND, int MINB>
__global__ void __launch_bounds__(kTemporalBx* kTemporalBy, MINB) temporal_kernel(const TemporalArgs a) {
    // PDL: the launch itself overlaps the tail of the previous kernel in the stream; nothing is read before that
    // kernel (possibly the caller's producer of the input planes) has completed
    pdl_wait();
    pdl_launch_dependents();
    const int x = blockIdx.x * kTemporalBx + threadIdx.x;
    const int y = a.row_begin + blockIdx.y * kTemporalBy + threadIdx.y;
    const int W = a.W, H = a.H, Wp = a.Wp;
    // rows of the history planes that hold valid history: the whole plane, or (one band of a frame) the own rows plus
    // the rows the neighbours refreshed; a reprojection tap beyond them counts as outside the image (disoccluded)
    const int ylo = BAND ? a.hist_row_lo : 0, yhi = BAND ? a.hist_row_hi : H;
    // L2 prefetch for a CTA that will run a few waves from now: its inputs, and the history texels at its pixels' OWN
    // position (motion is a few pixels: the lines its reprojection gathers will touch are these or their neighbours).
    // Worth 1.5-2 % of the pass (profiles/r2_notes.md): the two dependent round trips per pixel (inputs, then the
    // gathers they address) become L2 hits, but the pass is limited by DRAM efficiency over its 15 concurrent streams
    // (8 read, 7 written), not by latency.
    if (a.prefetch_ctas > 0) {
        const int gx = gridDim.x;
        const long long b = (long long)blockIdx.y * gx + blockIdx.x + a.prefetch_ctas;
        const int py = a.row_begin + (int)(b / gx) * kTemporalBy + threadIdx.y, px = (int)(b % gx) * kTemporalBx + threadIdx.x;
        if (px < W && py < a.row_end) {
            const size_t qi = (size_t)py * W + px, qo = (size_t)py * Wp + px;
            prefetch_l2(a.color + qi); prefetch_l2(a.guide + qi); prefetch_l2(a.albedo + qi); prefetch_l2(a.motion + qi);
            if (a.have_history) {
                prefetch_l2(a.hist_c4 + qo); prefetch_l2(a.prev_g4 + qo); prefetch_l2(a.hist_m + qo);
                if ((threadIdx.x & 31) == 0) prefetch_l2(a.hist_n + qo);
            }
        }
    }
    bool short_hist = false;
    if (BAND && x < W && y < a.row_end && (y < a.full_begin || y >= a.full_end)) {
        // band mode: halo rows beyond the temporal range only need the decoded guide (the a-trous levels read it up
        // to 33 rows beyond the rows they produce)
        a.out_g4[(size_t)y * Wp + x] = decode_guide(__ldg(a.guide + (size_t)y * W + x));
    } else if (x < W && y < a.row_end) {
        const size_t pi = (size_t)y * W + x;    // caller planes: pitch W
        const size_t po = (size_t)y * Wp + x;   // context planes: pitch Wp
        // ---- round trip 1: everything that depends only on (x, y) --------------------------------
        const int xr = min(x + 1, W - 1), yd = min(y + 1, H - 1);
        const uint32_t mraw = a.have_history ? __ldg(a.motion + pi) : 0u;
        const uint2 graw = __ldg(a.guide + pi);
        const uint2 graw_x = __ldg(a.guide + (size_t)y * W + xr);
        const uint2 graw_y = __ldg(a.guide + (size_t)yd * W + x);
        const uint2 craw = __ldg(a.color + pi);
        const uint32_t araw = __ldg(a.albedo + pi);
        // ---- round trip 2: the reprojection footprint depends only on the motion vector, so all 13
        //      gathers are issued before the guide is even decoded (clamped coordinates; validity only
        //      decides which of them contribute) -----------------------------------------------------
        const float2 mv = __half22float2(*reinterpret_cast<const __half2*>(&mraw));
        const float qx = __fadd_rn((float)x, mv.x), qy = __fadd_rn((float)y, mv.y);
        const float q0x = floorf(qx), q0y = floorf(qy);
        const int ix = (int)q0x, iy = (int)q0y;
        const int rx = (int)floorf(__fadd_rn(qx, 0.5f)), ry = (int)floorf(__fadd_rn(qy, 0.5f));
        float4 gq[4], hc[4];
        float2 hm[4];
        bool inside[4];
        int Nr = 0;
        if (a.have_history) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int tx = ix + (t & 1), ty = iy + (t >> 1);
                inside[t] = tx >= 0 && ty >= ylo && tx < W && ty < yhi;
                const size_t q = (size_t)min(max(ty, 0), H - 1) * Wp + min(max(tx, 0), W - 1);
                gq[t] = __ldg(a.prev_g4 + q);
                hc[t] = __ldg(a.hist_c4 + q);
                hm[t] = __ldg(a.hist_m + q);
            }
            Nr = a.hist_n[(size_t)min(max(ry, ylo), yhi - 1) * Wp + min(max(rx, 0), W - 1)];
        }
        const float4 gp = decode_guide(graw);
        const float4 c = half4_to_float4(craw);
        if (gp.w == 0.0f) {  // sky: pass through, no history
            a.out_c4[po] = make_float4(c.x, c.y, c.z, luminance(c.x, c.y, c.z));
            a.out_v[po] = 0.0f;
            a.out_m[po] = make_float2(0.f, 0.f);
            a.out_n[po] = 0;
            a.out_g4[po] = gp;
            a.out_dz[po] = 0.0f;
        } else {
            // depth slope: forward differences, clamped at the image edge (spec S1)
            const float4 gx = decode_guide(graw_x);
            const float4 gy = decode_guide(graw_y);
            const float dz = fmaxf(fabsf(__fsub_rn(gx.w, gp.w)), fabsf(__fsub_rn(gy.w, gp.w)));
            // demodulate (spec S2): i = c / max(albedo, floor), IEEE division
            const float ar = fmaxf(__fmul_rn((float)(araw & 255u), 1.0f / 255.0f), a.k.afloor);
            const float ag = fmaxf(__fmul_rn((float)((araw >> 8) & 255u), 1.0f / 255.0f), a.k.afloor);
            const float ab = fmaxf(__fmul_rn((float)((araw >> 16) & 255u), 1.0f / 255.0f), a.k.afloor);
            // approximate reciprocals (<= 2 ulp) on the common path; the FP64 moment path below re-divides exactly
            const float ir = c.x * fast_rcp(ar), ig = c.y * fast_rcp(ag), ib = c.z * fast_rcp(ab);
            const float Lf = luminance(ir, ig, ib);
            float Cr = ir, Cg = ig, Cb = ib;
            // history moments: mode 0 = none (disoccluded), 1 = bilinear taps hm[]/wt[], 2 = 3x3 sums (FP64)
            int mode = 0;
            float wt[4] = {0.f, 0.f, 0.f, 0.f};
            bool ok[4] = {false, false, false, false};
            float sumw = 0.0f;
            float rhs = 0.0f;
            double fb_m0 = 0.0, fb_m1 = 0.0;  // moment means of the 3x3 fallback
            int N = 0;
            if (a.have_history) {
                const float fx = __fsub_rn(qx, q0x), fy = __fsub_rn(qy, q0y);
                rhs = __fadd_rn(__fmul_rn(a.k.dtol, gp.w), __fmul_rn(2.0f, dz));
                const float gx1 = __fsub_rn(1.0f, fx), gy1 = __fsub_rn(1.0f, fy);
                wt[0] = __fmul_rn(gx1, gy1); wt[1] = __fmul_rn(fx, gy1); wt[2] = __fmul_rn(gx1, fy); wt[3] = __fmul_rn(fx, fy);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    ok[t] = tap_test(inside[t], gq[t], gp, rhs, a.k.nthr);
                    if (ok[t]) sumw = __fadd_rn(sumw, wt[t]);
                }
                if (sumw >= 0.01f) {
                    float sr = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (ok[t]) { sr = fmaf(wt[t], hc[t].x, sr); sg = fmaf(wt[t], hc[t].y, sg); sb = fmaf(wt[t], hc[t].z, sb); }
                    const float inv = fast_rcp(sumw);
                    Cr = sr * inv; Cg = sg * inv; Cb = sb * inv;
                    mode = 1;
                } else {
                    // 3x3 search around round(q), unweighted mean of the valid taps (rare path)
                    const Search3x3 fb = search_3x3(a.prev_g4, a.hist_c4, a.hist_m, W, ylo, yhi, Wp, rx, ry, gp, rhs, a.k.nthr);
                    if (fb.cnt > 0) {
                        const float inv = fast_rcp((float)fb.cnt);
                        Cr = fb.r * inv; Cg = fb.g * inv; Cb = fb.b * inv;
                        const double dinv = 1.0 / (double)fb.cnt;
                        fb_m0 = fb.m0 * dinv; fb_m1 = fb.m1 * dinv;
                        mode = 2;
                    }
                }
                if (mode) N = Nr;
            }
            const int Nn = min(N + 1, a.k.cap);
            const float invN = __fdiv_rn(1.0f, (float)Nn);
            const float ac = fmaxf(invN, a.k.alpha_c), am = fmaxf(invN, a.k.alpha_m);
            Cr = fmaf(ac, ir - Cr, Cr); Cg = fmaf(ac, ig - Cg, Cg); Cb = fmaf(ac, ib - Cb, Cb);
            // ---- luminance moments ----
            float2 mom;
            float var;
            float hmax = Lf;
            if (mode == 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (ok[t]) hmax = fmaxf(hmax, hm[t].x);
            }
            if (mode == 2 || hmax > kBrightLuminance) {
                // FP64, same operation order as oracle pass_temporal
                const double Lc = 0.2126f * (double)__fdiv_rn(c.x, ar) + 0.7152f * (double)__fdiv_rn(c.y, ag) +
                                  0.0722f * (double)__fdiv_rn(c.z, ab);
                double M0 = Lc, M1 = Lc * Lc;
                if (mode == 1) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (ok[t]) { s0 += (double)wt[t] * (double)hm[t].x; s1 += (double)wt[t] * (double)hm[t].y; }
                    if (sumw != 1.0f) {  // all four taps valid: the bilinear weights sum to exactly 1
                        const double dinv = 1.0 / (double)sumw;
                        s0 *= dinv; s1 *= dinv;
                    }
                    M0 = s0; M1 = s1;
                } else if (mode == 2) {
                    M0 = fb_m0; M1 = fb_m1;
                }
                moments_finish<double>(M0, M1, Lc, Lc * Lc, am, mom, var);
            } else {
                float M0 = Lf, M1 = Lf * Lf;
                if (mode == 1) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (ok[t]) { s0 = fmaf(wt[t], hm[t].x, s0); s1 = fmaf(wt[t], hm[t].y, s1); }
                    const float inv = fast_rcp(sumw);
                    M0 = s0 * inv; M1 = s1 * inv;
                }
                moments_finish<float>(M0, M1, Lf, Lf * Lf, am, mom, var);
            }
            const float4 oc = make_float4(Cr, Cg, Cb, luminance(Cr, Cg, Cb));
            a.out_c4[po] = oc;
            a.out_v[po] = var;
            a.out_m[po] = mom;
            a.out_n[po] = (uint8_t)Nn;
            a.out_g4[po] = gp;
            a.out_dz[po] = dz;
            short_hist = Nn < a.k.short_hist;
            if (short_hist) a.side_c4[po] = oc;  // untouched copy for the in-place variance pass
        }
    }
    // does the 7x7 variance pass have work in this 32x8 tile?  Then append it to the compact work list the
    // (persistent) variance kernel walks; the order is arbitrary, the results do not depend on it
    const int any = __syncthreads_or(short_hist ? 1 : 0);
    if (any && threadIdx.x == 0 && threadIdx.y == 0) {
        const uint32_t slot = atomicAdd(a.tile_count, 1u);
        if (slot < a.tile_capacity)  // always true unless an earlier frame failed between its two passes
            a.tile_list[slot] = (blockIdx.x << 16) | (uint32_t)(a.row_begin + blockIdx.y * kTemporalBy);
    }
}

}  // namespace

int launch_temporal(const TemporalArgs& a_in, cudaStream_t s, bool pdl) {
    const TemporalArgs& a0 = a_in;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((a0.W + kTemporalBx - 1) / kTemporalBx, (a0.row_end - a0.row_begin + kTemporalBy - 1) / kTemporalBy);
    cfg.blockDim = dim3(kTemporalBx, kTemporalBy);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    static const int prefetch = [] {  // CTAs of look-ahead for the L2 prefetch (RMD_TEMPORAL_PREFETCH=0 switches it off)
        const char* e = getenv("RMD_TEMPORAL_PREFETCH");
        return e ? atoi(e) : 148;  // measured at 4K: 0 -> 231.5 us, 74..592 -> 227-229 us, 1184 -> 235 us, 2368+ -> 262 us
    }();
    TemporalArgs a = a_in;
    a.prefetch_ctas = prefetch;
    const bool band = a.full_begin != a.row_begin || a.full_end != a.row_end || a.hist_row_lo != 0 || a.hist_row_hi != a.H;
    static const int minb = [] {  // A/B switch: RMD_TEMPORAL_CTAS=3 trades occupancy (24 warps/SM) for 85 registers (no spills)
        const char* e = getenv("RMD_TEMPORAL_CTAS");
        return e && atoi(e) == 3 ? 3 : RMD_TEMPORAL_MINB;
    }();
    if (minb == 3)
        return band ? (int)cudaLaunchKernelEx(&cfg, temporal_kernel<true, 3>, a) : (int)cudaLaunchKernelEx(&cfg, temporal_kernel<false, 3>, a);
    return band ? (int)cudaLaunchKernelEx(&cfg, temporal_kernel<true, RMD_TEMPORAL_MINB>, a)
                : (int)cudaLaunchKernelEx(&cfg, temporal_kernel<false, RMD_TEMPORAL_MINB>, a);
}

namespace {
// decoded guide only (no temporal work) for rows outside a band's temporal range: the a-trous levels read the
// guide up to 32 rows beyond the rows they produce
__global__ void guide_rows_kernel(const uint2* __restrict__ guide, float4* __restrict__ out_g4, int W, int Wp, int row_begin,
                                  int row_end) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (x < W && y < row_end) out_g4[(size_t)y * Wp + x] = decode_guide(__ldg(guide + (size_t)y * W + x));
}
}  // namespace

int launch_guide_rows(const uint2* guide, float4* out_g4, int W, int H, int Wp, int row_begin, int row_end, cudaStream_t s) {
    (void)H;
    if (row_end <= row_begin) return 0;
    dim3 block(32, 8), grid((W + 31) / 32, (row_end - row_begin + 7) / 8);
    guide_rows_kernel<<<grid, block, 0, s>>>(guide, out_g4, W, Wp, row_begin, row_end);
    return (int)cudaGetLastError();
}

}  // namespace rmd
