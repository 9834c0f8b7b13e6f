// svgf_atrous.cu — passes 3..7 of the SVGF path: one edge-avoiding a-trous wavelet
// level (5x5 B3-spline taps dilated by step = 2^level; normal, depth and luminance
// edge-stopping weights; variance propagation), DESIGN.md spec S4-S5, checked
// against oracle/oracle_svgf.c:pass_atrous.
//
// Reference hooks: the taps are the reference's `waveletSpline = {3/8, 1/4, 1/16}`
// (src/filter.cu:10); the border rule is its "skip the tap and renormalise"
// (src/filter.cu:38-39, 46, 49); the global->shared halo tile that the reference
// fills with a strided cooperative copy (`cacheTile`, src/filter.cu:60-85) is
// filled here by TMA tensor loads.
//
// Tiling (DESIGN.md "A-trous kernel"):
//   * Polyphase rows.  At step S a pixel only ever reads rows with the same
//     (y mod S), so a CTA works on ONE row phase: its tile is WT dense columns x TY
//     lattice rows (y = phase + S*k).  The planes are described to TMA as
//     {x, phase, k} tensors (strides pitch, S*pitch), so one box fetches the
//     (WT + 2*max(2S,4)) x (TY + 4) texels the tile needs: the vertical halo is 2 lattice
//     rows at every level instead of 2*S image rows.
//   * TMA zero-fills texels outside the image (and the planes' padding rows are
//     zero), which decodes to "normal = 0": the normal weight max(0, n.n')^sigma is
//     then exactly 0, i.e. the tap is skipped and the sum renormalised, with no
//     bounds test in the tap loop.
//   * Register blocking.  Each thread owns one column and 4 consecutive lattice
//     rows; the 8x5 texels it stages through registers feed 100 taps (2.5 taps per
//     shared-memory load), which keeps the loop issue-bound instead of
//     shared-memory-bandwidth-bound.  Lanes are consecutive in x, so every LDS.128
//     is conflict-free.
//   * The three edge-stopping terms and the spline weight are merged into ONE
//     exponent: w*h = 2^(sigma_n*lg2(n.n') - |dz|*iz - |dL|*il + lg2 h).
//
// Roofline: HBM for traffic (per pixel: read colour+lum 16, variance 4, guide 16,
// slope 4; write 16 + 4 = 60 B; the last level writes the 16-B output and reads
// 4 B of albedo instead = 60 B), but the kernel is fp32-issue bound (~17 issue
// slots per tap x 24 taps); both ceilings are reported by bench.py.
#include "svgf.cuh"

namespace rmd {
namespace {

constexpr int align128(int v) { return (v + 127) & ~127; }

template <int S>
struct Tile {
    // x halo: 2*S texels are needed; TMA wants every box row to start on a 16-byte
    // boundary, and the variance plane has 4-byte texels, so the halo is a multiple of 4.
    static constexpr int HX = 2 * S < 4 ? 4 : 2 * S;
    static constexpr int TW = kAtrousWT + 2 * HX;
    static constexpr int TH = kAtrousTY + 4;
    // float4 planes are staged as two half-width column blocks [2][TH][TW/2]: a TMA box
    // dimension holds at most 256 elements, so one box of 8-byte elements covers TW/2
    // texels (<= 96) per row with 1-1.5 KB rows (16-byte inner rows made TMA request-bound).
    static constexpr int HW2 = TW / 2;
    static constexpr int HALF_BYTES = HW2 * TH * 16;
    static constexpr int C4_BYTES = 2 * HALF_BYTES;
    static constexpr int V_BYTES = TW * TH * 4;
    static constexpr int OFF_C4 = 0;
    static constexpr int OFF_G4 = align128(C4_BYTES);
    static constexpr int OFF_V = OFF_G4 + align128(C4_BYTES);
    static constexpr int OFF_BAR = OFF_V + align128(V_BYTES);
    static constexpr int SMEM = OFF_BAR + 16 + 128;  // + slack to align the dynamic base to 128 B
    static constexpr uint32_t TX_BYTES = 2u * C4_BYTES + V_BYTES;
    static_assert(HALF_BYTES % 128 == 0, "second column block must stay 128-B aligned for TMA");
    static_assert(TW % 2 == 0 && 2 * HW2 <= 256, "box limit");
    // texel offset of column `col` inside a float4 plane (row 0)
    __device__ static __forceinline__ int coloff(int col) { return col < HW2 ? col : TH * HW2 + col - HW2; }
};

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

// row-indexed loads: `jr` is a compile-time constant after unrolling, the switch folds away
template <class T, int C>
__device__ __forceinline__ float4 ld_c4(uint32_t colbase, int jr) {
    switch (jr) {
        case 0: return lds128<T::OFF_C4 + 0 * T::HW2 * 16>(colbase);
        case 1: return lds128<T::OFF_C4 + 1 * T::HW2 * 16>(colbase);
        case 2: return lds128<T::OFF_C4 + 2 * T::HW2 * 16>(colbase);
        case 3: return lds128<T::OFF_C4 + 3 * T::HW2 * 16>(colbase);
        case 4: return lds128<T::OFF_C4 + 4 * T::HW2 * 16>(colbase);
        case 5: return lds128<T::OFF_C4 + 5 * T::HW2 * 16>(colbase);
        case 6: return lds128<T::OFF_C4 + 6 * T::HW2 * 16>(colbase);
        default: return lds128<T::OFF_C4 + 7 * T::HW2 * 16>(colbase);
    }
}
template <class T, int C>
__device__ __forceinline__ float4 ld_g4(uint32_t colbase, int jr) {
    switch (jr) {
        case 0: return lds128<T::OFF_G4 + 0 * T::HW2 * 16>(colbase);
        case 1: return lds128<T::OFF_G4 + 1 * T::HW2 * 16>(colbase);
        case 2: return lds128<T::OFF_G4 + 2 * T::HW2 * 16>(colbase);
        case 3: return lds128<T::OFF_G4 + 3 * T::HW2 * 16>(colbase);
        case 4: return lds128<T::OFF_G4 + 4 * T::HW2 * 16>(colbase);
        case 5: return lds128<T::OFF_G4 + 5 * T::HW2 * 16>(colbase);
        case 6: return lds128<T::OFF_G4 + 6 * T::HW2 * 16>(colbase);
        default: return lds128<T::OFF_G4 + 7 * T::HW2 * 16>(colbase);
    }
}
template <class T, int COLS>
__device__ __forceinline__ float ld_v(uint32_t vbase, int jr) {
    switch (jr) {
        case 0: return lds32<(0 * T::TW + COLS) * 4>(vbase);
        case 1: return lds32<(1 * T::TW + COLS) * 4>(vbase);
        case 2: return lds32<(2 * T::TW + COLS) * 4>(vbase);
        case 3: return lds32<(3 * T::TW + COLS) * 4>(vbase);
        case 4: return lds32<(4 * T::TW + COLS) * 4>(vbase);
        case 5: return lds32<(5 * T::TW + COLS) * 4>(vbase);
        case 6: return lds32<(6 * T::TW + COLS) * 4>(vbase);
        default: return lds32<(7 * T::TW + COLS) * 4>(vbase);
    }
}

// 3x3 Gaussian {1/4,1/8,1/16} prefilter of the variance; fixed operation order so that the
// tile and the ring kernel agree bit for bit
__device__ __forceinline__ float vbar3x3(float tm, float tc, float tp, float mm, float mc, float mp, float bm, float bc,
                                         float bp) {
    const float top = __fadd_rn(fmaf(2.0f, tc, tm), tp);
    const float mid = __fadd_rn(fmaf(2.0f, mc, mm), mp);
    const float bot = __fadd_rn(fmaf(2.0f, bc, bm), bp);
    return __fmul_rn(__fadd_rn(fmaf(2.0f, mid, top), bot), 1.0f / 16.0f);
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

template <int S>
__device__ __forceinline__ void centre_setup(Centre& ctr, Acc& acc, const float4 c, const float4 g, const float v,
                                             const float vbar, const float dz, const AtrousArgs& a) {
    // il = log2(e) / phi_l and iz_k = log2(e) / (phi_z * d_k + 1e-6), with 1/log2(e) folded into the
    // operands so that each is one FFMA + one MUFU.RCP
    const float kLn2 = 0.6931471805599453f;
    ctr.nx = g.x; ctr.ny = g.y; ctr.nz = g.z; ctr.z = g.w; ctr.L = c.w;
    ctr.il = fast_rcp(fmaf(a.sigma_l * kLn2, sqrtf(fmaxf(vbar, 0.0f)), 1e-4f * kLn2));
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

// all taps of one staged texel (column index C, tile row JR) for the 4 outputs of a thread
template <int C, int JR>
__device__ __forceinline__ void taps_of_texel(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const float4 q,
                                              const float4 g, const float v, const float sigma_n) {
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        constexpr int adx = C < 2 ? 2 - C : C - 2;
        const int dy = JR - 2 - j;
        if (dy < -2 || dy > 2) continue;
        if (dy == 0 && C == 2) continue;  // centre tap, already accumulated with w = 1
        const int ady = dy < 0 ? -dy : dy;
        if (ady == 0) tap<adx, 0>(acc[j], ctr[j], q, g, v, sigma_n);
        else if (ady == 1) tap<adx, 1>(acc[j], ctr[j], q, g, v, sigma_n);
        else tap<adx, 2>(acc[j], ctr[j], q, g, v, sigma_n);
    }
}

template <int JR>
__device__ __forceinline__ void taps_row(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const float4 q,
                                         const float4 g, const float v, const float sigma_n, int c) {
    switch (c) {
        case 0: taps_of_texel<0, JR>(acc, ctr, q, g, v, sigma_n); break;
        case 1: taps_of_texel<1, JR>(acc, ctr, q, g, v, sigma_n); break;
        case 2: taps_of_texel<2, JR>(acc, ctr, q, g, v, sigma_n); break;
        case 3: taps_of_texel<3, JR>(acc, ctr, q, g, v, sigma_n); break;
        default: taps_of_texel<4, JR>(acc, ctr, q, g, v, sigma_n); break;
    }
}

__device__ __forceinline__ void store_output(const AtrousArgs& a, const Acc& acc, const Centre& ctr, const float4 cC,
                                             const float cV, int x, int y) {
    const float inv = fast_rcp(acc.w);
    float r = acc.r * inv, g = acc.g * inv, b = acc.b * inv, v = acc.v * inv * inv;
    const bool sky = ctr.z == 0.0f;
    if (sky) { r = cC.x; g = cC.y; b = cC.z; v = cV; }
    if (a.out_c4) {
        const size_t p = (size_t)y * a.Wp + x;
        a.out_c4[p] = make_float4(r, g, b, sky ? cC.w : luminance(r, g, b));
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

template <int S>
__global__ void __launch_bounds__(kAtrousWT* kAtrousTR, 512 / (kAtrousWT * kAtrousTR))
    atrous_kernel(const AtrousArgs a, const __grid_constant__ AtrousMaps maps) {
    using T = Tile<S>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const float4* sC4 = reinterpret_cast<const float4*>(smem + T::OFF_C4);
    const float4* sG4 = reinterpret_cast<const float4*>(smem + T::OFF_G4);
    const float* sV = reinterpret_cast<const float*>(smem + T::OFF_V);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);

    const int W = a.W, H = a.H, Wp = a.Wp;
    const int tx = threadIdx.x, tr = threadIdx.y;
    const int tid = tr * kAtrousWT + tx;
    // blockIdx.y enumerates (phase, lattice tile)
    const int lat_rows_max = (H + S - 1) / S;
    const int tiles_per_phase = (lat_rows_max + kAtrousTY - 1) / kAtrousTY;
    const int phase = blockIdx.y / tiles_per_phase;
    const int k0 = (blockIdx.y - phase * tiles_per_phase) * kAtrousTY;
    const int x0 = blockIdx.x * kAtrousWT;
    if (phase + S * k0 >= H) return;  // this phase has fewer lattice rows (uniform per CTA)
    {   // band mode: skip tiles none of whose rows are produced by this launch (uniform per CTA)
        const int y_first = phase + S * k0, y_last = phase + S * (k0 + kAtrousTY - 1);
        if (y_last < a.row0 || y_first >= a.row0 + a.rows) return;
    }

    // ---- stage the tile -------------------------------------------------------------
    if (a.use_tma) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();  // the barrier must be initialised before any thread polls it
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, T::TX_BYTES);
            const int cx = 2 * (x0 - T::HX);  // 8-byte elements: 2 per texel
            tma_load_3d(smem + T::OFF_C4, &maps.c4, bar, cx, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_C4 + T::HALF_BYTES, &maps.c4, bar, cx + 2 * T::HW2, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_G4, &maps.g4, bar, cx, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_G4 + T::HALF_BYTES, &maps.g4, bar, cx + 2 * T::HW2, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_V, &maps.v, bar, x0 - T::HX, phase, k0 - 2);
        }
    } else {
        float4* wC4 = reinterpret_cast<float4*>(smem + T::OFF_C4);
        float4* wG4 = reinterpret_cast<float4*>(smem + T::OFF_G4);
        float* wV = reinterpret_cast<float*>(smem + T::OFF_V);
        for (int i = tid; i < T::TW * T::TH; i += kAtrousWT * kAtrousTR) {
            const int row = i / T::TW, col = i - row * T::TW;
            const int gx = x0 - T::HX + col, k = k0 - 2 + row;
            const int gy = phase + S * k;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f), g = c;
            float v = 0.f;
            if (gx >= 0 && gx < W && k >= 0 && gy < H) {
                const size_t q = (size_t)gy * Wp + gx;
                c = a.in_c4[q];
                g = a.g4[q];
                v = a.in_v[q];
            }
            wC4[T::coloff(col) + row * T::HW2] = c;
            wG4[T::coloff(col) + row * T::HW2] = g;
            wV[i] = v;
        }
    }

    // ---- per-output centre terms that do not come from the tile (global, L1/L2) ----
    // 3x3 Gaussian prefilter of the variance at the centre (dense neighbours: other
    // row phases, so not in the tile) and the depth slope.  Issued before the tile
    // wait so their latency overlaps the TMA.
    const int x = x0 + tx;
    const int xc = min(x, W - 1);
    // clamp-to-edge in x by selection, not by address: all 36 loads are [row pointer + immediate] and independent;
    // at x = 0 / x = W-1 the neighbour load reads the adjacent padding element (the planes carry a guard at either
    // end, svgf_ctx.cu) and its value is replaced by the centre column's
    const bool has_l = xc > 0, has_r = xc < W - 1;
    float vbar[kAtrousOPT], dzv[kAtrousOPT];
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = min(phase + S * (k0 + kAtrousOPT * tr + j), H - 1);
        const int ym = max(y - 1, 0), yp = min(y + 1, H - 1);
        const float* r0 = a.in_v + ((size_t)ym * Wp + xc);
        const float* r1 = a.in_v + ((size_t)y * Wp + xc);
        const float* r2 = a.in_v + ((size_t)yp * Wp + xc);
        const float t0 = __ldg(r0 - 1), t1 = __ldg(r0), t2 = __ldg(r0 + 1);
        const float m0 = __ldg(r1 - 1), m1 = __ldg(r1), m2 = __ldg(r1 + 1);
        const float b0 = __ldg(r2 - 1), b1 = __ldg(r2), b2 = __ldg(r2 + 1);
        vbar[j] = vbar3x3(has_l ? t0 : t1, t1, has_r ? t2 : t1, has_l ? m0 : m1, m1, has_r ? m2 : m1,
                          has_l ? b0 : b1, b1, has_r ? b2 : b1);
        dzv[j] = __ldg(a.dz + ((size_t)y * Wp + xc));
    }

    if (a.use_tma) {
        mbar_wait(bar, 0);
    } else {
        __syncthreads();
    }

    // ---- centre set-up ---------------------------------------------------------------
    Centre ctr[kAtrousOPT];
    Acc acc[kAtrousOPT];
    float4 cC[kAtrousOPT];
    float cV[kAtrousOPT];
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int row = kAtrousOPT * tr + j + 2, col = tx + T::HX;
        cC[j] = sC4[T::coloff(col) + row * T::HW2];
        const float4 g = sG4[T::coloff(col) + row * T::HW2];
        cV[j] = sV[row * T::TW + col];
        centre_setup<S>(ctr[j], acc[j], cC[j], g, cV[j], vbar[j], dzv[j], a);
    }

    // ---- 100 taps from 40 staged texels ---------------------------------------------
    const float sigma_n = a.sigma_n;
    // per-thread column bases (shared-window byte addresses); every load below is
    // [register + compile-time immediate]
    const uint32_t sbase = smem_u32(smem);
    uint32_t cb[5];
#pragma unroll
    for (int c = 0; c < 5; ++c)
        cb[c] = sbase + 16u * (uint32_t)(T::coloff(tx + (T::HX - 2 * S) + c * S) + kAtrousOPT * tr * T::HW2);
    const uint32_t vb = sbase + T::OFF_V + 4u * (uint32_t)(kAtrousOPT * tr * T::TW + tx + (T::HX - 2 * S));
#pragma unroll
    for (int jr = 0; jr < kAtrousOPT + 4; ++jr) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            float4 q, g;
            float v;
            // c is a compile-time constant after unrolling: select the immediate-offset load
            if (c == 0) { q = ld_c4<T, 0>(cb[0], jr); g = ld_g4<T, 0>(cb[0], jr); v = ld_v<T, 0 * S>(vb, jr); }
            if (c == 1) { q = ld_c4<T, 1>(cb[1], jr); g = ld_g4<T, 1>(cb[1], jr); v = ld_v<T, 1 * S>(vb, jr); }
            if (c == 2) { q = ld_c4<T, 2>(cb[2], jr); g = ld_g4<T, 2>(cb[2], jr); v = ld_v<T, 2 * S>(vb, jr); }
            if (c == 3) { q = ld_c4<T, 3>(cb[3], jr); g = ld_g4<T, 3>(cb[3], jr); v = ld_v<T, 3 * S>(vb, jr); }
            if (c == 4) { q = ld_c4<T, 4>(cb[4], jr); g = ld_g4<T, 4>(cb[4], jr); v = ld_v<T, 4 * S>(vb, jr); }
            switch (jr) {  // jr and c are compile-time after unrolling
                case 0: taps_row<0>(acc, ctr, q, g, v, sigma_n, c); break;
                case 1: taps_row<1>(acc, ctr, q, g, v, sigma_n, c); break;
                case 2: taps_row<2>(acc, ctr, q, g, v, sigma_n, c); break;
                case 3: taps_row<3>(acc, ctr, q, g, v, sigma_n, c); break;
                case 4: taps_row<4>(acc, ctr, q, g, v, sigma_n, c); break;
                case 5: taps_row<5>(acc, ctr, q, g, v, sigma_n, c); break;
                case 6: taps_row<6>(acc, ctr, q, g, v, sigma_n, c); break;
                default: taps_row<7>(acc, ctr, q, g, v, sigma_n, c); break;
            }
        }
    }

    // ---- epilogue ---------------------------------------------------------------------
    if (x >= W) return;
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = phase + S * (k0 + kAtrousOPT * tr + j);
        if (y >= a.row0 && y < a.row0 + a.rows) store_output(a, acc[j], ctr[j], cC[j], cV[j], x, y);
    }
}


// taps of one staged texel in a column with |dx| = ADX (centre column when ADX == 0), tile row JR
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

// =====================================================================================
// Ring kernel: persistent CTAs march down column strips; TMA streams 4-row chunks of the
// strip into a shared-memory ring one step ahead of the arithmetic.
//
//   * work      one "step" = 128 columns x 8 lattice rows of one row phase.  The steps of
//               the whole level are numbered (phase, strip, step-in-strip) and every CTA
//               owns a contiguous, equally long range of them, so vertically consecutive
//               steps reuse the 4 halo rows they share (each input row is fetched ~1.07x
//               instead of 1.5-2x with independent tiles).
//   * ring      NRC chunk slots of 4 rows x TW texels x 36 B (colour+lum 16, guide 16,
//               variance 4).  Chunk j of a run holds lattice rows [8*ks0 - 2 + 4j, +4).
//   * warps     8 warps per CTA: warp = (row group tr, column block); it produces 32
//               columns x 4 lattice rows per step from chunks (2s + tr, 2s + tr + 1) and
//               free-runs: it waits on the chunks' "full" mbarriers only.
//   * refill    every chunk is read by a known number of warps; the LAST warp to finish
//               with a slot (shared-memory counter) re-arms its mbarrier and issues the
//               TMA loads of the chunk NRC ahead.  No producer warp, no CTA-wide barrier
//               after start-up.
//   * extras    the centre pixel additionally needs the variance at rows y-1 / y+1 (other
//               row phases, hence not in the ring) and the depth slope: each warp gathers
//               them for its NEXT step with clamped 4-byte cp.async into a private 1.7 KB
//               buffer while the current step computes.
// =====================================================================================
template <int S>
struct Ring {
    static constexpr int HX = 2 * S < 4 ? 4 : 2 * S;
    static constexpr int TW = kAtrousWT + 2 * HX;
    static constexpr int HW2 = TW / 2;
    static constexpr int CR = 4;                                   // rows per chunk
    static constexpr int NRC = S <= 2 ? 5 : (S <= 8 ? 4 : 3);      // chunk slots (shared-memory budget, 2 CTAs/SM)
    static constexpr int HALF_BYTES = HW2 * CR * 16;
    static constexpr int OFF_C4 = 0;
    static constexpr int OFF_G4 = 2 * HALF_BYTES;
    static constexpr int OFF_V = 4 * HALF_BYTES;
    static constexpr int V_BYTES = TW * CR * 4;
    static constexpr int CHUNK_BYTES = 4 * HALF_BYTES + V_BYTES;
    static constexpr int CHUNK_STRIDE = align128(CHUNK_BYTES);
    static constexpr int EX_COLS = 36;                              // 34 used: x-1 .. x+32
    static constexpr int EX_ROWS = 12;                              // 4 x V(y-1), 4 x V(y+1), 4 x dz
    static constexpr int EX_BYTES = EX_ROWS * EX_COLS * 4;
    static constexpr int NWARPS = 8;
    static constexpr int OFF_EX = NRC * CHUNK_STRIDE;
    static constexpr int OFF_BAR = OFF_EX + NWARPS * EX_BYTES;      // full[NRC] (8 B each) then cnt[NRC] (4 B each)
    static constexpr int SMEM = align128(OFF_BAR + NRC * 12) + 128;
    static_assert(HALF_BYTES % 128 == 0 && (4 * HALF_BYTES) % 128 == 0, "TMA destinations must be 128-B aligned");
    static_assert(OFF_BAR % 8 == 0, "mbarrier alignment");
    __device__ static __forceinline__ int coloff_bytes(int col) {
        return col < HW2 ? col * 16 : HALF_BYTES + (col - HW2) * 16;
    }
};

// One column of the 5x5 footprint for the 4 outputs of a thread: 8 staged texels (tile rows
// 0..3 from the first chunk, 4..7 from the second), all addresses [register + immediate].
template <int S, int ADX>
__device__ __forceinline__ void ring_column(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const uint32_t b0,
                                            const uint32_t b1, const uint32_t v0, const uint32_t v1, const float sigma_n) {
    using R = Ring<S>;
#define RMD_RING_LOAD(JR, B, V, Q, G, VV)                                    \
    const float4 Q = lds128<R::OFF_C4 + ((JR)&3) * R::HW2 * 16>(B);          \
    const float4 G = lds128<R::OFF_G4 + ((JR)&3) * R::HW2 * 16>(B);          \
    const float VV = lds32<R::OFF_V + ((JR)&3) * R::TW * 4>(V);
    // loads run one texel ahead of the arithmetic that consumes them
    RMD_RING_LOAD(0, b0, v0, q0, g0, w0)
    RMD_RING_LOAD(1, b0, v0, q1, g1, w1)
    taps_of_texel_adx<ADX, 0>(acc, ctr, q0, g0, w0, sigma_n);
    RMD_RING_LOAD(2, b0, v0, q2, g2, w2)
    taps_of_texel_adx<ADX, 1>(acc, ctr, q1, g1, w1, sigma_n);
    RMD_RING_LOAD(3, b0, v0, q3, g3, w3)
    taps_of_texel_adx<ADX, 2>(acc, ctr, q2, g2, w2, sigma_n);
    RMD_RING_LOAD(4, b1, v1, q4, g4, w4)
    taps_of_texel_adx<ADX, 3>(acc, ctr, q3, g3, w3, sigma_n);
    RMD_RING_LOAD(5, b1, v1, q5, g5, w5)
    taps_of_texel_adx<ADX, 4>(acc, ctr, q4, g4, w4, sigma_n);
    RMD_RING_LOAD(6, b1, v1, q6, g6, w6)
    taps_of_texel_adx<ADX, 5>(acc, ctr, q5, g5, w5, sigma_n);
    RMD_RING_LOAD(7, b1, v1, q7, g7, w7)
    taps_of_texel_adx<ADX, 6>(acc, ctr, q6, g6, w6, sigma_n);
    taps_of_texel_adx<ADX, 7>(acc, ctr, q7, g7, w7, sigma_n);
#undef RMD_RING_LOAD
}

struct RingWork {
    int t0, t1, nsteps, nbx;
};

// chunk g of this CTA's work range -> (strip id sigma, first lattice row); false past the end.
// Runs: the first one may start mid-strip, every later one starts at the top of a strip.
__device__ __forceinline__ bool ring_map_chunk(const RingWork& w, int g, int& sigma, int& row) {
    const int ks0 = w.t0 % w.nsteps;
    const int n0 = min(w.nsteps - ks0, w.t1 - w.t0);
    if (g < 2 * n0 + 1) {
        sigma = w.t0 / w.nsteps;
        row = 8 * ks0 - 2 + 4 * g;
        return true;
    }
    g -= 2 * n0 + 1;
    const int per = 2 * w.nsteps + 1;         // chunks of a full strip
    const int r = g / per, j = g - r * per;   // r-th later run, chunk j inside it
    const int start = w.t0 + n0 + r * w.nsteps;
    if (start >= w.t1) return false;
    const int n = min(w.nsteps, w.t1 - start);
    if (j >= 2 * n + 1) return false;
    sigma = start / w.nsteps;
    row = 4 * j - 2;
    return true;
}

template <int S>
__device__ __noinline__ void ring_issue_chunk(const RingWork w, const AtrousMaps* maps, uint32_t sbase, int g) {
    using R = Ring<S>;
    int sigma, row;
    if (!ring_map_chunk(w, g, sigma, row)) return;
    const int phase = sigma / w.nbx, bx = sigma - phase * w.nbx;
    const int slot = g % R::NRC;
    const uint32_t dst = sbase + slot * R::CHUNK_STRIDE;
    const uint32_t full = sbase + R::OFF_BAR + 8 * slot;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(R::CHUNK_BYTES) : "memory");
    const int x0 = bx * kAtrousWT - R::HX;
#define RMD_TMA3(DST, MAP, C0)                                                                                          \
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" \
                 ::"r"(DST), "l"(MAP), "r"(full), "r"(C0), "r"(phase), "r"(row) : "memory")
    RMD_TMA3(dst + R::OFF_C4, &maps->c4, 2 * x0);
    RMD_TMA3(dst + R::OFF_C4 + R::HALF_BYTES, &maps->c4, 2 * (x0 + R::HW2));
    RMD_TMA3(dst + R::OFF_G4, &maps->g4, 2 * x0);
    RMD_TMA3(dst + R::OFF_G4 + R::HALF_BYTES, &maps->g4, 2 * (x0 + R::HW2));
    RMD_TMA3(dst + R::OFF_V, &maps->v, x0);
#undef RMD_TMA3
}

// gathers (clamped) V(y-1), V(y+1) and dz for the 4 outputs of the warp's step into its private buffer
template <int S>
__device__ __forceinline__ void ring_issue_extras(const AtrousArgs& a, uint32_t ex, int lane, int xw0, int phase,
                                                  int kfirst) {
    using R = Ring<S>;
    const int W = a.W, H = a.H, Wp = a.Wp;
    const int xs = min(xw0 + lane, W - 1);
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = min(phase + S * (kfirst + j), H - 1);
        const int ym = max(y - 1, 0), yp = min(y + 1, H - 1);
        cp_async_4(ex + ((j)*R::EX_COLS + lane + 1) * 4, a.in_v + (size_t)ym * Wp + xs);
        cp_async_4(ex + ((4 + j) * R::EX_COLS + lane + 1) * 4, a.in_v + (size_t)yp * Wp + xs);
        cp_async_4(ex + ((8 + j) * R::EX_COLS + lane + 1) * 4, a.dz + (size_t)y * Wp + xs);
    }
    if (lane < 16) {  // the two edge columns (x-1 of lane 0, x+1 of lane 31) of the 8 variance rows
        const int r = lane & 7, right = lane >> 3;
        const int j = r & 3;
        const int y = min(phase + S * (kfirst + j), H - 1);
        const int yy = r < 4 ? max(y - 1, 0) : min(y + 1, H - 1);
        const int xe = right ? min(xw0 + 32, W - 1) : max(min(xw0, W - 1) - 1, 0);
        cp_async_4(ex + (r * R::EX_COLS + (right ? 33 : 0)) * 4, a.in_v + (size_t)yy * Wp + xe);
    }
}

template <int S>
__global__ void __launch_bounds__(256, 2) atrous_ring_kernel(const AtrousArgs a, const __grid_constant__ AtrousMaps maps) {
    using R = Ring<S>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + R::OFF_BAR);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + R::OFF_BAR + R::NRC * 8);
    const uint32_t sbase = smem_u32(smem);

    const int W = a.W, H = a.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tr = warp >> 2;                 // row group: lattice rows 4*tr .. 4*tr+3 of the step
    const int tx = (warp & 3) * 32 + lane;    // column inside the strip
    RingWork w;
    {
        const int lat_rows = (H + S - 1) / S;
        w.nsteps = (lat_rows + 7) / 8;
        w.nbx = (W + kAtrousWT - 1) / kAtrousWT;
        const int phases = S < H ? S : H;
        const long long T = (long long)phases * w.nbx * w.nsteps;
        w.t0 = (int)(T * blockIdx.x / gridDim.x);
        w.t1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
    }
    if (w.t0 >= w.t1) return;

    if (tid == 0) {
        for (int i = 0; i < R::NRC; ++i) {
            mbar_init(&full[i], 1);
            cnt[i] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int g = 0; g < R::NRC; ++g) ring_issue_chunk<S>(w, &maps, sbase, g);

    const uint32_t ex = sbase + R::OFF_EX + warp * R::EX_BYTES;
    const float* exf = reinterpret_cast<const float*>(smem + R::OFF_EX + warp * R::EX_BYTES);
    // per-thread column offsets inside a chunk slot
    uint32_t co[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) co[c] = (uint32_t)R::coloff_bytes(tx + (R::HX - 2 * S) + c * S);
    const uint32_t vco = 4u * (uint32_t)(tx + (R::HX - 2 * S));
    const float sigma_n = a.sigma_n;

    // (phase, strip column, step-in-strip) of the current and of the next step, advanced incrementally
    int ks = w.t0 % w.nsteps, bx, phase;
    {
        const int sigma = w.t0 / w.nsteps;
        phase = sigma / w.nbx;
        bx = sigma - phase * w.nbx;
    }
    int ks_n = ks, bx_n = bx, phase_n = phase;
    auto advance = [&](int& k, int& b, int& p) {
        if (++k == w.nsteps) {
            k = 0;
            if (++b == w.nbx) { b = 0; ++p; }
        }
    };
    advance(ks_n, bx_n, phase_n);
    const int xw = (warp & 3) * 32;
    ring_issue_extras<S>(a, ex, lane, bx * kAtrousWT + xw, phase, 8 * ks + 4 * tr);  // extras of the first step

    int gbase = 0, run_start = w.t0;
    int run_n = min(w.nsteps - ks, w.t1 - w.t0);
    for (int t = w.t0; t < w.t1; ++t) {
        if (t == run_start + run_n) {  // next strip: its chunks follow in the chunk stream
            gbase += 2 * run_n + 1;
            run_start = t;
            run_n = min(w.nsteps, w.t1 - t);
        }
        const int s = t - run_start;
        const int x = bx * kAtrousWT + tx;
        const int kfirst = 8 * ks + 4 * tr;
        const int out_phase = phase;
        const int j0 = 2 * s + tr;                      // run-local index of the warp's first chunk
        const int g0 = gbase + j0;
        const int slot0 = g0 % R::NRC, slot1 = (g0 + 1) % R::NRC;
        const uint32_t sb0 = sbase + slot0 * R::CHUNK_STRIDE, sb1 = sbase + slot1 * R::CHUNK_STRIDE;

        // ---- extras gathered during the previous step ----
        cp_async_wait_all();
        __syncwarp();
        float vbar[kAtrousOPT], dzv[kAtrousOPT], vup[kAtrousOPT][3], vdn[kAtrousOPT][3];
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                vup[j][d] = exf[j * R::EX_COLS + lane + d];
                vdn[j][d] = exf[(4 + j) * R::EX_COLS + lane + d];
            }
            dzv[j] = exf[(8 + j) * R::EX_COLS + lane + 1];
        }
        __syncwarp();
        if (t + 1 < w.t1)  // gather for the next step while this one computes
            ring_issue_extras<S>(a, ex, lane, bx_n * kAtrousWT + xw, phase_n, 8 * ks_n + 4 * tr);
        ks = ks_n; bx = bx_n; phase = phase_n;
        advance(ks_n, bx_n, phase_n);

        // ---- wait for the two chunks of this step ----
        mbar_wait(&full[slot0], (uint32_t)((g0 / R::NRC) & 1));
        mbar_wait(&full[slot1], (uint32_t)(((g0 + 1) / R::NRC) & 1));

        // ---- centre set-up (tile rows 2..5 of the warp's 8-row window) ----
        Centre ctr[kAtrousOPT];
        Acc acc[kAtrousOPT];
        const uint32_t cco = (uint32_t)R::coloff_bytes(tx + R::HX);
        // own-row neighbours clamp at the image edge (TMA zero-fills there, the spec clamps)
        const int xl = x > 0 ? -4 : 0, xr = x < W - 1 ? 4 : 0;
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
            const int jr = j + 2;
            const uint32_t sb = jr < 4 ? sb0 : sb1;
            const int rr = jr & 3;
            const float4 c = lds128_dyn(sb + cco + R::OFF_C4 + rr * R::HW2 * 16);
            const float4 g = lds128_dyn(sb + cco + R::OFF_G4 + rr * R::HW2 * 16);
            const uint32_t va = sb + R::OFF_V + (uint32_t)(rr * R::TW + tx + R::HX) * 4u;
            const float vc = lds32_dyn(va);
            const float vm = lds32_dyn(va + xl), vp = lds32_dyn(va + xr);
            vbar[j] = vbar3x3(vup[j][0], vup[j][1], vup[j][2], vm, vc, vp, vdn[j][0], vdn[j][1], vdn[j][2]);
            centre_setup<S>(ctr[j], acc[j], c, g, vc, vbar[j], dzv[j], a);
        }

        // ---- 100 taps from 40 staged texels: columns grouped by |dx| so that the unrolled
        //      body is 56 taps instead of 100 (instruction-cache footprint) ----
        {
            const uint32_t vb0 = sb0 + vco, vb1 = sb1 + vco;
#pragma unroll 1
            for (int it = 0; it < 2; ++it) {  // |dx| = 2: columns 0 and 4
                const uint32_t o = it ? co[4] : co[0], vo = it ? 16u * S : 0u;
                ring_column<S, 2>(acc, ctr, sb0 + o, sb1 + o, vb0 + vo, vb1 + vo, sigma_n);
            }
#pragma unroll 1
            for (int it = 0; it < 2; ++it) {  // |dx| = 1: columns 1 and 3
                const uint32_t o = it ? co[3] : co[1], vo = it ? 12u * S : 4u * S;
                ring_column<S, 1>(acc, ctr, sb0 + o, sb1 + o, vb0 + vo, vb1 + vo, sigma_n);
            }
            ring_column<S, 0>(acc, ctr, sb0 + co[2], sb1 + co[2], vb0 + 8u * S, vb1 + 8u * S, sigma_n);
        }

        // ---- centre colour/variance again (sky pass-through needs the exact input; keeping them
        //      in registers through the tap loop would cost 20 registers) ----
        float4 cC[kAtrousOPT];
        float cV[kAtrousOPT];
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
            const int jr = j + 2;
            const uint32_t sb = jr < 4 ? sb0 : sb1;
            const int rr = jr & 3;
            cC[j] = lds128_dyn(sb + cco + R::OFF_C4 + rr * R::HW2 * 16);
            cV[j] = lds32_dyn(sb + R::OFF_V + (uint32_t)(rr * R::TW + tx + R::HX) * 4u);
        }

        // ---- release the two chunks; the last reader of a slot refills it ----
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + u, g = g0 + u;
                const uint32_t expected = (j == 0 || j == 2 * run_n) ? 4u : 8u;
                const int slot = g % R::NRC;
                if (atomicAdd(&cnt[slot], 1u) == expected - 1u) {
                    cnt[slot] = 0u;
                    fence_proxy_async();
                    ring_issue_chunk<S>(w, &maps, sbase, g + R::NRC);
                }
            }
        }

        // ---- epilogue ----
        if (x < W) {
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                const int y = out_phase + S * (kfirst + j);
                if (y >= a.row0 && y < a.row0 + a.rows) store_output(a, acc[j], ctr[j], cC[j], cV[j], x, y);
            }
        }
    }
}

__global__ void remodulate_kernel(const float4* __restrict__ c4, const float* __restrict__ v,
                                  const float4* __restrict__ g4, const uchar4* __restrict__ albedo, float4* out,
                                  uchar4* out8, int W, int H, int Wp, float afloor) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t pi = (size_t)y * Wp + x, po = (size_t)y * W + x;
    const float4 c = c4[pi];
    float r = c.x, g = c.y, b = c.z;
    if (g4[pi].w != 0.0f) {
        const uchar4 al = albedo[po];
        r *= fmaxf(__fmul_rn((float)al.x, 1.0f / 255.0f), afloor);
        g *= fmaxf(__fmul_rn((float)al.y, 1.0f / 255.0f), afloor);
        b *= fmaxf(__fmul_rn((float)al.z, 1.0f / 255.0f), afloor);
    }
    out[po] = make_float4(r, g, b, v[pi]);
    if (out8)
        out8[po] = make_uchar4((unsigned char)(__saturatef(r) * 255.0f), (unsigned char)(__saturatef(g) * 255.0f),
                               (unsigned char)(__saturatef(b) * 255.0f), 255);
}

template <int S>
int launch_level(const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s) {
    const int lat_rows_max = (a.H + S - 1) / S;
    const int tiles_per_phase = (lat_rows_max + kAtrousTY - 1) / kAtrousTY;
    const int phases = S < a.H ? S : a.H;
    dim3 grid((a.W + kAtrousWT - 1) / kAtrousWT, phases * tiles_per_phase);
    dim3 block(kAtrousWT, kAtrousTR);
    atrous_kernel<S><<<grid, block, Tile<S>::SMEM, s>>>(a, maps);
    return (int)cudaGetLastError();
}

int g_num_sms = 0;

template <int S>
int launch_ring(const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s) {
    const int lat_rows = (a.H + S - 1) / S;
    const long long T = (long long)(S < a.H ? S : a.H) * ((a.W + kAtrousWT - 1) / kAtrousWT) * ((lat_rows + 7) / 8);
    const long long slots = 2LL * g_num_sms;
    const int grid = (int)(T < slots ? T : slots);
    atrous_ring_kernel<S><<<grid, 256, Ring<S>::SMEM, s>>>(a, maps);
    return (int)cudaGetLastError();
}

}  // namespace

int atrous_configure() {
    int dev = 0;
    RMD_CUDA_TRY(cudaGetDevice(&dev));
    RMD_CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<1>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<2>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<4>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<8>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<16>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<1>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<2>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<4>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<8>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<16>::SMEM));
    return 0;
}

int launch_atrous_ring(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s) {
    switch (level) {
        case 0: return launch_ring<1>(a, maps, s);
        case 1: return launch_ring<2>(a, maps, s);
        case 2: return launch_ring<4>(a, maps, s);
        case 3: return launch_ring<8>(a, maps, s);
        case 4: return launch_ring<16>(a, maps, s);
        default: return RMD_E_PARAM;
    }
}

int launch_atrous(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s) {
    switch (level) {
        case 0: return launch_level<1>(a, maps, s);
        case 1: return launch_level<2>(a, maps, s);
        case 2: return launch_level<4>(a, maps, s);
        case 3: return launch_level<8>(a, maps, s);
        case 4: return launch_level<16>(a, maps, s);
        default: return RMD_E_PARAM;
    }
}

int launch_remodulate(const float4* c4, const float* v, const float4* g4, const uchar4* albedo, float4* out,
                      uchar4* out8, int W, int H, int Wp, float afloor, cudaStream_t s) {
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    remodulate_kernel<<<grid, block, 0, s>>>(c4, v, g4, albedo, out, out8, W, H, Wp, afloor);
    return (int)cudaGetLastError();
}

}  // namespace rmd
