// svgf_atrous_tile.cu — passes 3..7 of the SVGF path, independent-tile kernel: one edge-avoiding
// a-trous wavelet level (5x5 B3-spline taps dilated by step = 2^level; normal, depth and luminance
// edge-stopping weights; variance propagation), DESIGN.md spec S4-S5, checked against
// oracle/oracle_svgf.c:pass_atrous.
//
// Reference hooks: the global->shared halo tile that the reference fills with a strided
// cooperative copy (`cacheTile`, src/filter.cu:60-85) is filled here by TMA tensor loads; taps and
// border rule: svgf_atrous.cuh.
//
// Tiling (DESIGN.md "A-trous kernel"):
//   * Polyphase rows.  At step S a pixel only ever reads rows with the same (y mod S), so a CTA
//     works on ONE row phase: its tile is WT dense columns x TY lattice rows (y = phase + S*k).
//     The planes are described to TMA as {x, phase, k} tensors (strides pitch, S*pitch), so one
//     box fetches the (WT + 2*max(2S,4)) x (TY + 4) texels the tile needs: the vertical halo is
//     2 lattice rows at every level instead of 2*S image rows.
//   * TMA zero-fills texels outside the image (and the planes' padding rows are zero), which
//     decodes to "normal = 0": the normal weight max(0, n.n')^sigma is then exactly 0, i.e. the
//     tap is skipped and the sum renormalised, with no bounds test in the tap loop.
//   * Centre terms from shared memory (MODE & 1).  The 3x3 variance pre-filter of an output needs
//     the image rows y-1 and y+1, which belong to the neighbouring row phases and are therefore
//     not in the tile: two more TMA boxes on the variance map (phase -/+ 1, wrapping to the
//     previous / next lattice row at the phase ends; at step 1 the rows are already in the tile)
//     and one on the slope map deliver them, so the prologue has no global load, no 64-bit address
//     arithmetic and a single wait.  (Round 1 issued 40 `ld.global.nc` per thread here: a quarter
//     of the kernel's instructions and half of its warp-stall samples, profiles/r2_notes.md.)
//   * Register blocking.  Each thread owns one column and 4 consecutive lattice rows; the 8x5
//     texels it stages through registers feed 100 taps (2.5 taps per shared-memory load).
//     Lanes are consecutive in x, so every LDS.128 is conflict-free.
//   * MODE & 2: the 5 tap columns are walked grouped by |dx| (2 x 20 + 2 x 20 + 16 taps in three
//     unrolled bodies) instead of as one 100-tap body: 43 KB of code becomes ~25 KB.
//
// This file is compiled once per variant (-DRMD_VARIANT=n, build.py); svgf_atrous.cu dispatches.
#include "svgf_atrous.cuh"

#ifndef RMD_VARIANT
#define RMD_VARIANT 1
#endif

namespace rmd {
namespace {

#if RMD_VARIANT == 0
constexpr int kMode = 0, kMinB = 4;
#elif RMD_VARIANT == 1
constexpr int kMode = 1, kMinB = 4;
#elif RMD_VARIANT == 2
constexpr int kMode = 3, kMinB = 4;
#elif RMD_VARIANT == 3
constexpr int kMode = 2, kMinB = 4;
#elif RMD_VARIANT == 4
constexpr int kMode = 1, kMinB = 3;
#elif RMD_VARIANT == 5
constexpr int kMode = 3, kMinB = 3;
#else
#error "unknown RMD_VARIANT"
#endif

template <int S, int MODE>
struct Tile {
    static constexpr bool PRO = (MODE & 1) != 0;
    static constexpr bool GROUPED = (MODE & 2) != 0;
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
    // neighbour-phase variance rows (y-1 / y+1 of the TY output rows), columns x0-4 .. x0+WT+3
    static constexpr int NBW = kAtrousWT + 8;
    static constexpr int NB_BYTES = (PRO && S > 1) ? NBW * kAtrousTY * 4 : 0;
    static constexpr int DZ_BYTES = PRO ? kAtrousWT * kAtrousTY * 4 : 0;
    static constexpr int OFF_C4 = 0;
    static constexpr int OFF_G4 = align128(C4_BYTES);
    static constexpr int OFF_V = OFF_G4 + align128(C4_BYTES);
    static constexpr int OFF_VM = OFF_V + align128(V_BYTES);
    static constexpr int OFF_VP = OFF_VM + align128(NB_BYTES);
    static constexpr int OFF_DZ = OFF_VP + align128(NB_BYTES);
    static constexpr int OFF_BAR = OFF_DZ + align128(DZ_BYTES);
    static constexpr int SMEM = OFF_BAR + 16 + 128;  // + slack to align the dynamic base to 128 B
    static constexpr uint32_t TX_BYTES = 2u * C4_BYTES + V_BYTES + 2u * NB_BYTES + DZ_BYTES;
    static_assert(HALF_BYTES % 128 == 0, "second column block must stay 128-B aligned for TMA");
    static_assert(TW % 2 == 0 && 2 * HW2 <= 256, "box limit");
    static_assert((NBW * 4) % 16 == 0 && (kAtrousWT * 4) % 16 == 0, "TMA inner box bytes");
    // texel offset of column `col` inside a float4 plane (row 0)
    __device__ static __forceinline__ int coloff(int col) { return col < HW2 ? col : TH * HW2 + col - HW2; }
};

// One column of the 5x5 footprint for the kAtrousOPT outputs of a thread: TH staged texels, all
// addresses [register + immediate]; loads run one texel ahead of the arithmetic that consumes them.
template <class T, int ADX>
__device__ __forceinline__ void tile_column(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const uint32_t b,
                                            const uint32_t vb, const float sigma_n) {
    static_assert(T::TH == 8, "unrolled for 4 outputs per thread");
#define RMD_TILE_LOAD(JR, Q, G, VV)                                   \
    const float4 Q = lds128<T::OFF_C4 + (JR)*T::HW2 * 16>(b);         \
    const float4 G = lds128<T::OFF_G4 + (JR)*T::HW2 * 16>(b);         \
    const float VV = lds32<T::OFF_V + (JR)*T::TW * 4>(vb);
    RMD_TILE_LOAD(0, q0, g0, w0)
    RMD_TILE_LOAD(1, q1, g1, w1)
    taps_of_texel_adx<ADX, 0>(acc, ctr, q0, g0, w0, sigma_n);
    RMD_TILE_LOAD(2, q2, g2, w2)
    taps_of_texel_adx<ADX, 1>(acc, ctr, q1, g1, w1, sigma_n);
    RMD_TILE_LOAD(3, q3, g3, w3)
    taps_of_texel_adx<ADX, 2>(acc, ctr, q2, g2, w2, sigma_n);
    RMD_TILE_LOAD(4, q4, g4, w4)
    taps_of_texel_adx<ADX, 3>(acc, ctr, q3, g3, w3, sigma_n);
    RMD_TILE_LOAD(5, q5, g5, w5)
    taps_of_texel_adx<ADX, 4>(acc, ctr, q4, g4, w4, sigma_n);
    RMD_TILE_LOAD(6, q6, g6, w6)
    taps_of_texel_adx<ADX, 5>(acc, ctr, q5, g5, w5, sigma_n);
    RMD_TILE_LOAD(7, q7, g7, w7)
    taps_of_texel_adx<ADX, 6>(acc, ctr, q6, g6, w6, sigma_n);
    taps_of_texel_adx<ADX, 7>(acc, ctr, q7, g7, w7, sigma_n);
#undef RMD_TILE_LOAD
}

template <int S, int MODE, int MINB>
__global__ void __launch_bounds__(kAtrousWT* kAtrousTR, MINB)
    atrous_kernel(const AtrousArgs a, const __grid_constant__ AtrousMaps maps) {
    using T = Tile<S, MODE>;
    static_assert(kAtrousTR == 1, "one thread row per CTA");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);

    // PDL: nothing below may run ahead of the previous kernel in the stream.  The wait sits before the early
    // exits so that this grid can never complete before its predecessor has (a grid whose CTAs all returned
    // early would otherwise release ITS successor too soon).
    pdl_wait();
    const int W = a.W, H = a.H, Wp = a.Wp;
    const int tx = threadIdx.x;
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
    // neighbour-phase rows: image row y-1 of lattice row k is (phase-1, k), or (S-1, k-1) when phase == 0;
    // image row y+1 is (phase+1, k), or (0, k+1) when phase == S-1
    const int pm = phase > 0 ? phase - 1 : S - 1, km = phase > 0 ? k0 : k0 - 1;
    const int pp = phase < S - 1 ? phase + 1 : 0, kp = phase < S - 1 ? k0 : k0 + 1;

    // ---- stage the tile -------------------------------------------------------------
    if (a.use_tma) {
        if (tx == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();  // the barrier must be initialised before any thread polls it
        if (tx == 0) {
            mbar_arrive_expect_tx(bar, T::TX_BYTES);
            const int cx = 2 * (x0 - T::HX);  // 8-byte elements: 2 per texel
            tma_load_3d(smem + T::OFF_C4, &maps.c4, bar, cx, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_C4 + T::HALF_BYTES, &maps.c4, bar, cx + 2 * T::HW2, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_G4, &maps.g4, bar, cx, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_G4 + T::HALF_BYTES, &maps.g4, bar, cx + 2 * T::HW2, phase, k0 - 2);
            tma_load_3d(smem + T::OFF_V, &maps.v, bar, x0 - T::HX, phase, k0 - 2);
            if constexpr (T::PRO) {
                if constexpr (S > 1) {
                    tma_load_3d(smem + T::OFF_VM, &maps.vn, bar, x0 - 4, pm, km);
                    tma_load_3d(smem + T::OFF_VP, &maps.vn, bar, x0 - 4, pp, kp);
                }
                tma_load_3d(smem + T::OFF_DZ, &maps.dzm, bar, x0, phase, k0);
            }
        }
    } else {
        float4* wC4 = reinterpret_cast<float4*>(smem + T::OFF_C4);
        float4* wG4 = reinterpret_cast<float4*>(smem + T::OFF_G4);
        float* wV = reinterpret_cast<float*>(smem + T::OFF_V);
        for (int i = tx; i < T::TW * T::TH; i += kAtrousWT) {
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
        if constexpr (T::PRO) {  // the same boxes the TMA path fetches, zero outside the image
            float* wDZ = reinterpret_cast<float*>(smem + T::OFF_DZ);
            for (int i = tx; i < kAtrousWT * kAtrousTY; i += kAtrousWT) {
                const int row = i / kAtrousWT, col = i - row * kAtrousWT;
                const int gx = x0 + col, gy = phase + S * (k0 + row);
                wDZ[i] = (gx < W && gy < H) ? a.dz[(size_t)gy * Wp + gx] : 0.f;
            }
            if constexpr (S > 1) {
                float* wVM = reinterpret_cast<float*>(smem + T::OFF_VM);
                float* wVP = reinterpret_cast<float*>(smem + T::OFF_VP);
                for (int i = tx; i < T::NBW * kAtrousTY; i += kAtrousWT) {
                    const int row = i / T::NBW, col = i - row * T::NBW;
                    const int gx = x0 - 4 + col;
                    const int ym = pm + S * (km + row), yp = pp + S * (kp + row);
                    const bool xin = gx >= 0 && gx < W;
                    wVM[i] = (xin && ym >= 0 && ym < H) ? a.in_v[(size_t)ym * Wp + gx] : 0.f;
                    wVP[i] = (xin && yp >= 0 && yp < H) ? a.in_v[(size_t)yp * Wp + gx] : 0.f;
                }
            }
        }
    }
    pdl_launch_dependents();

    const int x = x0 + tx;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ccol = sbase + 16u * (uint32_t)T::coloff(tx + T::HX);                    // centre column, float4 planes
    const uint32_t vcol = sbase + T::OFF_V + 4u * (uint32_t)(tx + T::HX);                   // centre column, variance
    float vbar[kAtrousOPT], dzv[kAtrousOPT];
    if constexpr (!T::PRO) {
        // ---- per-output centre terms from global memory (L1/L2), issued before the tile wait ----
        // clamp-to-edge in x by selection, not by address: all 36 loads are [row pointer + immediate] and
        // independent; at x = 0 / x = W-1 the neighbour load reads the adjacent padding element (the planes
        // carry a guard at either end, svgf_ctx.cu) and its value is replaced by the centre column's
        const int xc = min(x, W - 1);
        const bool has_l = xc > 0, has_r = xc < W - 1;
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
            const int y = min(phase + S * (k0 + j), H - 1);
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
    }

    if (a.use_tma) {
        mbar_wait(bar, 0);
    } else {
        __syncthreads();
    }

    if constexpr (T::PRO) {
        // ---- per-output centre terms from the staged rows: 3x3 Gaussian of the variance with clamped
        //      coordinates (spec S4) and the depth slope ----
        const bool has_l = x > 0, has_r = x < W - 1;
        const uint32_t nb = sbase + 4u * (uint32_t)(tx + 4);
        const uint32_t dzb = sbase + T::OFF_DZ + 4u * (uint32_t)tx;
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
            const int y = phase + S * (k0 + j);
            float t0, t1, t2, b0, b1, b2;
            const float m0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW - 1));
            const float m1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW));
            const float m2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW + 1));
            if constexpr (S > 1) {
                t0 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW - 1));
                t1 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW));
                t2 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW + 1));
                b0 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW - 1));
                b1 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW));
                b2 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW + 1));
            } else {  // step 1: the neighbouring image rows are the neighbouring tile rows
                t0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW - 1));
                t1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW));
                t2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW + 1));
                b0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW - 1));
                b1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW));
                b2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW + 1));
            }
            if (y <= 0) { t0 = m0; t1 = m1; t2 = m2; }        // row y-1 clamps to row y at the top edge
            if (y >= H - 1) { b0 = m0; b1 = m1; b2 = m2; }    // row y+1 clamps to row y at the bottom edge
            vbar[j] = vbar3x3(has_l ? t0 : t1, t1, has_r ? t2 : t1, has_l ? m0 : m1, m1, has_r ? m2 : m1,
                              has_l ? b0 : b1, b1, has_r ? b2 : b1);
            dzv[j] = lds32_dyn(dzb + 4u * (uint32_t)(j * kAtrousWT));
        }
    }

    // ---- centre set-up ---------------------------------------------------------------
    Centre ctr[kAtrousOPT];
    Acc acc[kAtrousOPT];
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const float4 c = lds128_dyn(ccol + T::OFF_C4 + 16u * (uint32_t)((j + 2) * T::HW2));
        const float4 g = lds128_dyn(ccol + T::OFF_G4 + 16u * (uint32_t)((j + 2) * T::HW2));
        const float v = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW));
        centre_setup<S>(ctr[j], acc[j], c, g, v, vbar[j], dzv[j], a);
    }

    // ---- 100 taps from 40 staged texels ---------------------------------------------
    const float sigma_n = a.sigma_n;
    // per-thread column bases (shared-window byte addresses); every load below is
    // [register + compile-time immediate]
    uint32_t cb[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) cb[c] = sbase + 16u * (uint32_t)T::coloff(tx + (T::HX - 2 * S) + c * S);
    const uint32_t vb = sbase + 4u * (uint32_t)(tx + (T::HX - 2 * S));
    if constexpr (T::GROUPED) {
#pragma unroll 1
        for (int it = 0; it < 2; ++it)  // |dx| = 2: columns 0 and 4
            tile_column<T, 2>(acc, ctr, it ? cb[4] : cb[0], vb + (it ? 16u * S : 0u), sigma_n);
#pragma unroll 1
        for (int it = 0; it < 2; ++it)  // |dx| = 1: columns 1 and 3
            tile_column<T, 1>(acc, ctr, it ? cb[3] : cb[1], vb + (it ? 12u * S : 4u * S), sigma_n);
        tile_column<T, 0>(acc, ctr, cb[2], vb + 8u * S, sigma_n);
    } else {
        tile_column<T, 2>(acc, ctr, cb[0], vb, sigma_n);
        tile_column<T, 1>(acc, ctr, cb[1], vb + 4u * S, sigma_n);
        tile_column<T, 0>(acc, ctr, cb[2], vb + 8u * S, sigma_n);
        tile_column<T, 1>(acc, ctr, cb[3], vb + 12u * S, sigma_n);
        tile_column<T, 2>(acc, ctr, cb[4], vb + 16u * S, sigma_n);
    }

    // ---- epilogue ---------------------------------------------------------------------
    if (x >= W) return;
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = phase + S * (k0 + j);
        if (y >= a.row0 && y < a.row0 + a.rows)
            store_output(a, acc[j], ctr[j], ccol + T::OFF_C4 + 16u * (uint32_t)((j + 2) * T::HW2),
                         vcol + 4u * (uint32_t)((j + 2) * T::TW), x, y);
    }
}

template <int S>
int launch_level(const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s, bool pdl) {
    const int lat_rows_max = (a.H + S - 1) / S;
    const int tiles_per_phase = (lat_rows_max + kAtrousTY - 1) / kAtrousTY;
    const int phases = S < a.H ? S : a.H;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((a.W + kAtrousWT - 1) / kAtrousWT, phases * tiles_per_phase);
    cfg.blockDim = dim3(kAtrousWT, kAtrousTR);
    cfg.dynamicSmemBytes = Tile<S, kMode>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, atrous_kernel<S, kMode, kMinB>, a, maps);
}

template <int S>
int configure_level() {
    return (int)cudaFuncSetAttribute(atrous_kernel<S, kMode, kMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Tile<S, kMode>::SMEM);
}

}  // namespace

#define RMD_CAT2(a, b) a##b
#define RMD_CAT(a, b) RMD_CAT2(a, b)

int RMD_CAT(atrous_tile_configure_v, RMD_VARIANT)() {
    int rc = configure_level<1>(); if (rc) return rc;
    rc = configure_level<2>(); if (rc) return rc;
    rc = configure_level<4>(); if (rc) return rc;
    rc = configure_level<8>(); if (rc) return rc;
    return configure_level<16>();
}

int RMD_CAT(launch_atrous_tile_v, RMD_VARIANT)(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s,
                                               bool pdl) {
    switch (level) {
        case 0: return launch_level<1>(a, maps, s, pdl);
        case 1: return launch_level<2>(a, maps, s, pdl);
        case 2: return launch_level<4>(a, maps, s, pdl);
        case 3: return launch_level<8>(a, maps, s, pdl);
        case 4: return launch_level<16>(a, maps, s, pdl);
        default: return RMD_E_PARAM;
    }
}

}  // namespace rmd
