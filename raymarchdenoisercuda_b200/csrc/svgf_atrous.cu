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
    static constexpr int C4_BYTES = TW * TH * 16;
    static constexpr int V_BYTES = TW * TH * 4;
    static constexpr int OFF_C4 = 0;
    static constexpr int OFF_G4 = align128(C4_BYTES);
    static constexpr int OFF_V = OFF_G4 + align128(C4_BYTES);
    static constexpr int OFF_BAR = OFF_V + align128(V_BYTES);
    static constexpr int SMEM = OFF_BAR + 16 + 128;  // + slack to align the dynamic base to 128 B
    static constexpr uint32_t TX_BYTES = 2u * C4_BYTES + V_BYTES;
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
    const float d = fmaxf(fmaf(c.nz, g.z, fmaf(c.ny, g.y, c.nx * g.x)), 0.0f);
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
__global__ void __launch_bounds__(kAtrousWT* kAtrousTR, 2)
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

    // ---- stage the tile -------------------------------------------------------------
    if (a.use_tma) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();  // the barrier must be initialised before any thread polls it
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, T::TX_BYTES);
            tma_load_4d(smem + T::OFF_C4, &maps.c4, bar, 0, x0 - T::HX, phase, k0 - 2);
            tma_load_4d(smem + T::OFF_G4, &maps.g4, bar, 0, x0 - T::HX, phase, k0 - 2);
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
            wC4[i] = c;
            wG4[i] = g;
            wV[i] = v;
        }
    }

    // ---- per-output centre terms that do not come from the tile (global, L1/L2) ----
    // 3x3 Gaussian prefilter of the variance at the centre (dense neighbours: other
    // row phases, so not in the tile) and the depth slope.  Issued before the tile
    // wait so their latency overlaps the TMA.
    const int x = x0 + tx;
    const int xc = min(x, W - 1), xm = max(xc - 1, 0), xp = min(xc + 1, W - 1);
    float vbar[kAtrousOPT], dzv[kAtrousOPT];
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = min(phase + S * (k0 + kAtrousOPT * tr + j), H - 1);
        const int ym = max(y - 1, 0), yp = min(y + 1, H - 1);
        const float* r0 = a.in_v + (size_t)ym * Wp;
        const float* r1 = a.in_v + (size_t)y * Wp;
        const float* r2 = a.in_v + (size_t)yp * Wp;
        const float top = __ldg(r0 + xm) + 2.0f * __ldg(r0 + xc) + __ldg(r0 + xp);
        const float mid = __ldg(r1 + xm) + 2.0f * __ldg(r1 + xc) + __ldg(r1 + xp);
        const float bot = __ldg(r2 + xm) + 2.0f * __ldg(r2 + xc) + __ldg(r2 + xp);
        vbar[j] = (top + 2.0f * mid + bot) * (1.0f / 16.0f);
        dzv[j] = __ldg(a.dz + (size_t)y * Wp + xc);
    }

    if (a.use_tma) {
        mbar_wait(bar, 0);
    } else {
        __syncthreads();
    }

    // ---- centre set-up ---------------------------------------------------------------
    const float kLog2e = 1.4426950408889634f;
    Centre ctr[kAtrousOPT];
    Acc acc[kAtrousOPT];
    float4 cC[kAtrousOPT];
    float cV[kAtrousOPT];
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int row = kAtrousOPT * tr + j + 2, col = tx + T::HX;
        const float4 c = sC4[row * T::TW + col];
        const float4 g = sG4[row * T::TW + col];
        const float v = sV[row * T::TW + col];
        cC[j] = c;
        cV[j] = v;
        ctr[j].nx = g.x; ctr[j].ny = g.y; ctr[j].nz = g.z; ctr[j].z = g.w; ctr[j].L = c.w;
        const float phi_l = fmaf(a.sigma_l, sqrtf(fmaxf(vbar[j], 0.0f)), 1e-4f);
        ctr[j].il = kLog2e * fast_rcp(phi_l);
        const float zs = a.sigma_z * fmaxf(dzv[j], 1e-8f) * (float)S;
        ctr[j].iz[0] = kLog2e * fast_rcp(fmaf(zs, 1.0f, 1e-6f));
        ctr[j].iz[1] = kLog2e * fast_rcp(fmaf(zs, 1.4142135623730951f, 1e-6f));
        ctr[j].iz[2] = kLog2e * fast_rcp(fmaf(zs, 2.0f, 1e-6f));
        ctr[j].iz[3] = kLog2e * fast_rcp(fmaf(zs, 2.23606797749979f, 1e-6f));
        ctr[j].iz[4] = kLog2e * fast_rcp(fmaf(zs, 2.8284271247461903f, 1e-6f));
        const float h0 = 0.140625f;  // (3/8)^2
        acc[j].w = h0;
        acc[j].r = h0 * c.x; acc[j].g = h0 * c.y; acc[j].b = h0 * c.z;
        acc[j].v = h0 * h0 * v;
    }

    // ---- 100 taps from 40 staged texels ---------------------------------------------
    const float sigma_n = a.sigma_n;
#pragma unroll
    for (int jr = 0; jr < kAtrousOPT + 4; ++jr) {
        const int row = kAtrousOPT * tr + jr;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int idx = row * T::TW + tx + (T::HX - 2 * S) + c * S;
            const float4 q = sC4[idx];
            const float4 g = sG4[idx];
            const float v = sV[idx];
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                const int dy = jr - 2 - j;
                if (dy < -2 || dy > 2) continue;
                if (dy == 0 && c == 2) continue;  // centre tap, already accumulated with w = 1
                const int adx = c < 2 ? 2 - c : c - 2, ady = dy < 0 ? -dy : dy;
                // adx/ady are compile-time after unrolling; dispatch to the constexpr tap
                if (adx == 0 && ady == 1) tap<0, 1>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 0 && ady == 2) tap<0, 2>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 1 && ady == 0) tap<1, 0>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 1 && ady == 1) tap<1, 1>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 1 && ady == 2) tap<1, 2>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 2 && ady == 0) tap<2, 0>(acc[j], ctr[j], q, g, v, sigma_n);
                else if (adx == 2 && ady == 1) tap<2, 1>(acc[j], ctr[j], q, g, v, sigma_n);
                else tap<2, 2>(acc[j], ctr[j], q, g, v, sigma_n);
            }
        }
    }

    // ---- epilogue ---------------------------------------------------------------------
    if (x >= W) return;
#pragma unroll
    for (int j = 0; j < kAtrousOPT; ++j) {
        const int y = phase + S * (k0 + kAtrousOPT * tr + j);
        if (y >= H) continue;
        const float inv = fast_rcp(acc[j].w);
        float r = acc[j].r * inv, g = acc[j].g * inv, b = acc[j].b * inv, v = acc[j].v * inv * inv;
        const bool sky = ctr[j].z == 0.0f;
        if (sky) { r = cC[j].x; g = cC[j].y; b = cC[j].z; v = cV[j]; }
        if (a.out_c4) {
            const size_t p = (size_t)y * Wp + x;
            a.out_c4[p] = make_float4(r, g, b, sky ? cC[j].w : luminance(r, g, b));
            a.out_v[p] = v;
        }
        if (a.final_out) {
            const size_t p = (size_t)y * W + x;  // caller planes: pitch W
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

}  // namespace

int atrous_configure() {
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<1>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<2>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<4>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<8>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<16>::SMEM));
    return 0;
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
