// svgf_atrous.cu — a-trous level dispatch (independent-tile kernel variants live in
// svgf_atrous_tile.cu, compiled once per variant), the persistent ring kernel kept as a measured
// alternative (RMD_ATROUS_RING=1), and the depth-0 remodulate kernel.  DESIGN.md spec S4-S5,
// checked against oracle/oracle_svgf.c:pass_atrous.
#include "svgf_atrous.cuh"

namespace rmd {
namespace {

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
                if (y >= a.row0 && y < a.row0 + a.rows) store_output_vals(a, acc[j], ctr[j], cC[j], cV[j], x, y);
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

// independent-tile kernel variants (svgf_atrous_tile.cu, one object per RMD_VARIANT)
#define RMD_DECL_VARIANT(n)            \
    int atrous_tile_configure_v##n();  \
    int atrous_tile_width_v##n(int level); \
    int launch_atrous_tile_v##n(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s, bool pdl); \
    int atrous_tile_cover_v##n(int level, const AtrousArgs& a, int* cover, int* nbx_out, int* tiles_with_work);
RMD_ATROUS_VARIANTS(RMD_DECL_VARIANT)
#undef RMD_DECL_VARIANT

int atrous_configure() {
    int dev = 0;
    RMD_CUDA_TRY(cudaGetDevice(&dev));
    RMD_CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<1>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<2>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<4>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<8>::SMEM));
    RMD_CUDA_TRY(cudaFuncSetAttribute(atrous_ring_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<16>::SMEM));
#define RMD_CFG_VARIANT(n) { const int rc = atrous_tile_configure_v##n(); if (rc) return rc; }
    RMD_ATROUS_VARIANTS(RMD_CFG_VARIANT)
#undef RMD_CFG_VARIANT
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

bool atrous_variant_exists(int variant) {
#define RMD_HAS_VARIANT(n) if (variant == n) return true;
    RMD_ATROUS_VARIANTS(RMD_HAS_VARIANT)
#undef RMD_HAS_VARIANT
    return false;
}

int atrous_variant_tile_width(int variant, int level) {
#define RMD_WT_VARIANT(n) if (variant == n) return atrous_tile_width_v##n(level);
    RMD_ATROUS_VARIANTS(RMD_WT_VARIANT)
#undef RMD_WT_VARIANT
    return kAtrousWT;
}

int launch_atrous(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s, int variant, bool pdl) {
#define RMD_RUN_VARIANT(n) if (variant == n) return launch_atrous_tile_v##n(level, a, maps, s, pdl);
    RMD_ATROUS_VARIANTS(RMD_RUN_VARIANT)
#undef RMD_RUN_VARIANT
    return RMD_E_PARAM;
}

int atrous_cover(int level, const AtrousArgs& a, int variant, int* cover, int* nbx_out, int* tiles_with_work) {
#define RMD_COVER_VARIANT(n) if (variant == n) return atrous_tile_cover_v##n(level, a, cover, nbx_out, tiles_with_work);
    RMD_ATROUS_VARIANTS(RMD_COVER_VARIANT)
#undef RMD_COVER_VARIANT
    return RMD_E_PARAM;
}

int launch_remodulate(const float4* c4, const float* v, const float4* g4, const uchar4* albedo, float4* out,
                      uchar4* out8, int W, int H, int Wp, float afloor, cudaStream_t s) {
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    remodulate_kernel<<<grid, block, 0, s>>>(c4, v, g4, albedo, out, out8, W, H, Wp, afloor);
    return (int)cudaGetLastError();
}

}  // namespace rmd
