// svgf_variance.cu — pass 2 of the SVGF path: 7x7 cross-bilateral estimate of colour,
// moments and variance for pixels whose history is shorter than `short_history`
// (DESIGN.md spec S3; oracle/oracle_svgf.c:pass_variance).  No reference counterpart.
//
// The pass is sparse in steady state (only disoccluded regions qualify), so it is
// driven by the compact list of 32x8 tiles with short-history pixels that the
// temporal kernel appended to: one persistent CTA per SM slot walks the list with a
// grid stride (round 1 launched one CTA per tile and let the un-flagged ones exit;
// a third of the kernel's warp-time was spent waiting for that one flag load).
// For each tile the CTA compacts the qualifying pixels into a shared-memory list and
// gives ONE pixel to each of the first `count` threads, so the 49-tap loop runs on
// dense warps instead of on the scattered lanes that happen to qualify.
// Jacobi semantics without a second pass: the temporal kernel also stores the
// colour of every short-history pixel in a side plane that nobody modifies; a tap
// reads the side plane when the neighbour is itself a short-history pixel (which
// this pass may already have overwritten) and the temporal plane otherwise, and
// the result is written in place.
//
// Roofline: HBM.  Worst case (every pixel qualifies, e.g. the first frames of a
// sequence): reads side colour 16 + moments 8 + guide 16 + slope 4 + histlen 1,
// writes colour 16 + variance 4 => 65 B/px (+16 B/px the temporal pass spent on
// the side plane).  Steady state: 4 B per 32x8 tile plus the disoccluded pixels.
#include "svgf.cuh"

namespace rmd {
namespace {

constexpr int kVarHalo = 3;
constexpr int kVarTW = kTemporalBx + 2 * kVarHalo, kVarTH = kTemporalBy + 2 * kVarHalo;

// index of the squared tap distance d2 = dx^2 + dy^2 (7x7 window: 9 distinct non-zero values)
__device__ __forceinline__ constexpr int var_dist_class(int d2) {
    return d2 == 1 ? 0 : d2 == 2 ? 1 : d2 == 4 ? 2 : d2 == 5 ? 3 : d2 == 8 ? 4 : d2 == 9 ? 5 : d2 == 10 ? 6 : d2 == 13 ? 7 : 8;
}

constexpr int kVarTilePx = kTemporalBx * kTemporalBy;  // pixels of a flagged tile (32 x 8)

// NT threads per CTA work on one 32x8 tile at a time (NT = 256: 4 CTAs per SM, the default; NT = 128: 8 CTAs per SM,
// twice as many tiles in flight — measured equal, see launch_variance).
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) variance_kernel(const VarianceArgs a) {
    __shared__ int s_count;
    __shared__ unsigned short s_list[kVarTilePx];
    __shared__ float4 sG[kVarTW * kVarTH];   // guide of the tile + 3-texel halo (0 outside the image => weight 0)
    __shared__ float4 sC[kVarTW * kVarTH];   // Jacobi colour: side copy for short-history texels, temporal output otherwise
    __shared__ float2 sM[kVarTW * kVarTH];
    __shared__ float sDZ[kVarTilePx];        // depth slope of the tile's own pixels
    __shared__ uint8_t sN[kVarTW * kVarTH];  // history length (compaction and the 4/N factor read it again)
    pdl_wait();  // the tile list is written by the temporal kernel
    pdl_launch_dependents();
    const int W = a.W, H = a.H, Wp = a.Wp;
    const int tid = threadIdx.x;
    const int short_hist = a.k.short_hist;
    const unsigned ntiles = min(*a.tile_count, a.tile_capacity);
    if (blockIdx.x == 0 && tid == 0) *a.next_count = 0u;  // nobody reads or appends to that counter during this kernel
    // persistent CTAs over the compact list of flagged tiles (strided: every CTA gets the same number +- 1)
    for (unsigned ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const uint32_t entry = a.tile_list[ti];
    const int x0 = (int)(entry >> 16) * kTemporalBx, y0 = (int)(entry & 0xFFFFu);
    if (tid == 0) s_count = 0;
    // ---- stage the neighbourhood once (coalesced rows); all planes in ONE round trip per batch: the side colour is
    //      loaded speculatively and selected in registers ----
#pragma unroll 2
    for (int i = tid; i < kVarTW * kVarTH; i += NT) {
        const int ty = i / kVarTW, tx = i - ty * kVarTW;
        const int gx = x0 - kVarHalo + tx, gy = y0 - kVarHalo + ty;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), c = g;
        float2 m = make_float2(0.f, 0.f);
        int nq = 255;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t q = (size_t)gy * Wp + gx;
            g = a.g4[q];
            nq = a.n[q];
            const float4 cs = a.side_c4[q], ct = a.c4[q];
            m = a.m[q];
            c = (nq < short_hist && g.w != 0.0f) ? cs : ct;
        }
        sG[i] = g; sC[i] = c; sM[i] = m; sN[i] = (uint8_t)nq;
    }
#pragma unroll
    for (int i = tid; i < kVarTilePx; i += NT) {
        const int x = x0 + (i & (kTemporalBx - 1)), y = y0 + i / kTemporalBx;
        sDZ[i] = (x < W && y < H) ? a.dz[(size_t)y * Wp + x] : 0.0f;
    }
    __syncthreads();
    // compaction: which pixels of the tile take the spatial estimate?  (every warp covers whole tile rows)
#pragma unroll
    for (int i = tid; i < kVarTilePx; i += NT) {
        const int lx = i & (kTemporalBx - 1), ly = i / kTemporalBx;
        const int x = x0 + lx, y = y0 + ly;
        bool need = false;
        if (x < W && y >= a.row_begin && y < a.row_end) {
            const int ci = (ly + kVarHalo) * kVarTW + lx + kVarHalo;
            need = sN[ci] < short_hist && sG[ci].w != 0.0f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        int base = 0;
        if (lx == 0 && m) base = atomicAdd(&s_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (need) s_list[base + __popc(m & ((1u << lx) - 1u))] = (unsigned short)i;
    }
    __syncthreads();
    const int count = s_count;
    for (int li = tid; li < count; li += NT) {
    const int id = s_list[li];
    const int lx = id & (kTemporalBx - 1), ly = id / kTemporalBx;
    const int x = x0 + lx, y = y0 + ly;
    const size_t p = (size_t)y * Wp + x;
    const int ci = (ly + kVarHalo) * kVarTW + lx + kVarHalo;
    const float4 gp = sG[ci];
    const float4 cp = sC[ci];
    const float2 mp = sM[ci];
    const int Nn = sN[ci];
    const float kLog2e = 1.4426950408889634f;
    const float zs = a.k.sigma_z * fmaxf(sDZ[id], 1e-8f);
    const float il = kLog2e / a.k.lscale;
    const float sigma_n = a.k.sigma_n;
    float iz[9];
    {
        const float dist[9] = {1.0f, 1.4142135623730951f, 2.0f, 2.23606797749979f, 2.8284271247461903f, 3.0f,
                               3.1622776601683795f, 3.605551275463989f, 4.242640687119285f};
#pragma unroll
        for (int k = 0; k < 9; ++k) iz[k] = kLog2e * fast_rcp(fmaf(zs, dist[k], 1e-6f));
    }
    // Var = M2 - M1^2 cancels catastrophically in fp32 once the demodulated luminance is large (albedo at the
    // floor: L ~ 1e3, M2 ~ 1e6; the oracle accumulates in double).  The moment sums are therefore taken of the
    // DIFFERENCES to the centre's moments (exact when the neighbourhood is coherent, relative accuracy otherwise)
    // and combined in FP64 once per pixel.
    float sw = 1.0f, sr = cp.x, sg = cp.y, sb = cp.z;
    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
    for (int dx = -3; dx <= 3; ++dx) {
#pragma unroll
        for (int dy = -3; dy <= 3; ++dy) {
            if (dx == 0 && dy == 0) continue;
            const int qi = ci + dy * kVarTW + dx;
            const float4 gq = sG[qi];
            const float4 cq = sC[qi];
            const float2 mq = sM[qi];
            const float d = __saturatef(fmaf(gp.z, gq.z, fmaf(gp.y, gq.y, gp.x * gq.x)));
            float e = sigma_n * fast_lg2(d);  // out-of-image / sky / back-facing taps: d = 0 -> -inf -> w = 0
            e = fmaf(-fabsf(gp.w - gq.w), iz[var_dist_class(dx * dx + dy * dy)], e);
            e = fmaf(-fabsf(cp.w - cq.w), il, e);
            const float w = fast_ex2(e);
            sw += w;
            sr = fmaf(w, cq.x, sr); sg = fmaf(w, cq.y, sg); sb = fmaf(w, cq.z, sb);
            d0 = fmaf(w, mq.x - mp.x, d0); d1 = fmaf(w, mq.y - mp.y, d1);
        }
    }
    const float inv = 1.0f / fmaxf(sw, 1e-6f);
    const float r = sr * inv, g = sg * inv, b = sb * inv;
    // m0 = mp.x + D0, m1 = mp.y + D1  =>  m1 - m0^2 = (mp.y - mp.x^2) + D1 - 2 mp.x D0 - D0^2
    const double D0 = (double)(d0 * inv), D1 = (double)(d1 * inv), c0 = (double)mp.x;
    const double v = ((double)mp.y - c0 * c0) + D1 - 2.0 * c0 * D0 - D0 * D0;
    const float var = (float)(fmax(0.0, v) * (double)(4.0f / (float)Nn));
    a.patch_c4[p] = make_float4(r, g, b, luminance(r, g, b));
    a.patch_v[p] = var;
    }
    __syncthreads();  // the shared tile and list are reused by the next entry
    }
}

}  // namespace

int launch_variance(const VarianceArgs& a, cudaStream_t s, bool pdl) {
    static const int sms = [] {  // same for every B200 in the box; initialised once, thread-safe
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    static const int threads = [] {  // A/B switch: RMD_VAR_THREADS=128 runs 8 CTAs of 4 warps per SM (measured equal:
        const char* e = getenv("RMD_VAR_THREADS");  // 43.0 vs 44.2 us at 1080p, 72 vs 75 us at 4K — the tap loop is
        return e && atoi(e) == 128 ? 128 : 256;     // shared-memory bound, not staging bound, profiles/r2_notes.md)
    }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms * (1024 / threads));
    cfg.blockDim = dim3(threads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return threads == 256 ? (int)cudaLaunchKernelEx(&cfg, variance_kernel<256>, a)
                          : (int)cudaLaunchKernelEx(&cfg, variance_kernel<128>, a);
}

}  // namespace rmd
