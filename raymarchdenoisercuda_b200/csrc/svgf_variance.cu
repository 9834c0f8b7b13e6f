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
// dense warps instead of on the scattered lanes that happen to qualify.  A tile in
// which at least half of the pixels qualify (the first frames of a sequence, large
// disocclusions) is walked by position instead, two vertically adjacent pixels per
// thread: a staged texel then feeds two taps and the loop stops being bound by
// shared-memory bandwidth.  Both paths produce the same bits for a pixel.
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
constexpr int kVarDenseNT = kVarTilePx / 2;            // threads the position-mapped (dense) path uses: 2 pixels each

// per-pixel constants of the 49-tap estimate
struct VarCentre {
    float4 g;      // guide (normal, depth)
    float l;       // luminance of the Jacobi colour
    float2 m;      // moments
    float iz[9];   // log2(e) / (sigma_z * slope * distance + 1e-6) by distance class
};
struct VarAcc {
    float sw, sr, sg, sb, d0, d1;
};

__device__ __forceinline__ void var_centre(VarCentre& c, VarAcc& acc, const float4 gp, const float4 cp, const float2 mp,
                                           const float slope, const float sigma_z) {
    const float kLog2e = 1.4426950408889634f;
    const float dist[9] = {1.0f, 1.4142135623730951f, 2.0f, 2.23606797749979f, 2.8284271247461903f, 3.0f,
                           3.1622776601683795f, 3.605551275463989f, 4.242640687119285f};
    const float zs = sigma_z * fmaxf(slope, 1e-8f);
    c.g = gp; c.l = cp.w; c.m = mp;
#pragma unroll
    for (int k = 0; k < 9; ++k) c.iz[k] = kLog2e * fast_rcp(fmaf(zs, dist[k], 1e-6f));
    acc.sw = 1.0f; acc.sr = cp.x; acc.sg = cp.y; acc.sb = cp.z; acc.d0 = 0.0f; acc.d1 = 0.0f;
}

// One tap.  Both paths of the kernel accumulate a pixel's 48 taps through this function in the same order (dx outer,
// dy inner), so a pixel gets the same bits whichever path its tile takes (band mode tiles the frame differently).
// Var = M2 - M1^2 cancels catastrophically in fp32 once the demodulated luminance is large (albedo at the floor:
// L ~ 1e3, M2 ~ 1e6; the oracle accumulates in double).  The moment sums are therefore taken of the DIFFERENCES to
// the centre's moments (exact when the neighbourhood is coherent, relative accuracy otherwise) and combined in FP64
// once per pixel.
template <int DX, int DY>
__device__ __forceinline__ void var_tap(VarAcc& acc, const VarCentre& c, const float4 gq, const float4 cq, const float2 mq,
                                        const float sigma_n, const float il) {
    const float d = __saturatef(fmaf(c.g.z, gq.z, fmaf(c.g.y, gq.y, c.g.x * gq.x)));
    float e = sigma_n * fast_lg2(d);  // out-of-image / sky / back-facing taps: d = 0 -> -inf -> w = 0
    e = fmaf(-fabsf(c.g.w - gq.w), c.iz[var_dist_class(DX * DX + DY * DY)], e);
    e = fmaf(-fabsf(c.l - cq.w), il, e);
    const float w = fast_ex2(e);
    acc.sw += w;
    acc.sr = fmaf(w, cq.x, acc.sr); acc.sg = fmaf(w, cq.y, acc.sg); acc.sb = fmaf(w, cq.z, acc.sb);
    acc.d0 = fmaf(w, mq.x - c.m.x, acc.d0); acc.d1 = fmaf(w, mq.y - c.m.y, acc.d1);
}

__device__ __forceinline__ void var_finish(const VarianceArgs& a, const VarAcc& acc, const VarCentre& c, const int Nn, const size_t p) {
    const float inv = 1.0f / fmaxf(acc.sw, 1e-6f);
    const float r = acc.sr * inv, g = acc.sg * inv, b = acc.sb * inv;
    // m0 = mp.x + D0, m1 = mp.y + D1  =>  m1 - m0^2 = (mp.y - mp.x^2) + D1 - 2 mp.x D0 - D0^2
    const double D0 = (double)(acc.d0 * inv), D1 = (double)(acc.d1 * inv), c0 = (double)c.m.x;
    const double v = ((double)c.m.y - c0 * c0) + D1 - 2.0 * c0 * D0 - D0 * D0;
    const float var = (float)(fmax(0.0, v) * (double)(4.0f / (float)Nn));
    a.patch_c4[p] = make_float4(r, g, b, luminance(r, g, b));
    a.patch_v[p] = var;
}

// One column (fixed dx) of the 7x7 windows of two vertically adjacent pixels: 8 staged texels feed 14 taps
// (0.57 shared-memory texel loads per tap instead of 1: the one-pixel-per-thread loop moves 40 B per tap per lane and
// is shared-memory bound at 1.8x the SM's bandwidth, profiles/r2_notes.md).
template <int DX>
__device__ __forceinline__ void var_column2(VarAcc (&acc)[2], const VarCentre (&c)[2], const float4* sG, const float4* sC,
                                            const float2* sM, const int qb, const float sigma_n, const float il) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int qi = qb + j * kVarTW + DX;
        const float4 gq = sG[qi];
        const float4 cq = sC[qi];
        const float2 mq = sM[qi];
        // texel j is tap dy = j - 3 of the upper pixel and dy = j - 4 of the lower one
        if (j <= 6 && !(DX == 0 && j == 3)) {
            if (j == 0) var_tap<DX, -3>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 1) var_tap<DX, -2>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 2) var_tap<DX, -1>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 3) var_tap<DX, 0>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 4) var_tap<DX, 1>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 5) var_tap<DX, 2>(acc[0], c[0], gq, cq, mq, sigma_n, il);
            if (j == 6) var_tap<DX, 3>(acc[0], c[0], gq, cq, mq, sigma_n, il);
        }
        if (j >= 1 && !(DX == 0 && j == 4)) {
            if (j == 1) var_tap<DX, -3>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 2) var_tap<DX, -2>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 3) var_tap<DX, -1>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 4) var_tap<DX, 0>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 5) var_tap<DX, 1>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 6) var_tap<DX, 2>(acc[1], c[1], gq, cq, mq, sigma_n, il);
            if (j == 7) var_tap<DX, 3>(acc[1], c[1], gq, cq, mq, sigma_n, il);
        }
    }
}

// One column of the window of one pixel (the compacted path)
template <int DX>
__device__ __forceinline__ void var_column1(VarAcc& acc, const VarCentre& c, const float4* sG, const float4* sC,
                                            const float2* sM, const int ci, const float sigma_n, const float il) {
#define RMD_VAR_TAP1(DY)                                                                                   \
    if (!(DX == 0 && (DY) == 0)) {                                                                         \
        const int qi = ci + (DY)*kVarTW + DX;                                                              \
        var_tap<DX, (DY)>(acc, c, sG[qi], sC[qi], sM[qi], sigma_n, il);                                    \
    }
    RMD_VAR_TAP1(-3) RMD_VAR_TAP1(-2) RMD_VAR_TAP1(-1) RMD_VAR_TAP1(0) RMD_VAR_TAP1(1) RMD_VAR_TAP1(2) RMD_VAR_TAP1(3)
#undef RMD_VAR_TAP1
}

// NT threads per CTA work on one 32x8 tile at a time.  A tile with at least `dense_min` qualifying pixels is walked by
// position (thread = one column x two rows, kVarDenseNT threads), any other through the compacted list (one pixel per
// thread).  NT = 128 (default): every thread has work in the dense path; NT = 256: the round-1/2 shape.
template <int NT>
__global__ void __launch_bounds__(NT, NT == 128 ? 6 : 4) variance_kernel(const VarianceArgs a) {
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
    const float kLog2e = 1.4426950408889634f;
    const float il = kLog2e / a.k.lscale;
    const float sigma_n = a.k.sigma_n;
    const unsigned ntiles = min(*a.tile_count, a.tile_capacity);
    if (blockIdx.x == 0 && tid == 0) *a.next_count = 0u;  // nobody reads or appends to that counter during this kernel
    // persistent CTAs over the compact list of flagged tiles (strided: every CTA gets the same number +- 1)
    for (unsigned ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const uint32_t entry = a.tile_list[a.reverse ? ntiles - 1u - ti : ti];
        const int x0 = (int)(entry >> 16) * kTemporalBx, y0 = (int)(entry & 0xFFFFu);
        if (tid == 0) s_count = 0;
        // ---- stage the neighbourhood once (coalesced rows); all planes in ONE round trip per batch: the side colour
        //      is loaded speculatively and selected in registers ----
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
        if (count >= a.dense_min) {
            // ---- dense tile (the first frames of a sequence, large disocclusions): by position, two rows per thread ----
            for (int t = tid; t < kVarDenseNT; t += NT) {
                const int lx = t & (kTemporalBx - 1), ly = 2 * (t / kTemporalBx);
                const int ci = (ly + kVarHalo) * kVarTW + lx + kVarHalo;
                VarCentre c[2];
                VarAcc acc[2];
                var_centre(c[0], acc[0], sG[ci], sC[ci], sM[ci], sDZ[ly * kTemporalBx + lx], a.k.sigma_z);
                var_centre(c[1], acc[1], sG[ci + kVarTW], sC[ci + kVarTW], sM[ci + kVarTW], sDZ[(ly + 1) * kTemporalBx + lx], a.k.sigma_z);
                const int qb = ci - 3 * kVarTW;
                var_column2<-3>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<-2>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<-1>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<0>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<1>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<2>(acc, c, sG, sC, sM, qb, sigma_n, il);
                var_column2<3>(acc, c, sG, sC, sM, qb, sigma_n, il);
                const int x = x0 + lx;
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int y = y0 + ly + o;
                    const int co = ci + o * kVarTW;
                    const int Nn = sN[co];
                    if (x < W && y >= a.row_begin && y < a.row_end && Nn < short_hist && c[o].g.w != 0.0f)
                        var_finish(a, acc[o], c[o], Nn, (size_t)y * Wp + x);
                }
            }
        } else {
            // ---- sparse tile: ONE pixel of the compacted list per thread ----
            for (int li = tid; li < count; li += NT) {
                const int id = s_list[li];
                const int lx = id & (kTemporalBx - 1), ly = id / kTemporalBx;
                const int ci = (ly + kVarHalo) * kVarTW + lx + kVarHalo;
                VarCentre c;
                VarAcc acc;
                var_centre(c, acc, sG[ci], sC[ci], sM[ci], sDZ[id], a.k.sigma_z);
                var_column1<-3>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<-2>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<-1>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<0>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<1>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<2>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_column1<3>(acc, c, sG, sC, sM, ci, sigma_n, il);
                var_finish(a, acc, c, sN[ci], (size_t)(y0 + ly) * Wp + x0 + lx);
            }
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
    const int threads = a.threads == 256 ? 256 : 128;  // A/B switches of the context (RMD_VAR_THREADS, RMD_VAR_DENSE_MIN)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms * (threads == 128 ? 6 : 4));
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
