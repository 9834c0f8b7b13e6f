// box_filter.cu — the reference's legacy denoise path: unweighted (2r+1)^2 box mean
// on RGBA8 planes.  Replaces the caller-side launches of
//   filterKernelBaseline  reference src/filter.cu:13-58   (call site src/test.cu:73-75)
//   filterKernelTiled     reference src/filter.cu:87-158  (call site src/test.cu:85-87)
// and the cooperative halo copy `cacheTile` (src/filter.cu:60-85).
//
// Semantics (bit-exact, checked against oracle/oracle_box.c and the reference's own
// source built for CPU and for sm_100a, oracle/ref_build/):
//   out = (uchar) ( float(sum of in-image taps) / float(count of in-image taps) )
// The reference accumulates uchar values in fp32 (exact: sums < 2^24) and divides
// with IEEE division (helper_math `float3 /= float`), so integer sums followed by
// one __fdiv_rn and a truncating cast reproduce it bit for bit.
//
// B200 design: the path moves 8 B/px (4 read + 4 written) and is HBM-bound if the
// arithmetic is kept below ~100 instructions per pixel, so the 2-D window is
// evaluated separably from shared memory: a CTA stages a (64+2r) x (32+2r) tile,
// builds horizontal window sums with two channels packed per 32-bit word
// (0x00FF00FF SIMD-within-register lanes, 16-bit fields hold (2r+1)*255 <= 16575),
// then sums those vertically.  Out-of-image texels are staged as 0 and the tap
// count is the product of the clipped window extents.  That generic kernel serves any radius and any
// alignment.  The radii the reference's harness uses (1..4; src/test.cu runs radius 2) on 16-byte aligned planes
// with W % 4 == 0 take `box_strip_kernel<R>` instead, which keeps everything in registers: a warp owns a
// 120-pixel wide column strip, every lane loads ONE 16-byte quad of texels per row (lanes 0 and 31 only feed the
// halo), gets its neighbours' texels by shuffle, forms the four horizontal window sums by sliding, and carries the
// vertical window as a running sum over a register ring of the last 2R+1 rows — ~25 instructions per pixel,
// no shared memory, no barrier, rows prefetched one ring ahead.
#include <stdlib.h>

#include "common.cuh"

namespace rmd {
namespace {

constexpr int kBoxTW = 64, kBoxTH = 32, kBoxThreads = 256;
constexpr int kBoxBX = 64, kBoxBY = kBoxThreads / kBoxBX;  // thread block 64 x 4: no div/mod in the loops

// `magic` = ceil(2^32 / (2r+1)^2) when floor(n / count) == umulhi(n, magic) for every reachable sum
// (n * count < 2^32, i.e. radius <= 31), else 0.  Used by tiles whose every tap is inside the image
// (count is then the same for every pixel); border tiles divide per pixel like the reference.
__global__ void __launch_bounds__(kBoxThreads) box_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                          int W, int H, int r, int replicate_r, uint32_t magic) {
    extern __shared__ uint32_t sm[];
    const int tw = kBoxTW + 2 * r, th = kBoxTH + 2 * r;
    uint32_t* tile = sm;                                              // tw * th raw texels
    uint2* hs = reinterpret_cast<uint2*>(sm + ((tw * th + 1) & ~1));  // th * kBoxTW horizontal sums {RB, GA}
    const int x0 = blockIdx.x * kBoxTW, y0 = blockIdx.y * kBoxTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // ---- stage the tile, out-of-image texels as 0 ----
    for (int row = ty; row < th; row += kBoxBY) {
        const int gy = y0 - r + row;
        const bool rowin = gy >= 0 && gy < H;
        const uint32_t* src = in + (size_t)(rowin ? gy : 0) * W;
        for (int col = tx; col < tw; col += kBoxBX) {
            const int gx = x0 - r + col;
            tile[row * tw + col] = (rowin && gx >= 0 && gx < W) ? __ldg(src + gx) : 0u;
        }
    }
    __syncthreads();
    // ---- horizontal window sums, two channels per 32-bit word ----
    for (int row = ty; row < th; row += kBoxBY) {
        const uint32_t* p = tile + row * tw + tx;
        uint32_t rb = 0u, ga = 0u;
        for (int d = 0; d <= 2 * r; ++d) {
            const uint32_t v = p[d];
            rb += v & 0x00FF00FFu;
            ga += (v >> 8) & 0x00FF00FFu;
        }
        hs[row * kBoxTW + tx] = make_uint2(rb, ga);
    }
    __syncthreads();
    // ---- vertical sums + division ----
    const int x = x0 + tx;
    const bool interior = magic != 0u && x0 - r >= 0 && x0 + kBoxTW + r <= W && y0 - r >= 0 && y0 + kBoxTH + r <= H;
    const int cx = min(x + r, W - 1) - max(x - r, 0) + 1;
    for (int row = ty; row < kBoxTH; row += kBoxBY) {
        const int y = y0 + row;
        if (x >= W || y >= H) continue;
        uint32_t sr = 0u, sg = 0u, sb = 0u;
        const uint2* p = hs + row * kBoxTW + tx;
        for (int d = 0; d <= 2 * r; ++d) {
            const uint2 h = p[d * kBoxTW];
            sr += h.x & 0xFFFFu;
            sb += h.x >> 16;
            sg += h.y & 0xFFFFu;
        }
        uint32_t R, G, B;
        if (interior) {
            R = __umulhi(sr, magic);
            G = __umulhi(sg, magic);
            B = __umulhi(sb, magic);
        } else {
            // the reference's arithmetic verbatim: fp32 sum / fp32 count, truncated (src/filter.cu:48-53)
            const int cy = min(y + r, H - 1) - max(y - r, 0) + 1;
            const float norm = (float)(cx * cy);
            R = (uint32_t)(unsigned char)__fdiv_rn((float)sr, norm);
            G = (uint32_t)(unsigned char)__fdiv_rn((float)sg, norm);
            B = (uint32_t)(unsigned char)__fdiv_rn((float)sb, norm);
        }
        if (replicate_r) {  // filterKernelBaseline writes acum.x to all three channels (src/filter.cu:51-53)
            G = R; B = R;
        }
        out[(size_t)y * W + x] = R | (G << 8) | (B << 16);  // .w = 0
    }
}

// ---- register strip kernel (radius 1..4, W % 4 == 0, 16-byte aligned planes) -------------------------------
#ifndef RMD_STRIP_MINB
#define RMD_STRIP_MINB 4
#endif
constexpr int kStripWarps = 4;        // warps per CTA, side by side in x
constexpr int kStripQuads = 30;       // output quads (4 px) per warp; lanes 0 and 31 are halo only

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Clipped windows (the frame's outermost R rows and columns): the reference divides the fp32 sum by the fp32
// in-image tap count and truncates (src/filter.cu:48-53).  Both are integers and a non-integer quotient is at
// least 1/count away from the next integer — far more than fp32 rounding can move it — so the result IS the
// integer quotient, taken here by multiplication with ceil(2^32 / count) (exact while 255 * count^2 < 2^32;
// tests/test_oracle_box.py checks every reachable sum).  A per-count table instead of twelve IEEE divisions per
// quad: the warps on the left/right frame edge take this path on every row and used to set the kernel's tail.
constexpr int kMagicCounts = 82;  // counts 1 .. (2*4+1)^2
struct MagicTable { uint32_t v[kMagicCounts]; };
constexpr MagicTable make_magic_table() {
    MagicTable t{};
    for (int c = 2; c < kMagicCounts; ++c) t.v[c] = (uint32_t)(((1ull << 32) + c - 1) / c);
    return t;  // v[0], v[1] = 0: "no division"
}
__constant__ MagicTable c_magic = make_magic_table();

__device__ __forceinline__ uint32_t pack_rgb(uint32_t r8, uint32_t g8, uint32_t b8) {
    return prmt(prmt(r8, g8, 0x4440u), b8, 0x7410u);  // {r, g, b, 0}
}

template <int R, bool REP>
__global__ void __launch_bounds__(kStripWarps * 32, (R <= 2 ? RMD_STRIP_MINB : 2)) box_strip_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                    int W4, int H, int strip, uint32_t magic) {
    constexpr int K = 2 * R + 1;
    const int lane = threadIdx.x & 31;
    const int wx = blockIdx.x * kStripWarps + (threadIdx.x >> 5);
    if (wx * kStripQuads >= W4) return;  // warp-uniform
    const int q = wx * kStripQuads - 1 + lane;  // this lane's quad column
    const int ys = blockIdx.y * strip, ye = min(ys + strip, H);
    const int n = ye - ys + 2 * R;              // rows streamed: the strip plus R above and below
    const bool col_in = q >= 0 && q < W4;
    const bool writes = lane >= 1 && lane <= kStripQuads && q < W4;
    const int W = W4 * 4, px0 = q * 4;
    const bool cols_full = px0 - R >= 0 && px0 + 3 + R <= W - 1;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int i_lo = max(0, R - ys), i_hi = min(n, H + R - ys);  // stream rows inside the frame
    const uint4* src = in + (ptrdiff_t)(ys - R) * W4 + q;        // stream row 0 (dereferenced inside the frame only)
    uint4* dst = out + (size_t)ys * W4 + q;                      // next output row
    uint32_t ring_rb[K][4], ring_g[K][4], v_rb[4], v_g[4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) ring_rb[k][j] = ring_g[k][j] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) v_rb[j] = v_g[j] = 0u;
    uint4 cur[K], nxt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        cur[k] = (col_in && k >= i_lo && k < i_hi) ? __ldg(src) : zero;
        src += W4;
    }
    for (int base = 0; base < n; base += K) {
#pragma unroll
        for (int k = 0; k < K; ++k) {  // one ring ahead
            const int i = base + K + k;
            nxt[k] = (col_in && i >= i_lo && i < i_hi) ? __ldg(src) : zero;
            src += W4;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = base + k;
            if (i < n) {  // warp-uniform
                // texels x-R .. x+3+R of this row: own quad + R words of each neighbour lane
                const uint4 c = cur[k];
                uint32_t t[4 + 2 * R];
                const uint32_t own[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    t[j] = __shfl_up_sync(0xffffffffu, own[4 - R + j], 1);
                    t[4 + R + j] = __shfl_down_sync(0xffffffffu, own[j], 1);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) t[R + j] = own[j];
                uint32_t rb[4 + 2 * R], g[4 + 2 * R];
#pragma unroll
                for (int j = 0; j < 4 + 2 * R; ++j) {
                    rb[j] = prmt(t[j], 0u, 0x4240u);  // {r, 0, b, 0}: two 16-bit fields
                    g[j] = prmt(t[j], 0u, 0x4441u);   // {g, 0, 0, 0}
                }
                // the row leaving the vertical window goes first (fields never borrow: it is part of the sum), so the
                // new horizontal sums can be formed in its ring slot
#pragma unroll
                for (int j = 0; j < 4; ++j) { v_rb[j] -= ring_rb[k][j]; v_g[j] -= ring_g[k][j]; }
                uint32_t a_rb = rb[0], a_g = g[0];
#pragma unroll
                for (int j = 1; j < K; ++j) { a_rb += rb[j]; a_g += g[j]; }
                ring_rb[k][0] = a_rb; ring_g[k][0] = a_g;
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    ring_rb[k][j] = ring_rb[k][j - 1] + rb[j + 2 * R] - rb[j - 1];
                    ring_g[k][j] = ring_g[k][j - 1] + g[j + 2 * R] - g[j - 1];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { v_rb[j] += ring_rb[k][j]; v_g[j] += ring_g[k][j]; }
                if (i >= 2 * R) {  // warp-uniform: this row completes the window of output row ys + i - 2R
                    if (writes) {
                        const int y = ys + i - 2 * R;
                        uint32_t o[4];
                        if (cols_full && y >= R && y + R <= H - 1) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t r8 = __umulhi(v_rb[j] & 0xFFFFu, magic);
                                o[j] = REP ? r8 * 0x010101u : pack_rgb(r8, __umulhi(v_g[j], magic), __umulhi(v_rb[j] >> 16, magic));
                            }
                        } else {
                            const int cy = min(y + R, H - 1) - max(y - R, 0) + 1;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int x = px0 + j;
                                const uint32_t m = c_magic.v[(min(x + R, W - 1) - max(x - R, 0) + 1) * cy];
                                const uint32_t sr = v_rb[j] & 0xFFFFu, sb = v_rb[j] >> 16, sg = v_g[j];
                                const uint32_t r8 = m ? __umulhi(sr, m) : sr;
                                o[j] = REP ? r8 * 0x010101u : pack_rgb(r8, m ? __umulhi(sg, m) : sg, m ? __umulhi(sb, m) : sb);
                            }
                        }
                        *dst = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    dst += W4;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) cur[k] = nxt[k];
    }
}

template <int R>
void launch_strip(dim3 grid, cudaStream_t s, const uint4* in, uint4* out, int W4, int H, int strip, int replicate_r, uint32_t magic) {
    if (replicate_r) box_strip_kernel<R, true><<<grid, kStripWarps * 32, 0, s>>>(in, out, W4, H, strip, magic);
    else box_strip_kernel<R, false><<<grid, kStripWarps * 32, 0, s>>>(in, out, W4, H, strip, magic);
}

int strip_rows(int W4, int H) {
    if (const char* e = getenv("RMD_BOX_STRIP")) { const int v = atoi(e); if (v > 0) return v; }
    // tallest strip that still gives every SM two rounds of 16 warps; 16 rows at least (halo re-reads: 2R / strip)
    const long long xw = (W4 + kStripQuads - 1) / kStripQuads;
    for (int s = 128; s > 16; s >>= 1)
        if (xw * ((H + s - 1) / s) >= 148LL * 32) return s;
    return 16;
}

size_t box_smem_bytes(int r) {
    const int tw = kBoxTW + 2 * r, th = kBoxTH + 2 * r;
    return (size_t)((tw * th + 1) & ~1) * 4 + (size_t)th * kBoxTW * 8;
}

int box_filter(const RmdGBuffer* f, const RmdFilterParams* p, int replicate_r, cudaStream_t s) {
    if (!f || !p) return RMD_E_NULL;
    if (f->width <= 0 || f->height <= 0 || (long long)f->width * f->height > 0x7FFFFFFFLL) return RMD_E_SHAPE;
    if (p->type != RMD_FILTER_AVERAGE)  // GAUSSIAN / CROSS were dispatched before; WAVELET needs rmd_svgf_*
        return p->type == RMD_FILTER_WAVELET ? RMD_E_UNSUPPORTED : RMD_E_PARAM;
    if (p->radius < 0 || p->radius > RMD_BOX_MAX_RADIUS || p->depth < 1) return RMD_E_PARAM;
    if (!f->render || !f->denoised) return RMD_E_NULL;
    if (p->depth > 1 && (!f->buffer[0] || !f->buffer[1])) return RMD_E_NULL;
    if (((uintptr_t)f->render | (uintptr_t)f->denoised | (uintptr_t)f->buffer[0] | (uintptr_t)f->buffer[1]) & 3u)
        return RMD_E_ALIGN;
    const size_t smem = box_smem_bytes(p->radius);
    // function attributes are per device: set the opt-in whenever a launch needs it (large radii only)
    if (smem > 48 * 1024)
        RMD_CUDA_TRY(cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((f->width + kBoxTW - 1) / kBoxTW, (f->height + kBoxTH - 1) / kBoxTH);
    dim3 block(kBoxBX, kBoxBY);
    const uint64_t cnt = (uint64_t)(2 * p->radius + 1) * (2 * p->radius + 1);
    // exact while 255 * cnt * cnt < 2^32 (radius <= 31); 0 selects the per-pixel division everywhere
    const uint32_t magic = (255ull * cnt * cnt < (1ull << 32)) ? (uint32_t)(((1ull << 32) + cnt - 1) / cnt) : 0u;
    const bool strip_ok = p->radius >= 1 && p->radius <= 4 && f->width % 4 == 0 && magic != 0u && !getenv("RMD_BOX_GENERIC") &&
                          !(((uintptr_t)f->render | (uintptr_t)f->denoised | (uintptr_t)f->buffer[0] | (uintptr_t)f->buffer[1]) & 15u);
    for (int level = 0; level < p->depth; ++level) {
        // ping-pong of the reference (src/filter.cu:24-25)
        const void* in = level == 0 ? f->render : f->buffer[level % 2];
        void* out = level == p->depth - 1 ? f->denoised : f->buffer[(level + 1) % 2];
        if (strip_ok) {
            const int W4 = f->width / 4, strip = strip_rows(W4, f->height);
            const int xw = (W4 + kStripQuads - 1) / kStripQuads;
            dim3 sgrid((xw + kStripWarps - 1) / kStripWarps, (f->height + strip - 1) / strip);
            const uint4* i4 = (const uint4*)in;
            uint4* o4 = (uint4*)out;
            switch (p->radius) {
                case 1: launch_strip<1>(sgrid, s, i4, o4, W4, f->height, strip, replicate_r, magic); break;
                case 2: launch_strip<2>(sgrid, s, i4, o4, W4, f->height, strip, replicate_r, magic); break;
                case 3: launch_strip<3>(sgrid, s, i4, o4, W4, f->height, strip, replicate_r, magic); break;
                default: launch_strip<4>(sgrid, s, i4, o4, W4, f->height, strip, replicate_r, magic); break;
            }
            continue;
        }
        box_kernel<<<grid, block, smem, s>>>((const uint32_t*)in, (uint32_t*)out, f->width, f->height, p->radius,
                                             replicate_r, magic);
    }
    return (int)cudaGetLastError();
}

}  // namespace
}  // namespace rmd

namespace rmd {
int weighted_filter(const RmdGBuffer* f, const RmdFilterParams* p, cudaStream_t s);  // weighted_filter.cu
}

// FilterParams::type selects the arithmetic (reference include/filter.cuh:12): AVERAGE = the reference's two box
// kernels; GAUSSIAN / CROSS = the weighted filters of weighted_filter.cu (identical through either entry point: the
// R-replication quirk of filterKernelBaseline belongs to its AVERAGE code, src/filter.cu:51-53); WAVELET = the SVGF
// path, which needs a per-sequence context (rmd_svgf_frame_gbuffer).
extern "C" int rmd_filter_baseline(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream) {
    if (params && (params->type == RMD_FILTER_GAUSSIAN || params->type == RMD_FILTER_CROSS))
        return rmd::weighted_filter(frame, params, (cudaStream_t)stream);
    return rmd::box_filter(frame, params, 1, (cudaStream_t)stream);
}
extern "C" int rmd_filter_tiled(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream) {
    if (params && (params->type == RMD_FILTER_GAUSSIAN || params->type == RMD_FILTER_CROSS))
        return rmd::weighted_filter(frame, params, (cudaStream_t)stream);
    return rmd::box_filter(frame, params, 0, (cudaStream_t)stream);
}
