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
// count is the product of the clipped window extents.
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

size_t box_smem_bytes(int r) {
    const int tw = kBoxTW + 2 * r, th = kBoxTH + 2 * r;
    return (size_t)((tw * th + 1) & ~1) * 4 + (size_t)th * kBoxTW * 8;
}

int box_filter(const RmdGBuffer* f, const RmdFilterParams* p, int replicate_r, cudaStream_t s) {
    if (!f || !p) return RMD_E_NULL;
    if (f->width <= 0 || f->height <= 0 || (long long)f->width * f->height > 0x7FFFFFFFLL) return RMD_E_SHAPE;
    if (p->type != RMD_FILTER_AVERAGE) return RMD_E_UNSUPPORTED;  // the reference reads no other type either
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
    for (int level = 0; level < p->depth; ++level) {
        // ping-pong of the reference (src/filter.cu:24-25)
        const void* in = level == 0 ? f->render : f->buffer[level % 2];
        void* out = level == p->depth - 1 ? f->denoised : f->buffer[(level + 1) % 2];
        box_kernel<<<grid, block, smem, s>>>((const uint32_t*)in, (uint32_t*)out, f->width, f->height, p->radius,
                                             replicate_r, magic);
    }
    return (int)cudaGetLastError();
}

}  // namespace
}  // namespace rmd

extern "C" int rmd_filter_baseline(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream) {
    return rmd::box_filter(frame, params, 1, (cudaStream_t)stream);
}
extern "C" int rmd_filter_tiled(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream) {
    return rmd::box_filter(frame, params, 0, (cudaStream_t)stream);
}
