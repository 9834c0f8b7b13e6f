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

__global__ void __launch_bounds__(kBoxThreads) box_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                          int W, int H, int r, int replicate_r) {
    extern __shared__ uint32_t sm[];
    const int tw = kBoxTW + 2 * r, th = kBoxTH + 2 * r;
    uint32_t* tile = sm;               // tw * th raw texels
    uint2* hs = reinterpret_cast<uint2*>(sm + ((tw * th + 1) & ~1));  // th * kBoxTW horizontal sums {RB, GA}
    const int x0 = blockIdx.x * kBoxTW, y0 = blockIdx.y * kBoxTH;
    const int tid = threadIdx.x;
    for (int i = tid; i < tw * th; i += kBoxThreads) {
        const int ty = i / tw, tx = i - ty * tw;
        const int gx = x0 - r + tx, gy = y0 - r + ty;
        uint32_t v = 0u;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = __ldg(in + (size_t)gy * W + gx);
        tile[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < th * kBoxTW; i += kBoxThreads) {
        const int ty = i / kBoxTW, tx = i - ty * kBoxTW;
        const uint32_t* row = tile + ty * tw + tx;
        uint32_t rb = 0u, ga = 0u;
        for (int d = 0; d <= 2 * r; ++d) {
            const uint32_t v = row[d];
            rb += v & 0x00FF00FFu;
            ga += (v >> 8) & 0x00FF00FFu;
        }
        hs[i] = make_uint2(rb, ga);
    }
    __syncthreads();
    for (int i = tid; i < kBoxTH * kBoxTW; i += kBoxThreads) {
        const int ty = i / kBoxTW, tx = i - ty * kBoxTW;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= W || y >= H) continue;
        uint32_t sr = 0u, sg = 0u, sb = 0u;
        for (int d = 0; d <= 2 * r; ++d) {
            const uint2 h = hs[(ty + d) * kBoxTW + tx];
            sr += h.x & 0xFFFFu;
            sb += h.x >> 16;
            sg += h.y & 0xFFFFu;
        }
        const int cx = min(x + r, W - 1) - max(x - r, 0) + 1;
        const int cy = min(y + r, H - 1) - max(y - r, 0) + 1;
        const float norm = (float)(cx * cy);
        const uint32_t R = (uint32_t)(unsigned char)__fdiv_rn((float)sr, norm);
        uint32_t G, B;
        if (replicate_r) {  // filterKernelBaseline writes acum.x to all three channels (src/filter.cu:51-53)
            G = R; B = R;
        } else {
            G = (uint32_t)(unsigned char)__fdiv_rn((float)sg, norm);
            B = (uint32_t)(unsigned char)__fdiv_rn((float)sb, norm);
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
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        RMD_CUDA_TRY(cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((f->width + kBoxTW - 1) / kBoxTW, (f->height + kBoxTH - 1) / kBoxTH);
    for (int level = 0; level < p->depth; ++level) {
        // ping-pong of the reference (src/filter.cu:24-25)
        const void* in = level == 0 ? f->render : f->buffer[level % 2];
        void* out = level == p->depth - 1 ? f->denoised : f->buffer[(level + 1) % 2];
        box_kernel<<<grid, kBoxThreads, smem, s>>>((const uint32_t*)in, (uint32_t*)out, f->width, f->height, p->radius,
                                                   replicate_r);
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
