// filter_compat.cu — the reference's two caller-launched entry points as real __global__ symbols
// (reference include/filter.cuh:25-26, defined in its src/filter.cu:13-58 and 87-158), built as relocatable device
// code into librmd_compat.a so that a caller's `filterKernelBaseline<<<grid, block, smem>>>(frame, params)`
// (src/test.cu:73-75, 85-87) links against this library instead of the reference's src/filter.cu, unchanged.
//
// The caller owns the launch geometry, so these kernels cannot use the library's own tiling (register strips,
// TMA tiles: csrc/box_filter.cu, csrc/svgf_atrous_tile.cu).  What they guarantee instead:
//   * correct for ANY 2-D block, ANY covering grid and ANY dynamic shared-memory size.  Each block stages the
//     neighbourhood of as large a sub-tile of its pixels as the launch's shared memory holds (two ping-pong windows of
//     (w + 2*depth*radius) x (h + 2*depth*radius) texels per staged plane) and walks its sub-tiles; with no usable
//     shared memory a single level reads its taps from global memory, like the reference's baseline kernel (several
//     levels then raise rmdCompatLastError(): they need at least (1 + 2*depth*radius)^2 * 8 bytes);
//   * `depth` levels inside ONE launch without the reference's cross-block race (its level loop ends in a
//     block-local __syncthreads, src/filter.cu:56): the block recomputes the halo of every intermediate level
//     itself ("fusing levels through halo recompute"), so the result equals `depth` host-iterated launches, which
//     is what oracle/oracle_box.c and rmd_filter_* compute.  The intermediate planes of the block's own pixels are
//     still written to frame.buffer[] when those pointers are non-null (the reference's ping-pong, :24-25);
//   * AVERAGE arithmetic identical to the reference (float sums, IEEE division, truncation; red channel replicated
//     by filterKernelBaseline, :51-53); GAUSSIAN / CROSS: csrc/weighted.cuh; WAVELET: not servable from a stateless
//     launch (needs history planes) -> no-op + rmdCompatLastError().
#include "filter.cuh"

#include "../weighted.cuh"
#include "../../../include/rmd_b200.h"

namespace {

__device__ int g_compat_error = 0;

struct LevelCtx {
    int W, H, radius, depth;
    int type;         // FilterParams::FilterType
    bool replicate;   // filterKernelBaseline: red channel only
    bool use_a, use_n;
    rmd::WeightScales k;
};

__device__ __forceinline__ unsigned dynamic_smem_bytes() {
    unsigned r;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(r));
    return r;
}

// one output texel of one level from the window `src` (row stride sw, origin (ox, oy) in image coordinates);
// guide windows ga / gn share the geometry of the level-0 window (origin (gx0, gy0), stride gsw)
__device__ __forceinline__ uint32_t level_texel(const LevelCtx& c, const uint32_t* src, int sw, int ox, int oy, int x, int y,
                                                const uint32_t* ga, const uint32_t* gn, int gsw, int gx0, int gy0) {
    const int r = c.radius;
    if (c.type == FilterParams::AVERAGE) {
        float ax = 0.f, ay = 0.f, az = 0.f, norm = 0.f;
        for (int dx = -r; dx <= r; ++dx)          // x outer, y inner: src/filter.cu:34-35
            for (int dy = -r; dy <= r; ++dy) {
                const int nx = x + dx, ny = y + dy;
                if (nx < 0 || nx >= c.W || ny < 0 || ny >= c.H) continue;   // :38-39
                const uint32_t m = src[(ny - oy) * sw + (nx - ox)];
                ax += (float)(m & 0xFFu);
                ay += (float)((m >> 8) & 0xFFu);
                az += (float)((m >> 16) & 0xFFu);
                norm += 1.0f;
            }
        const uint32_t ox8 = (uint32_t)(unsigned char)__fdiv_rn(ax, norm);    // :48-51: IEEE division, truncation
        if (c.replicate) return ox8 | (ox8 << 8) | (ox8 << 16);              // :51-53 (sic: .x three times)
        const uint32_t oy8 = (uint32_t)(unsigned char)__fdiv_rn(ay, norm), oz8 = (uint32_t)(unsigned char)__fdiv_rn(az, norm);
        return ox8 | (oy8 << 8) | (oz8 << 16);                               // :151-155, .w = 0
    }
    const uint32_t cp = src[(y - oy) * sw + (x - ox)];
    const uint32_t ap = c.use_a ? ga[(y - gy0) * gsw + (x - gx0)] : 0u;
    const uint32_t np = c.use_n ? gn[(y - gy0) * gsw + (x - gx0)] : 0u;
    rmd::WeightedAcc acc{0.f, 0.f, 0.f, 0.f};
    for (int dx = -r; dx <= r; ++dx)
        for (int dy = -r; dy <= r; ++dy) {
            const int nx = x + dx, ny = y + dy;
            if (nx < 0 || nx >= c.W || ny < 0 || ny >= c.H) continue;
            const uint32_t cq = src[(ny - oy) * sw + (nx - ox)];
            const uint32_t aq = c.use_a ? ga[(ny - gy0) * gsw + (nx - gx0)] : 0u;
            const uint32_t nq = c.use_n ? gn[(ny - gy0) * gsw + (nx - gx0)] : 0u;
            rmd::weighted_tap(acc, c.k, dx * dx + dy * dy, cp, cq, ap, aq, np, nq);
        }
    return rmd::weighted_finish(acc);
}

// single level, taps straight from global memory (no shared memory available)
__device__ void single_level_global(const LevelCtx& c, const GBuffer& f) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= c.W || y >= c.H) return;
    const uint32_t v = level_texel(c, (const uint32_t*)f.render, c.W, 0, 0, x, y, (const uint32_t*)f.albedo,
                                   (const uint32_t*)f.normal, c.W, 0, 0);
    ((uint32_t*)f.denoised)[(size_t)y * c.W + x] = v;
}

// `depth` levels of a (tw x th) sub-tile at (tx0, ty0) through ping-pong windows in `mem` (shared or local memory).
// The whole block cooperates.
__device__ void tile_levels(const LevelCtx& c, const GBuffer& f, uint32_t* mem, int tx0, int ty0, int tw, int th, int tid,
                            int nt) {
    const int R = c.radius * c.depth;
    const int ww = tw + 2 * R, wh = th + 2 * R, wn = ww * wh;
    const int wx0 = tx0 - R, wy0 = ty0 - R;
    uint32_t* win[2] = {mem, mem + wn};
    uint32_t* ga = mem + 2 * wn;                       // guide windows (CROSS only), level-0 geometry
    uint32_t* gn = ga + (c.use_a ? wn : 0);
    const uint32_t* render = (const uint32_t*)f.render;
    for (int i = tid; i < wn; i += nt) {
        const int gx = wx0 + i % ww, gy = wy0 + i / ww;
        const bool in = gx >= 0 && gx < c.W && gy >= 0 && gy < c.H;
        const size_t q = (size_t)gy * c.W + gx;
        win[0][i] = in ? render[q] : 0u;
        if (c.use_a) ga[i] = in ? ((const uint32_t*)f.albedo)[q] : 0u;
        if (c.use_n) gn[i] = in ? ((const uint32_t*)f.normal)[q] : 0u;
    }
    __syncthreads();
    for (int level = 0; level < c.depth; ++level) {
        const int hl = c.radius * (c.depth - 1 - level);   // halo still needed after this level
        const int ow = tw + 2 * hl, oh = th + 2 * hl, ox0 = tx0 - hl, oy0 = ty0 - hl;
        const uint32_t* src = win[level & 1];
        uint32_t* dst = win[(level + 1) & 1];
        const bool last = level == c.depth - 1;
        // the reference's ping-pong planes (src/filter.cu:24-25), own pixels only
        uint32_t* plane = last ? (uint32_t*)f.denoised : (uint32_t*)f.buffer[(level + 1) % 2];
        for (int i = tid; i < ow * oh; i += nt) {
            const int x = ox0 + i % ow, y = oy0 + i / ow;
            if (x < 0 || x >= c.W || y < 0 || y >= c.H) continue;
            const uint32_t v = level_texel(c, src, ww, wx0, wy0, x, y, ga, gn, ww, wx0, wy0);
            if (!last) dst[(y - wy0) * ww + (x - wx0)] = v;
            if (plane && x >= tx0 && x < tx0 + tw && y >= ty0 && y < ty0 + th) plane[(size_t)y * c.W + x] = v;
        }
        __syncthreads();
    }
}

__device__ void filter_entry(const GBuffer& frame, const FilterParams& params, bool replicate) {
    LevelCtx c;
    c.W = frame.shape.x; c.H = frame.shape.y; c.radius = params.radius; c.depth = params.depth;
    c.type = (int)params.type; c.replicate = replicate;
    c.use_a = c.use_n = false;
    c.k = rmd::WeightScales{0.f, 0.f, 0.f, 0.f};
    if (c.W <= 0 || c.H <= 0 || c.depth < 1 || c.radius < 0 || !frame.render || !frame.denoised) {
        g_compat_error = RMD_E_PARAM;
        return;
    }
    if (c.type == FilterParams::WAVELET || c.type < 0 || c.type > 3) {
        g_compat_error = RMD_E_UNSUPPORTED;   // SVGF needs a per-sequence context: rmd_svgf_frame_gbuffer
        return;
    }
    if (c.type != FilterParams::AVERAGE) {
        // same double-precision derivation as the host entry points (csrc/weighted.cuh) so that both agree bit for bit
        const double log2e = 1.4426950408889634;
        const double ss = params.sigmaSpace > 0 ? params.sigmaSpace : 0.5 * (c.radius > 1 ? c.radius : 1);
        c.k.ks = (float)(log2e / (2.0 * ss * ss));
        if (c.type == FilterParams::CROSS) {
            if (params.sigmaColor > 0) c.k.kc = (float)(log2e / (2.0 * (double)params.sigmaColor * params.sigmaColor * 65025.0));
            if (params.sigmaAlbedo > 0) c.k.ka = (float)(log2e / (2.0 * (double)params.sigmaAlbedo * params.sigmaAlbedo * 65025.0));
            if (params.sigmaNormal > 0) c.k.kn = (float)(log2e / (2.0 * (double)params.sigmaNormal * params.sigmaNormal * 65025.0));
        }
        c.use_a = c.k.ka > 0.f;
        c.use_n = c.k.kn > 0.f;
        if ((c.use_a && !frame.albedo) || (c.use_n && !frame.normal)) {
            g_compat_error = RMD_E_NULL;
            return;
        }
    }
    extern __shared__ uint32_t dyn_smem[];
    const int nt = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int planes = 2 + (c.use_a ? 1 : 0) + (c.use_n ? 1 : 0);
    const int R = c.radius * c.depth;
    const long long avail = (long long)dynamic_smem_bytes() / 4;
    // largest sub-tile (halving the longer side) whose windows fit the launch's dynamic shared memory
    int tw = blockDim.x, th = blockDim.y;
    while ((long long)(tw + 2 * R) * (th + 2 * R) * planes > avail && (tw > 1 || th > 1)) {
        if (tw >= th) tw = (tw + 1) / 2; else th = (th + 1) / 2;
    }
    const int bx0 = blockIdx.x * blockDim.x, by0 = blockIdx.y * blockDim.y;
    if ((long long)(tw + 2 * R) * (th + 2 * R) * planes <= avail) {
        for (int sy = 0; sy < (int)blockDim.y; sy += th)
            for (int sx = 0; sx < (int)blockDim.x; sx += tw) {
                const int w = min(tw, (int)blockDim.x - sx), h = min(th, (int)blockDim.y - sy);
                if (bx0 + sx >= c.W || by0 + sy >= c.H) continue;   // uniform per block
                tile_levels(c, frame, dyn_smem, bx0 + sx, by0 + sy, w, h, tid, nt);
            }
        return;
    }
    // no usable shared memory
    if (c.depth == 1) {
        single_level_global(c, frame);
        return;
    }
    // several levels need windows: (1 + 2*depth*radius)^2 * 8 bytes of dynamic shared memory at the very least
    // (3.5 KB for depth 5, radius 2; the reference's call sites pass 30-48 KB, src/test.cu:73, 85)
    g_compat_error = RMD_E_PARAM;
}

}  // namespace

KERNEL void filterKernelBaseline(GBuffer frame, const FilterParams params) { filter_entry(frame, params, true); }
KERNEL void filterKernelTiled(GBuffer frame, const FilterParams params) { filter_entry(frame, params, false); }

int rmdCompatLastError() {
    int e = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return RMD_E_STATE;
    if (cudaMemcpyFromSymbol(&e, g_compat_error, sizeof(e)) != cudaSuccess) return RMD_E_STATE;
    return e;
}
