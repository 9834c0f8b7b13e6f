// svgf_ctx.cu — per-sequence context (history planes, tensor maps, copy streams) and
// the extern "C" entry points of the SVGF path declared in include/rmd_b200.h.
//
// Reference shape being replaced: the never-implemented CudaGBuffer frame pipeline
// (reference include/gbuffer.h:20-33: allocate -> openImages(stream) -> kernel with
// depth=N ping-ponging buffer[0/1] -> denoisedCPU readback) and the caller-side
// kernel launches (src/test.cu:73-77).  Ownership follows the reference: caller
// planes are raw-pointer views (include/gbuffer.h:6-14) and are never freed here;
// history and scratch planes belong to the opaque context.
#include <new>
#include <stdlib.h>
#include <string.h>

#include "svgf.cuh"

using namespace rmd;

static_assert(sizeof(RmdGBuffer) == 56, "must mirror reference struct GBuffer (include/gbuffer.h:6-14)");
static_assert(offsetof(RmdGBuffer, render) == 8 && offsetof(RmdGBuffer, buffer) == 40, "GBuffer layout");
static_assert(sizeof(RmdFilterParams) == 36, "must mirror reference struct FilterParams (include/filter.cuh:11-23)");
static_assert(offsetof(RmdFilterParams, sigmaSpace) == 16 && offsetof(RmdFilterParams, cacheInput) == 32,
              "FilterParams layout");

namespace {

enum { kC4Hist = 0, kC4A = 1, kC4B = 2 };

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

struct rmd_svgf_ctx {
    int W = 0, H = 0, Wp = 0, Hp = 0, device = 0;
    size_t texels = 0;
    float4* c4[3] = {};
    float* v[3] = {};      // variance planes; each sits 256 B inside its allocation (v_raw): the a-trous prologue
    float* v_raw[3] = {};  // reads one element beyond either end of the plane and discards it
    float4* g4[2] = {};
    float* dz = nullptr;
    float2* m[2] = {};
    uint8_t* n[2] = {};
    float4* side_c4 = nullptr;
    uint32_t* tile_list = nullptr;    // compact list of flagged tiles (temporal -> variance)
    uint32_t* tile_count = nullptr;   // two counters, used alternately by frame parity
    uint32_t tile_capacity = 0;
    int parity = 0;
    int have_history = 0;
    int stop_after = 0;
    int last_launches = 0;
    int use_tma = 1;
    AtrousMaps maps[kMaxLevels][2];       // [level][guide parity], boxes of TY+4 rows (tile kernel)
    AtrousMaps ring_maps[kMaxLevels][2];  // same planes, boxes of 4 rows (ring kernel)
    int use_ring = 0;
    int variant[kMaxLevels] = {};  // tile-kernel variant per level (RMD_ATROUS_VARIANT = "n" or "n0,n1,n2,n3,n4")
    int var_dense_min = 128;       // variance pass: tiles with at least this many short-history pixels take the position-mapped path (RMD_VAR_DENSE_MIN; 257 = never, 0 = always)
    int var_threads = 128;         // variance pass CTA size (RMD_VAR_THREADS=256: the round-1 shape)
    int var_reverse = 1;           // variance pass walks the tile list last to first (RMD_VAR_REVERSE=0: first to last)
    int atrous_prefetch = 0;       // a-trous levels: tiles of look-ahead of the L2 tensor prefetch (RMD_ATROUS_PREFETCH; measured slower, off)
    int pdl = 5;                   // programmatic dependent launch, bit 0: level kernels, bit 1: temporal (measured slower: +16 us at 1080p, +64 us at 4K), bit 2: variance (RMD_PDL=<mask>)
    // host-frame path
    cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_compute[2] = {}, ev_d2h[2] = {};
    void *d_color[2] = {}, *d_albedo[2] = {}, *d_guide[2] = {}, *d_motion[2] = {}, *d_out[2] = {}, *d_out8[2] = {};
    unsigned long long host_frames = 0;
    int host_ready = 0;      // staging of the host-frame path exists (rmd_svgf_prepare_host or the first host frame)
    int device_frames = 0;   // rmd_svgf_frame / _gbuffer / band stages were used on this context
    // conversion planes of rmd_svgf_frame_gbuffer (allocated on first use)
    void *g_color = nullptr, *g_guide = nullptr, *g_motion = nullptr, *g_out = nullptr;
    // row-band mode (rmd_svgf_band_*)
    int band_row0 = 0, band_rows = 0;
    unsigned long long band_frame = 0;
    unsigned int* band_counter = nullptr;      // two last-block counters: [0] main-stream transfers, [1] push stream
    unsigned int* band_err_host = nullptr;     // pinned, device-mapped: number of flag waits that gave up (sticky error)
    unsigned int* band_err_dev = nullptr;
    cudaStream_t s_push = nullptr;             // peer pushes run beside the interior launch
    cudaEvent_t ev_boundary = nullptr, ev_pushed = nullptr;
    int push_pending = 0;
    int band_edge_stream = 1;   // a split level's boundary-tile launch runs on the push stream, beside the interior launch (RMD_BAND_EDGE_STREAM=0: before it, on the frame's stream)
    int band_early_unpack = 1;  // a level's halo rows are unpacked on the push stream right after the own push of the level before, beside that level's interior launch (RMD_BAND_EARLY_UNPACK=0: on the frame's stream, at the start of the level)
    unsigned band_unpacked = 0; // bit r: region r has already been unpacked on the push stream
    int band_serpentine = 1;    // band levels alternate the direction of their tile walk like the single-context frame (RMD_BAND_SERPENTINE=0: off)
    int band_launches = 0;
    // per-pass profiling
    int profiling = 0;
    int n_marks = 0;
    cudaEvent_t marks[kMaxLevels + 4] = {};
};

namespace {

int resolve(const RmdFilterParams* fp, const RmdSvgfParams* sp, SvgfConsts* r) {
    if (!fp) return RMD_E_NULL;
    if (fp->type != RMD_FILTER_WAVELET) return fp->type >= 0 && fp->type <= 3 ? RMD_E_UNSUPPORTED : RMD_E_PARAM;
    if (fp->radius != 2) return RMD_E_PARAM;  // 5x5 B3-spline footprint (reference waveletSpline has 3 taps)
    if (fp->depth < 0 || fp->depth > RMD_SVGF_MAX_LEVELS) return RMD_E_PARAM;
    if (fp->sigmaSpace < 0 || fp->sigmaColor < 0 || fp->sigmaNormal < 0) return RMD_E_PARAM;
    r->depth = fp->depth;
    r->sigma_z = fp->sigmaSpace > 0 ? fp->sigmaSpace : 1.0f;
    r->sigma_l = fp->sigmaColor > 0 ? fp->sigmaColor : 4.0f;
    r->sigma_n = fp->sigmaNormal > 0 ? fp->sigmaNormal : 128.0f;
    r->alpha_c = sp && sp->alpha_color > 0 ? sp->alpha_color : 0.05f;
    r->alpha_m = sp && sp->alpha_moments > 0 ? sp->alpha_moments : 0.2f;
    r->cap = sp && sp->history_cap > 0 ? sp->history_cap : 32;
    if (r->cap > 255) r->cap = 255;
    r->short_hist = sp && sp->short_history > 0 ? sp->short_history : 4;
    r->dtol = sp && sp->depth_tolerance > 0 ? sp->depth_tolerance : 0.1f;
    r->nthr = sp && sp->normal_threshold > 0 ? sp->normal_threshold : 0.9f;
    r->afloor = sp && sp->albedo_floor > 0 ? sp->albedo_floor : 1e-3f;
    r->lscale = sp && sp->variance_lum_scale > 0 ? sp->variance_lum_scale : 10.0f;
    return 0;
}

template <typename T>
int dev_alloc_zero(T** p, size_t bytes) {
    RMD_CUDA_TRY(cudaMalloc((void**)p, bytes));
    RMD_CUDA_TRY(cudaMemset(*p, 0, bytes));
    return 0;
}

// level input plane: level 0 reads the temporal output A, level 1 the history plane
// (level-0 output), then B/A alternate.
int level_in(int l) { return l == 0 ? kC4A : (l == 1 ? kC4Hist : ((l & 1) ? kC4A : kC4B)); }
int level_out(int l) { return l == 0 ? kC4Hist : ((l & 1) ? kC4B : kC4A); }

int build_maps(rmd_svgf_ctx* c, bool ring) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return RMD_E_DRIVER;
    for (int l = 0; l < kMaxLevels; ++l) {
        const int S = 1 << l;
        const int wt = ring ? kAtrousWT : atrous_variant_tile_width(c->variant[l], l);
        const cuuint32_t tw = (cuuint32_t)(wt + 2 * (2 * S < 4 ? 4 : 2 * S)), th = ring ? 4u : (cuuint32_t)(kAtrousTY + 4);
        for (int par = 0; par < 2; ++par) {
            AtrousMaps& mp = ring ? c->ring_maps[l][par] : c->maps[l][par];
            // float4 planes as {x in 8-byte elements (2 per texel), phase, lattice row}: one box row is
            // TW/2 texels = 1-1.5 KB (a 16-byte inner dimension made TMA request-bound); two boxes per tile
            const cuuint64_t dims4[3] = {(cuuint64_t)c->W * 2, (cuuint64_t)S, (cuuint64_t)(c->Hp / S)};
            const cuuint64_t strides4[2] = {(cuuint64_t)c->Wp * 16, (cuuint64_t)c->Wp * 16 * S};
            const cuuint32_t box4[3] = {tw, 1, th};  // tw/2 texels * 2 elements
            const cuuint32_t ones4[4] = {1, 1, 1, 1};
            CUresult r = enc(&mp.c4, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, c->c4[level_in(l)], dims4, strides4, box4, ones4,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return RMD_E_DRIVER;
            r = enc(&mp.g4, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, c->g4[par], dims4, strides4, box4, ones4,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return RMD_E_DRIVER;
            const cuuint64_t dims3[3] = {(cuuint64_t)c->W, (cuuint64_t)S, (cuuint64_t)(c->Hp / S)};
            const cuuint64_t strides3[2] = {(cuuint64_t)c->Wp * 4, (cuuint64_t)c->Wp * 4 * S};
            const cuuint32_t box3[3] = {tw, 1, th};
            r = enc(&mp.v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->v[level_in(l)], dims3, strides3, box3, ones4,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return RMD_E_DRIVER;
            // rows y-1 / y+1 of the TY output rows (neighbour phases) for the 3x3 variance pre-filter, and the slope
            const cuuint32_t boxn[3] = {(cuuint32_t)(wt + 8), 1, (cuuint32_t)kAtrousTY};
            r = enc(&mp.vn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->v[level_in(l)], dims3, strides3, boxn, ones4,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return RMD_E_DRIVER;
            const cuuint32_t boxd[3] = {(cuuint32_t)wt, 1, (cuuint32_t)kAtrousTY};
            r = enc(&mp.dzm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->dz, dims3, strides3, boxd, ones4,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return RMD_E_DRIVER;
        }
    }
    return 0;
}

void free_all(rmd_svgf_ctx* c) {
    for (int i = 0; i < 3; ++i) { cudaFree(c->c4[i]); cudaFree(c->v_raw[i]); }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->g4[i]); cudaFree(c->m[i]); cudaFree(c->n[i]);
        cudaFree(c->d_color[i]); cudaFree(c->d_albedo[i]); cudaFree(c->d_guide[i]); cudaFree(c->d_motion[i]);
        cudaFree(c->d_out[i]); cudaFree(c->d_out8[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_compute[i]) cudaEventDestroy(c->ev_compute[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    for (auto& e : c->marks) if (e) cudaEventDestroy(e);
    cudaFree(c->g_color); cudaFree(c->g_guide); cudaFree(c->g_motion); cudaFree(c->g_out);
    cudaFree(c->band_counter);
    if (c->band_err_host) cudaFreeHost(c->band_err_host);
    if (c->s_push) cudaStreamDestroy(c->s_push);
    if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
    if (c->ev_pushed) cudaEventDestroy(c->ev_pushed);
    cudaFree(c->dz); cudaFree(c->side_c4); cudaFree(c->tile_list); cudaFree(c->tile_count);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
}

int create_impl(rmd_svgf_ctx* c) {
    const size_t t = c->texels;
    for (int i = 0; i < 3; ++i) {
        int rc = dev_alloc_zero(&c->c4[i], t * 16); if (rc) return rc;
        rc = dev_alloc_zero(&c->v_raw[i], t * 4 + 512); if (rc) return rc;
        c->v[i] = c->v_raw[i] + 64;
    }
    for (int i = 0; i < 2; ++i) {
        int rc = dev_alloc_zero(&c->g4[i], t * 16); if (rc) return rc;
        rc = dev_alloc_zero(&c->m[i], t * 8); if (rc) return rc;
        rc = dev_alloc_zero(&c->n[i], t); if (rc) return rc;
    }
    int rc = dev_alloc_zero(&c->dz, t * 4); if (rc) return rc;
    rc = dev_alloc_zero(&c->side_c4, t * 16); if (rc) return rc;
    // one entry per 32x8 temporal tile (+ one tile row of slack: a band's temporal launch starts at an arbitrary row)
    const size_t ntiles = (size_t)((c->W + kTemporalBx - 1) / kTemporalBx) * ((c->H + kTemporalBy - 1) / kTemporalBy + 1);
    rc = dev_alloc_zero(&c->tile_list, ntiles * 4); if (rc) return rc;
    c->tile_capacity = (uint32_t)ntiles;
    rc = dev_alloc_zero(&c->tile_count, 2 * 4); if (rc) return rc;
    rc = atrous_configure(); if (rc) return rc;
    const char* no_tma = getenv("RMD_NO_TMA");
    c->use_tma = !(no_tma && no_tma[0] == '1');
    // tile-kernel variant per level (A/B switch; the default is the measured-fastest, profiles/r2_notes.md)
    for (int l = 0; l < kMaxLevels; ++l) c->variant[l] = kAtrousDefaultVariant[l];
    if (const char* v = getenv("RMD_ATROUS_VARIANT")) {
        int l = 0, last = -1;
        for (const char* p = v; *p && l < kMaxLevels;) {
            char* end = nullptr;
            const long n = strtol(p, &end, 10);
            if (end == p) break;
            if (!atrous_variant_exists((int)n)) return RMD_E_PARAM;
            c->variant[l++] = last = (int)n;
            p = (*end == ',') ? end + 1 : end;
        }
        for (; l < kMaxLevels && last >= 0; ++l) c->variant[l] = last;
    }
    rc = build_maps(c, false);
    if (rc) return rc;
    rc = build_maps(c, true);
    if (rc) return rc;
    // default: independent TMA tiles (measured faster on B200 than the ring: 57.5 vs 65 us per level at 1080p,
    // 199 vs 232 us at 4K when both were last compared, profiles/r1_notes.md); RMD_ATROUS_RING=1 selects the persistent ring kernel for levels 0..3
    const char* ring = getenv("RMD_ATROUS_RING");
    c->use_ring = c->use_tma && ring && ring[0] == '1';
    if (const char* pdl = getenv("RMD_PDL")) c->pdl = atoi(pdl) & 7;
    if (const char* e = getenv("RMD_VAR_DENSE_MIN")) { if (*e) c->var_dense_min = atoi(e); }
    if (const char* e = getenv("RMD_VAR_THREADS")) { if (atoi(e) == 256) c->var_threads = 256; }
    if (const char* e = getenv("RMD_VAR_REVERSE")) { if (*e) c->var_reverse = atoi(e) != 0; }
    if (const char* e = getenv("RMD_ATROUS_PREFETCH")) { if (*e && atoi(e) > 0) c->atrous_prefetch = atoi(e); }
    if (const char* e = getenv("RMD_BAND_EDGE_STREAM")) { if (*e) c->band_edge_stream = atoi(e) != 0; }
    if (const char* e = getenv("RMD_BAND_SERPENTINE")) { if (*e) c->band_serpentine = atoi(e) != 0; }
    if (const char* e = getenv("RMD_BAND_EARLY_UNPACK")) { if (*e) c->band_early_unpack = atoi(e) != 0; }
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_frame(const rmd_svgf_ctx* c, const RmdSvgfFrame* f, bool device_ptrs) {
    if (!c || !f) return RMD_E_NULL;
    if (f->width != c->W || f->height != c->H) return RMD_E_SHAPE;
    if (!f->color || !f->albedo || !f->guide || !f->motion || !f->out) return RMD_E_NULL;
    if (device_ptrs) {
        if (((uintptr_t)f->color | (uintptr_t)f->guide | (uintptr_t)f->out) & 15u) return RMD_E_ALIGN;
        if (((uintptr_t)f->albedo | (uintptr_t)f->motion | (uintptr_t)f->out_rgba8) & 3u) return RMD_E_ALIGN;
    }
    return 0;
}

#define RMD_MARK()                                                        \
    do {                                                                  \
        if (c->profiling) RMD_CUDA_TRY(cudaEventRecord(c->marks[c->n_marks++], s)); \
    } while (0)

int frame_impl(rmd_svgf_ctx* c, const RmdSvgfFrame* f, const SvgfConsts& k, cudaStream_t s) {
    int launches = 0;
    c->n_marks = 0;
    RMD_MARK();
    c->parity ^= 1;
    const int cur = c->parity, prv = cur ^ 1;
    TemporalArgs ta{};
    ta.color = (const uint2*)f->color; ta.albedo = (const uint32_t*)f->albedo;
    ta.guide = (const uint2*)f->guide; ta.motion = (const uint32_t*)f->motion;
    ta.hist_c4 = c->c4[kC4Hist]; ta.hist_m = c->m[prv]; ta.hist_n = c->n[prv]; ta.prev_g4 = c->g4[prv];
    ta.out_c4 = c->c4[kC4A]; ta.out_v = c->v[kC4A]; ta.out_m = c->m[cur]; ta.out_n = c->n[cur];
    ta.out_g4 = c->g4[cur]; ta.out_dz = c->dz; ta.side_c4 = c->side_c4;
    ta.tile_list = c->tile_list; ta.tile_count = c->tile_count + cur; ta.tile_capacity = c->tile_capacity;
    ta.W = c->W; ta.H = c->H; ta.Wp = c->Wp; ta.row_begin = 0; ta.row_end = c->H;
    ta.full_begin = 0; ta.full_end = c->H;
    ta.hist_row_lo = 0; ta.hist_row_hi = c->H;
    ta.have_history = c->have_history; ta.k = k;
    int rc = launch_temporal(ta, s, (c->pdl & 2) != 0); if (rc) return rc;
    launches += 1;
    RMD_MARK();
    c->have_history = 1;
    if (c->stop_after == 1) {  // test hook: the variance pass, which zeroes the next frame's tile counter, does not run
        RMD_CUDA_TRY(cudaMemsetAsync(c->tile_count, 0, 2 * 4, s));
        c->last_launches = launches;
        return 0;
    }

    VarianceArgs va{};
    va.c4 = c->c4[kC4A]; va.m = c->m[cur]; va.n = c->n[cur]; va.g4 = c->g4[cur]; va.dz = c->dz;
    va.side_c4 = c->side_c4; va.patch_c4 = c->c4[kC4A]; va.patch_v = c->v[kC4A];
    va.tile_list = c->tile_list; va.tile_count = c->tile_count + cur; va.tile_capacity = c->tile_capacity; va.next_count = c->tile_count + prv;
    va.W = c->W; va.H = c->H; va.Wp = c->Wp; va.k = k;
    va.row_begin = 0; va.row_end = c->H;
    va.dense_min = c->var_dense_min; va.threads = c->var_threads; va.reverse = c->var_reverse;
    rc = launch_variance(va, s, (c->pdl & 4) != 0); if (rc) return rc;
    launches += 1;
    RMD_MARK();
    if (c->stop_after == 2) { c->last_launches = launches; return 0; }

    if (k.depth == 0) {
        // no spatial level: the temporal output is both the history and the result
        RMD_CUDA_TRY(cudaMemcpyAsync(c->c4[kC4Hist], c->c4[kC4A], c->texels * 16, cudaMemcpyDeviceToDevice, s));
        rc = launch_remodulate(c->c4[kC4A], c->v[kC4A], c->g4[cur], (const uchar4*)f->albedo, (float4*)f->out,
                               (uchar4*)f->out_rgba8, c->W, c->H, c->Wp, k.afloor, s);
        if (rc) return rc;
        launches += 1;
        RMD_MARK();
    }
    for (int l = 0; l < k.depth; ++l) {
        const bool last = l == k.depth - 1;
        AtrousArgs aa{};
        aa.in_c4 = c->c4[level_in(l)]; aa.in_v = c->v[level_in(l)];
        aa.g4 = c->g4[cur]; aa.dz = c->dz;
        // level 0 always writes the history plane; the last level writes the caller's output
        const bool write_planes = !last || l == 0;
        aa.out_c4 = write_planes ? c->c4[level_out(l)] : nullptr;
        aa.out_v = write_planes ? c->v[level_out(l)] : nullptr;
        aa.final_out = last ? (float4*)f->out : nullptr;
        aa.final_rgba8 = last ? (uchar4*)f->out_rgba8 : nullptr;
        aa.albedo = (const uchar4*)f->albedo;
        aa.W = c->W; aa.H = c->H; aa.Wp = c->Wp; aa.Hp = c->Hp; aa.row0 = 0; aa.rows = c->H;
        aa.sigma_z = k.sigma_z; aa.sigma_l = k.sigma_l; aa.sigma_n = k.sigma_n; aa.afloor = k.afloor;
        aa.use_tma = c->use_tma; aa.prefetch_ahead = c->atrous_prefetch;
        aa.reverse = (l & 1) == 0;  // the temporal pass wrote top-down: level 0 walks bottom-up, level 1 top-down, ...
        // ring kernel for steps 1..8; at step 16 the ring (192-texel rows) has no shared memory left to
        // prefetch with and the independent-tile kernel is faster (profiles/r1_notes.md)
        rc = (c->use_ring && l < 4) ? launch_atrous_ring(l, aa, c->ring_maps[l][cur], s)
                                    : launch_atrous(l, aa, c->maps[l][cur], s, c->variant[l], (c->pdl & 1) != 0);
        if (rc) return rc;
        launches += 1;
        RMD_MARK();
    }
    c->last_launches = launches;
    return 0;
}

}  // namespace

extern "C" int rmd_svgf_create(rmd_svgf_ctx** out, int width, int height, int device) {
    if (!out) return RMD_E_NULL;
    *out = nullptr;
    if (width < 1 || height < 1 || width > 32768 || height > 32768) return RMD_E_SHAPE;
    int ndev = 0;
    RMD_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return RMD_E_PARAM;
    DeviceGuard guard(device);
    rmd_svgf_ctx* c = new (std::nothrow) rmd_svgf_ctx();
    if (!c) return RMD_E_NOMEM;
    c->W = width; c->H = height; c->device = device;
    c->Wp = (width + 31) & ~31;
    c->Hp = (height + 15) & ~15;
    c->texels = (size_t)c->Wp * c->Hp;
    int rc = create_impl(c);
    if (!rc) rc = (int)cudaDeviceSynchronize();
    if (rc) { free_all(c); delete c; return rc; }
    *out = c;
    return 0;
}

extern "C" int rmd_svgf_destroy(rmd_svgf_ctx* c) {
    if (!c) return RMD_E_NULL;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    free_all(c);
    delete c;
    return 0;
}

extern "C" int rmd_svgf_reset(rmd_svgf_ctx* c) {
    if (!c) return RMD_E_NULL;
    c->have_history = 0;
    return 0;
}

extern "C" int rmd_svgf_set_stop_after(rmd_svgf_ctx* c, int stage) {
    if (!c) return RMD_E_NULL;
    if (stage < 0 || stage > 2) return RMD_E_PARAM;
    c->stop_after = stage;
    return 0;
}

extern "C" int rmd_svgf_set_profiling(rmd_svgf_ctx* c, int enable) {
    if (!c) return RMD_E_NULL;
    DeviceGuard guard(c->device);
    if (enable && !c->marks[0])
        for (auto& e : c->marks) RMD_CUDA_TRY(cudaEventCreate(&e));
    c->profiling = enable ? 1 : 0;
    c->n_marks = 0;
    return 0;
}

extern "C" int rmd_svgf_get_pass_times(rmd_svgf_ctx* c, float* ms, int capacity) {
    if (!c || !ms) return RMD_E_NULL;
    if (!c->profiling || c->n_marks < 2) return RMD_E_STATE;
    DeviceGuard guard(c->device);
    RMD_CUDA_TRY(cudaEventSynchronize(c->marks[c->n_marks - 1]));
    int n = 0;
    for (int i = 0; i + 1 < c->n_marks && n < capacity; ++i, ++n)
        RMD_CUDA_TRY(cudaEventElapsedTime(&ms[n], c->marks[i], c->marks[i + 1]));
    return n;
}

extern "C" int rmd_svgf_last_launch_count(const rmd_svgf_ctx* c) { return c ? c->last_launches : RMD_E_NULL; }

extern "C" int rmd_svgf_frame(rmd_svgf_ctx* c, const RmdSvgfFrame* f, const RmdFilterParams* fp, const RmdSvgfParams* sp,
                              void* stream) {
    int rc = check_frame(c, f, true);
    if (rc) return rc;
    SvgfConsts k;
    rc = resolve(fp, sp, &k);
    if (rc) return rc;
    if (c->host_frames) return RMD_E_STATE;  // the host-frame path runs on the context's own streams: do not mix
    DeviceGuard guard(c->device);
    c->device_frames = 1;
    return frame_impl(c, f, k, (cudaStream_t)stream);
}

namespace {
// Staging slots, streams and events of the host-frame path.  Everything is created into locals and committed to the
// context only when every call has succeeded, so a failed first call leaves the context unchanged (and retryable).
int host_path_init(rmd_svgf_ctx* c) {
    if (c->host_ready) return 0;
    const size_t px = (size_t)c->W * c->H;
    cudaStream_t st[3] = {};
    cudaEvent_t ev[6] = {};
    void* buf[12] = {};
    const size_t bytes[6] = {px * 8, px * 4, px * 8, px * 4, px * 16, px * 4};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 12 && e == cudaSuccess; ++i) e = cudaMalloc(&buf[i], bytes[i % 6]);
    if (e != cudaSuccess) {
        for (auto b : buf) cudaFree(b);
        for (auto v : ev) if (v) cudaEventDestroy(v);
        for (auto s : st) if (s) cudaStreamDestroy(s);
        return (int)e;
    }
    c->s_h2d = st[0]; c->s_compute = st[1]; c->s_d2h = st[2];
    for (int i = 0; i < 2; ++i) {
        c->ev_h2d[i] = ev[3 * i]; c->ev_compute[i] = ev[3 * i + 1]; c->ev_d2h[i] = ev[3 * i + 2];
        c->d_color[i] = buf[6 * i]; c->d_albedo[i] = buf[6 * i + 1]; c->d_guide[i] = buf[6 * i + 2];
        c->d_motion[i] = buf[6 * i + 3]; c->d_out[i] = buf[6 * i + 4]; c->d_out8[i] = buf[6 * i + 5];
    }
    c->host_ready = 1;
    return 0;
}
}  // namespace

extern "C" int rmd_svgf_prepare_host(rmd_svgf_ctx* c) {
    if (!c) return RMD_E_NULL;
    DeviceGuard guard(c->device);
    return host_path_init(c);
}

extern "C" int rmd_svgf_frame_host(rmd_svgf_ctx* c, const RmdSvgfFrame* f, const RmdFilterParams* fp,
                                   const RmdSvgfParams* sp) {
    if (!c || !f) return RMD_E_NULL;
    if (f->width != c->W || f->height != c->H) return RMD_E_SHAPE;
    // `out` may be null when only the RGBA8 result is wanted (the reference's `denoised` format): 4 B/px come back
    // over PCIe instead of 16
    if (!f->color || !f->albedo || !f->guide || !f->motion || (!f->out && !f->out_rgba8)) return RMD_E_NULL;
    SvgfConsts k;
    int rc = resolve(fp, sp, &k);
    if (rc) return rc;
    if (c->device_frames) return RMD_E_STATE;  // one context = one entry-point family (the history planes are shared)
    DeviceGuard guard(c->device);
    const size_t px = (size_t)c->W * c->H;
    rc = host_path_init(c);
    if (rc) return rc;
    const int slot = (int)(c->host_frames & 1ull);
    // the slot's inputs were last read by the frame submitted two calls ago
    if (c->host_frames >= 2) RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_compute[slot], 0));
    RMD_CUDA_TRY(cudaMemcpyAsync(c->d_color[slot], f->color, px * 8, cudaMemcpyHostToDevice, c->s_h2d));
    RMD_CUDA_TRY(cudaMemcpyAsync(c->d_albedo[slot], f->albedo, px * 4, cudaMemcpyHostToDevice, c->s_h2d));
    RMD_CUDA_TRY(cudaMemcpyAsync(c->d_guide[slot], f->guide, px * 8, cudaMemcpyHostToDevice, c->s_h2d));
    RMD_CUDA_TRY(cudaMemcpyAsync(c->d_motion[slot], f->motion, px * 4, cudaMemcpyHostToDevice, c->s_h2d));
    RMD_CUDA_TRY(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
    RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_h2d[slot], 0));
    // the slot's output was last drained by the download submitted two calls ago
    if (c->host_frames >= 2) RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_d2h[slot], 0));
    RmdSvgfFrame df = *f;
    df.color = c->d_color[slot]; df.albedo = c->d_albedo[slot]; df.guide = c->d_guide[slot];
    df.motion = c->d_motion[slot]; df.out = c->d_out[slot]; df.out_rgba8 = f->out_rgba8 ? c->d_out8[slot] : nullptr;
    rc = frame_impl(c, &df, k, c->s_compute);
    if (rc) return rc;
    RMD_CUDA_TRY(cudaEventRecord(c->ev_compute[slot], c->s_compute));
    RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_compute[slot], 0));
    if (f->out) RMD_CUDA_TRY(cudaMemcpyAsync(f->out, c->d_out[slot], px * 16, cudaMemcpyDeviceToHost, c->s_d2h));
    if (f->out_rgba8)
        RMD_CUDA_TRY(cudaMemcpyAsync(f->out_rgba8, c->d_out8[slot], px * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    RMD_CUDA_TRY(cudaEventRecord(c->ev_d2h[slot], c->s_d2h));
    c->host_frames++;
    return 0;
}


// ---- the reference's RGBA8 G-buffer as SVGF input (rmd_svgf_frame_gbuffer) ------------------------------
namespace {
__global__ void gbuffer_convert_kernel(const uchar4* __restrict__ render, const uchar4* __restrict__ normal,
                                       uint2* __restrict__ color16f, uint2* __restrict__ guide, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float k = 1.0f / 255.0f;
    const uchar4 r = render[i];
    const __half2 rg = __floats2half2_rn(__fmul_rn((float)r.x, k), __fmul_rn((float)r.y, k));
    const __half2 ba = __floats2half2_rn(__fmul_rn((float)r.z, k), 1.0f);
    color16f[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&ba));
    const uchar4 nn = normal ? normal[i] : make_uchar4(0, 0, 255, 0);
    float x = __fmul_rn((float)nn.x, k), y = __fmul_rn((float)nn.y, k), z = __fmul_rn((float)nn.z, k);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    if (len < 1e-6f) { x = 0.f; y = 0.f; z = 1.f; }
    else { const float inv = __fdiv_rn(1.0f, len); x = __fmul_rn(x, inv); y = __fmul_rn(y, inv); z = __fmul_rn(z, inv); }
    // octahedral encode, snorm16 round-to-nearest-even
    const float s = __fadd_rn(__fadd_rn(fabsf(x), fabsf(y)), fabsf(z));
    float px = __fdiv_rn(x, s), py = __fdiv_rn(y, s);
    if (z < 0.0f) {
        const float ox = __fmul_rn(__fsub_rn(1.0f, fabsf(py)), px >= 0.0f ? 1.0f : -1.0f);
        const float oy = __fmul_rn(__fsub_rn(1.0f, fabsf(px)), py >= 0.0f ? 1.0f : -1.0f);
        px = ox; py = oy;
    }
    const int sx = __float2int_rn(__fmul_rn(px, 32767.0f)), sy = __float2int_rn(__fmul_rn(py, 32767.0f));
    guide[i] = make_uint2((uint32_t)(sx & 0xFFFF) | ((uint32_t)(sy & 0xFFFF) << 16), __float_as_uint(1.0f));
}
}  // namespace

namespace {
// conversion planes of the RGBA8 G-buffer entry point: allocated together, committed only when all succeeded
int gbuffer_path_init(rmd_svgf_ctx* c, cudaStream_t s) {
    if (c->g_color) return 0;
    const size_t px = (size_t)c->W * c->H;
    void* buf[4] = {};
    const size_t bytes[4] = {px * 8, px * 8, px * 4, px * 16};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaMalloc(&buf[i], bytes[i]);
    if (e == cudaSuccess) e = cudaMemsetAsync(buf[2], 0, px * 4, s);  // motion = 0: the reference's GBuffer has no such plane
    if (e != cudaSuccess) {
        for (auto b : buf) cudaFree(b);
        return (int)e;
    }
    c->g_guide = buf[1]; c->g_motion = buf[2]; c->g_out = buf[3];
    c->g_color = buf[0];  // last: marks the set as complete
    return 0;
}
}  // namespace

extern "C" int rmd_svgf_prepare_gbuffer(rmd_svgf_ctx* c, void* stream) {
    if (!c) return RMD_E_NULL;
    DeviceGuard guard(c->device);
    return gbuffer_path_init(c, (cudaStream_t)stream);
}

extern "C" int rmd_svgf_frame_gbuffer(rmd_svgf_ctx* c, const RmdGBuffer* g, const RmdFilterParams* fp,
                                      const RmdSvgfParams* sp, void* out_rgba32f, void* stream) {
    if (!c || !g) return RMD_E_NULL;
    if (g->width != c->W || g->height != c->H) return RMD_E_SHAPE;
    if (!g->render || !g->albedo || !g->denoised) return RMD_E_NULL;
    if (((uintptr_t)g->render | (uintptr_t)g->albedo | (uintptr_t)g->normal | (uintptr_t)g->denoised) & 3u) return RMD_E_ALIGN;
    if ((uintptr_t)out_rgba32f & 15u) return RMD_E_ALIGN;
    SvgfConsts k;
    int rc = resolve(fp, sp, &k);
    if (rc) return rc;
    DeviceGuard guard(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t px = (size_t)c->W * c->H;
    rc = gbuffer_path_init(c, s);
    if (rc) return rc;
    c->device_frames = 1;
    gbuffer_convert_kernel<<<(unsigned)((px + 255) / 256), 256, 0, s>>>((const uchar4*)g->render, (const uchar4*)g->normal,
                                                                     (uint2*)c->g_color, (uint2*)c->g_guide, (int)px);
    RMD_CUDA_TRY(cudaGetLastError());
    RmdSvgfFrame f{c->W, c->H, c->g_color, g->albedo, c->g_guide, c->g_motion, out_rgba32f ? out_rgba32f : c->g_out, g->denoised};
    rc = frame_impl(c, &f, k, s);
    c->last_launches += 1;
    return rc;
}

extern "C" int rmd_svgf_host_wait(rmd_svgf_ctx* c) {
    if (!c) return RMD_E_NULL;
    if (!c->s_compute) return 0;
    DeviceGuard guard(c->device);
    RMD_CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    RMD_CUDA_TRY(cudaStreamSynchronize(c->s_compute));
    RMD_CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    return 0;
}

extern "C" int rmd_svgf_read_plane(rmd_svgf_ctx* c, int plane, void* host_dst, size_t host_bytes, void* stream) {
    if (!c || !host_dst) return RMD_E_NULL;
    DeviceGuard guard(c->device);
    const void* src = nullptr;
    size_t elem = 0;
    switch (plane) {
        case RMD_PLANE_TEMPORAL_COLOR: src = c->c4[kC4A]; elem = 16; break;
        case RMD_PLANE_TEMPORAL_VAR: src = c->v[kC4A]; elem = 4; break;
        case RMD_PLANE_MOMENTS: src = c->m[c->parity]; elem = 8; break;
        case RMD_PLANE_HISTLEN: src = c->n[c->parity]; elem = 1; break;
        case RMD_PLANE_HISTORY_COLOR: src = c->c4[kC4Hist]; elem = 16; break;
        case RMD_PLANE_GUIDE: src = c->g4[c->parity]; elem = 16; break;
        case RMD_PLANE_SLOPE: src = c->dz; elem = 4; break;
        default: return RMD_E_PARAM;
    }
    if (host_bytes < (size_t)c->W * c->H * elem) return RMD_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    RMD_CUDA_TRY(cudaMemcpy2DAsync(host_dst, (size_t)c->W * elem, src, (size_t)c->Wp * elem, (size_t)c->W * elem, c->H,
                                   cudaMemcpyDeviceToHost, s));
    RMD_CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

extern "C" size_t rmd_svgf_history_bytes(const rmd_svgf_ctx* c, int nrows) {
    return c && nrows > 0 ? (size_t)c->W * nrows * 41 : 0;
}

namespace {
// direction: 0 = context planes -> packed buffer, 1 = packed buffer -> context planes
int history_copy(rmd_svgf_ctx* c, int row_begin, int nrows, void* buf, cudaStream_t s, int direction) {
    if (!c || !buf) return RMD_E_NULL;
    if (row_begin < 0 || nrows < 1 || row_begin + nrows > c->H) return RMD_E_SHAPE;
    DeviceGuard guard(c->device);
    const int cur = c->parity;
    struct { void* plane; size_t elem; } planes[4] = {
        {c->c4[kC4Hist], 16}, {c->g4[cur], 16}, {c->m[cur], 8}, {c->n[cur], 1}};
    uint8_t* b = (uint8_t*)buf;
    for (auto& pl : planes) {
        uint8_t* p = (uint8_t*)pl.plane + (size_t)row_begin * c->Wp * pl.elem;
        const size_t wbytes = (size_t)c->W * pl.elem, pitch = (size_t)c->Wp * pl.elem;
        if (direction == 0) RMD_CUDA_TRY(cudaMemcpy2DAsync(b, wbytes, p, pitch, wbytes, nrows, cudaMemcpyDeviceToDevice, s));
        else RMD_CUDA_TRY(cudaMemcpy2DAsync(p, pitch, b, wbytes, wbytes, nrows, cudaMemcpyDeviceToDevice, s));
        b += wbytes * nrows;
    }
    return 0;
}
}  // namespace

extern "C" int rmd_svgf_history_pack(rmd_svgf_ctx* c, int row_begin, int nrows, void* buf, void* stream) {
    return history_copy(c, row_begin, nrows, buf, (cudaStream_t)stream, 0);
}
extern "C" int rmd_svgf_history_unpack(rmd_svgf_ctx* c, int row_begin, int nrows, const void* buf, void* stream) {
    return history_copy(c, row_begin, nrows, const_cast<void*>(buf), (cudaStream_t)stream, 1);
}


// =====================================================================================
// Row bands with per-level halo exchange (include/rmd_b200.h "Row bands with per-level halo exchange")
// =====================================================================================
namespace {

constexpr int kBandTemporalExt = 6, kBandVarianceExt = 3, kBandHistoryRows = 21;  // 6 + max|mv_y| margin 15

struct RowJob {
    const char* src;
    char* dst;
    unsigned src_pitch, dst_pitch, row_units, nrows;  // row_units = row bytes / 16
};
struct BandXfer {
    RowJob job[6];
    int njobs;
    const unsigned long long* wait_flag[2];
    unsigned long long* signal_flag[2];
    unsigned long long value;
    unsigned int* counter;
    unsigned int* err;  // host-mapped: bumped when a flag wait gives up
};

// One kernel per exchange point: (optionally) wait until the neighbours' rows have landed, copy row blocks with
// 16-byte accesses (destination may be peer-mapped memory of the neighbour GPU), then (optionally) publish.
__global__ void __launch_bounds__(256) band_xfer_kernel(const BandXfer x) {
    pdl_wait();  // launched with programmatic stream serialisation: the launch overlaps the previous kernel's tail
    pdl_launch_dependents();
    if (x.wait_flag[0] || x.wait_flag[1]) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            for (int d = 0; d < 2; ++d) {
                if (!x.wait_flag[d]) continue;
                const volatile unsigned long long* f = x.wait_flag[d];
                while (*f < x.value) {
                    __nanosleep(100);
                    if (clock64() - t0 > 4000000000LL) {  // ~2 s: a protocol bug or a dead neighbour must not hang the GPU;
                        atomicAdd_system(x.err, 1u);      // the host sees the word and fails every later call (sticky)
                        break;
                    }
                }
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    for (int j = 0; j < x.njobs; ++j) {
        const RowJob jb = x.job[j];
        const unsigned total = jb.row_units * jb.nrows;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const unsigned r = i / jb.row_units, u = i - r * jb.row_units;
            const uint4 v = *reinterpret_cast<const uint4*>(jb.src + (size_t)r * jb.src_pitch + (size_t)u * 16);
            *reinterpret_cast<uint4*>(jb.dst + (size_t)r * jb.dst_pitch + (size_t)u * 16) = v;
        }
    }
    if (x.signal_flag[0] || x.signal_flag[1]) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(x.counter, 1u) == gridDim.x - 1) {  // last block: everything is written
            *x.counter = 0u;
            __threadfence_system();
            for (int d = 0; d < 2; ++d)
                if (x.signal_flag[d]) *reinterpret_cast<volatile unsigned long long*>(x.signal_flag[d]) = x.value;
            __threadfence_system();
        }
    }
}

int launch_xfer(const BandXfer& x, cudaStream_t s, bool pdl, int max_grid = 592) {
    if (x.njobs == 0 && !x.wait_flag[0] && !x.wait_flag[1] && !x.signal_flag[0] && !x.signal_flag[1]) return 0;
    size_t units = 0;
    for (int j = 0; j < x.njobs; ++j) units += (size_t)x.job[j].row_units * x.job[j].nrows;
    int grid = (int)((units + 255) / 256);
    grid = grid < 1 ? 1 : (grid > max_grid ? max_grid : grid);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, band_xfer_kernel, x);
}

// bytes per pixel column of every region, and the region offsets inside one (direction, parity) block
struct BandRegions {
    int nregions;              // R1 .. R(depth)
    int rows_c4[8], rows_v[8]; // rows of colour / variance in region r (r = 2..); region 1 = moments + history length
    size_t off[8];             // byte offset of region r inside a block
    size_t block_bytes;
};
BandRegions band_regions(int W, int depth) {
    BandRegions R{};
    size_t o = 0;
    R.off[1] = o; o += (size_t)W * kBandHistoryRows * (8 + 1);  // R1: moments (8 B) + history length (1 B)
    R.rows_c4[2] = kBandHistoryRows; R.rows_v[2] = 5;             // R2: level-0 output (history colour + L1 halo)
    R.off[2] = o; o += (size_t)W * (R.rows_c4[2] * 16 + R.rows_v[2] * 4);
    for (int l = 1; l <= depth - 2; ++l) {                        // R(l+2): output of level l, read by level l+1
        const int n = 2 * (2 << l) + 1;
        R.rows_c4[l + 2] = n; R.rows_v[l + 2] = n;
        R.off[l + 2] = o; o += (size_t)W * n * 20;
    }
    R.nregions = depth;
    R.block_bytes = (o + 255) & ~(size_t)255;
    return R;
}

}  // namespace

extern "C" int rmd_svgf_band_configure(rmd_svgf_ctx* c, int own_row0, int own_rows) {
    if (!c) return RMD_E_NULL;
    if (own_row0 < 0 || own_rows < 33 || own_row0 + own_rows > c->H) return RMD_E_SHAPE;
    if (c->W % 16) return RMD_E_UNSUPPORTED;
    if ((own_row0 != 0 && own_row0 < RMD_BAND_HALO) || (own_row0 + own_rows != c->H && c->H - own_row0 - own_rows < RMD_BAND_HALO))
        return RMD_E_SHAPE;
    // The flag words of a link carry a running sequence number: re-configuring a context that has already exchanged
    // frames would restart the sequence while the neighbours' words keep their old values (every wait would pass
    // immediately).  A band context is configured once.
    if (c->band_frame != 0 || c->band_rows != 0) return RMD_E_STATE;
    DeviceGuard guard(c->device);
    if (!c->band_counter) {
        RMD_CUDA_TRY(cudaMalloc((void**)&c->band_counter, 8));
        RMD_CUDA_TRY(cudaMemset(c->band_counter, 0, 8));
        RMD_CUDA_TRY(cudaHostAlloc((void**)&c->band_err_host, 4, cudaHostAllocMapped));
        *c->band_err_host = 0u;
        RMD_CUDA_TRY(cudaHostGetDevicePointer((void**)&c->band_err_dev, c->band_err_host, 0));
        // highest priority: the push's few CTAs must not queue behind the interior launch that follows it on the main
        // stream (a late push is a late neighbour)
        int prio_lo = 0, prio_hi = 0;
        RMD_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        RMD_CUDA_TRY(cudaStreamCreateWithPriority(&c->s_push, cudaStreamNonBlocking, prio_hi));
        RMD_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
        RMD_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pushed, cudaEventDisableTiming));
    }
    c->band_row0 = own_row0; c->band_rows = own_rows; c->band_frame = 0;
    return 0;
}

extern "C" int rmd_svgf_band_timeouts(const rmd_svgf_ctx* c) {
    if (!c) return RMD_E_NULL;
    return c->band_err_host ? (int)*(volatile unsigned int*)c->band_err_host : 0;
}

extern "C" int rmd_svgf_band_launch_count(const rmd_svgf_ctx* c) { return c ? c->band_launches : RMD_E_NULL; }

extern "C" size_t rmd_svgf_band_recv_bytes(const rmd_svgf_ctx* c) {
    return c ? 4 * band_regions(c->W, RMD_SVGF_MAX_LEVELS).block_bytes : 0;  // 2 directions x 2 frame parities
}

extern "C" int rmd_svgf_band_stage(rmd_svgf_ctx* c, const RmdSvgfFrame* f, const RmdFilterParams* fp,
                                   const RmdSvgfParams* sp, const RmdBandLink* link, int stage, void* stream) {
    int rc = check_frame(c, f, true);
    if (rc) return rc;
    if (!link || !link->recv || !link->flags) return RMD_E_NULL;
    if (c->band_rows == 0 || c->host_frames) return RMD_E_STATE;
    // sticky: once a flag wait has given up, the halo rows (and from then on the history) are not the neighbour's
    if (*(volatile unsigned int*)c->band_err_host) return RMD_E_TIMEOUT;
    c->device_frames = 1;
    SvgfConsts k;
    rc = resolve(fp, sp, &k);
    if (rc) return rc;
    if (k.depth < 2) return RMD_E_UNSUPPORTED;
    if (stage < 0 || stage > k.depth) return RMD_E_PARAM;
    DeviceGuard guard(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int W = c->W, E = c->H, Wp = c->Wp;
    const int o0 = c->band_row0, o1 = o0 + c->band_rows;
    const bool has[2] = {link->peer_recv[0] != nullptr, link->peer_recv[1] != nullptr};
    const BandRegions R = band_regions(W, RMD_SVGF_MAX_LEVELS);
    const int par = (int)(c->band_frame & 1ull);
    const unsigned long long seq0 = c->band_frame * 8ull;
    // my receive block for direction d (0 = from the upper neighbour) and the block I fill in neighbour d
    auto my_block = [&](int d) { return (char*)link->recv + (size_t)(2 * d + par) * R.block_bytes; };
    auto peer_block = [&](int d) { return (char*)link->peer_recv[d] + (size_t)(2 * (1 - d) + par) * R.block_bytes; };
    unsigned long long* my_flag[2] = {(unsigned long long*)link->flags, (unsigned long long*)link->flags + 1};

    // rows of my band that neighbour d needs (d = 0: my first rows, d = 1: my last rows) / my halo rows they fill
    auto src_row = [&](int d, int n) { return d == 0 ? o0 : o1 - n; };
    auto halo_row = [&](int d, int n) { return d == 0 ? o0 - n : o1; };

    // push: plane rows -> neighbour's receive block (region r), then bump its flag to seq0 + r
    // (`ordered`: the producer of the rows already ran on the push stream, no event needed)
    auto push = [&](int r, const void* p_a, int elem_a, int rows_a, const void* p_b, int elem_b, int rows_b, bool ordered = false) -> int {
        BandXfer x{};
        for (int d = 0; d < 2; ++d) {
            if (!has[d]) continue;
            char* dst = peer_block(d) + R.off[r];
            const struct { const void* p; int elem, rows; } pl[2] = {{p_a, elem_a, rows_a}, {p_b, elem_b, rows_b}};
            for (auto& q : pl) {
                if (!q.p || q.rows == 0) continue;
                RowJob& j = x.job[x.njobs++];
                j.src = (const char*)q.p + (size_t)src_row(d, q.rows) * Wp * q.elem;
                j.dst = dst;
                j.src_pitch = (unsigned)(Wp * q.elem); j.dst_pitch = (unsigned)(W * q.elem);
                j.row_units = (unsigned)(W * q.elem / 16); j.nrows = (unsigned)q.rows;
                dst += (size_t)W * q.elem * q.rows;
            }
            x.signal_flag[d] = (unsigned long long*)link->peer_flag[d];
        }
        x.value = seq0 + (unsigned long long)r;
        x.counter = c->band_counter + 1;
        x.err = c->band_err_dev;
        // the push runs on its own stream, after the boundary rows exist, beside whatever the main stream does next
        if (!ordered) {
            RMD_CUDA_TRY(cudaEventRecord(c->ev_boundary, s));
            RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_push, c->ev_boundary, 0));
        }
        const int rc2 = launch_xfer(x, c->s_push, false);  // first kernel after an event wait: nothing to overlap
        if (rc2) return rc2;
        RMD_CUDA_TRY(cudaEventRecord(c->ev_pushed, c->s_push));
        c->push_pending = 1;
        c->band_launches += 1;
        return 0;
    };
    // before the main stream overwrites rows a pending push may still be reading
    auto join_push = [&]() -> int {
        if (c->push_pending) {
            RMD_CUDA_TRY(cudaStreamWaitEvent(s, c->ev_pushed, 0));
            c->push_pending = 0;
        }
        return 0;
    };
    // unpack: wait for region r from both neighbours, copy it into my halo rows
    // (`early`: on the push stream, behind this band's own push of the same region — the neighbours push theirs at
    // about the same time — so the wait and the copy run beside the interior launch; few CTAs, they may spin)
    auto unpack = [&](int r, void* p_a, int elem_a, int rows_a, void* p_b, int elem_b, int rows_b, bool early = false) -> int {
        BandXfer x{};
        for (int d = 0; d < 2; ++d) {
            if (!has[d]) continue;
            const char* src = my_block(d) + R.off[r];
            const struct { void* p; int elem, rows; } pl[2] = {{p_a, elem_a, rows_a}, {p_b, elem_b, rows_b}};
            for (auto& q : pl) {
                if (!q.p || q.rows == 0) continue;
                RowJob& j = x.job[x.njobs++];
                j.src = src;
                j.dst = (char*)q.p + (size_t)halo_row(d, q.rows) * Wp * q.elem;
                j.src_pitch = (unsigned)(W * q.elem); j.dst_pitch = (unsigned)(Wp * q.elem);
                j.row_units = (unsigned)(W * q.elem / 16); j.nrows = (unsigned)q.rows;
                src += (size_t)W * q.elem * q.rows;
            }
            x.wait_flag[d] = my_flag[d];
        }
        x.value = seq0 + (unsigned long long)r;
        x.counter = c->band_counter;
        x.err = c->band_err_dev;
        c->band_launches += 1;
        if (early) {
            const int rc2 = launch_xfer(x, c->s_push, false, 148);
            if (rc2) return rc2;
            RMD_CUDA_TRY(cudaEventRecord(c->ev_pushed, c->s_push));  // the next stage joins after the unpack
            c->push_pending = 1;
            c->band_unpacked |= 1u << r;
            return 0;
        }
        return launch_xfer(x, s, (c->pdl & 1) != 0);
    };

    if (stage == 0) {
        c->band_launches = 0;
        c->parity ^= 1;
        const int cur = c->parity, prv = cur ^ 1;
        const int tb = o0 - kBandTemporalExt > 0 ? o0 - kBandTemporalExt : 0;
        const int te = o1 + kBandTemporalExt < E ? o1 + kBandTemporalExt : E;
        // one launch over every row of the context: full temporal pass on [tb, te), decoded guide only on the halo
        // rows outside it (the levels read the guide up to 33 rows out)
        TemporalArgs ta{};
        ta.color = (const uint2*)f->color; ta.albedo = (const uint32_t*)f->albedo;
        ta.guide = (const uint2*)f->guide; ta.motion = (const uint32_t*)f->motion;
        ta.hist_c4 = c->c4[kC4Hist]; ta.hist_m = c->m[prv]; ta.hist_n = c->n[prv]; ta.prev_g4 = c->g4[prv];
        ta.out_c4 = c->c4[kC4A]; ta.out_v = c->v[kC4A]; ta.out_m = c->m[cur]; ta.out_n = c->n[cur];
        ta.out_g4 = c->g4[cur]; ta.out_dz = c->dz; ta.side_c4 = c->side_c4;
        ta.tile_list = c->tile_list; ta.tile_count = c->tile_count + cur; ta.tile_capacity = c->tile_capacity;
        ta.W = W; ta.H = E; ta.Wp = Wp; ta.row_begin = 0; ta.row_end = E;
        ta.full_begin = tb; ta.full_end = te;
        // valid history = own rows + the kBandHistoryRows rows either neighbour refreshed after the last frame; a tap
        // beyond them (|motion_y| > RMD_BAND_MAX_MOTION_Y for an owned row) counts as disoccluded instead of reading
        // stale rows
        ta.hist_row_lo = has[0] ? o0 - kBandHistoryRows : 0;
        ta.hist_row_hi = has[1] ? o1 + kBandHistoryRows : E;
        ta.have_history = c->have_history; ta.k = k;
        rc = join_push(); if (rc) return rc;
        rc = launch_temporal(ta, s, (c->pdl & 2) != 0); if (rc) return rc;
        c->band_launches += 1;
        c->have_history = 1;
        return push(1, c->m[cur], 8, kBandHistoryRows, c->n[cur], 1, kBandHistoryRows);
    }
    const int cur = c->parity;
    if (stage == 1) {
        VarianceArgs va{};
        va.c4 = c->c4[kC4A]; va.m = c->m[cur]; va.n = c->n[cur]; va.g4 = c->g4[cur]; va.dz = c->dz;
        va.side_c4 = c->side_c4; va.patch_c4 = c->c4[kC4A]; va.patch_v = c->v[kC4A];
        va.tile_list = c->tile_list; va.tile_count = c->tile_count + cur; va.tile_capacity = c->tile_capacity; va.next_count = c->tile_count + (cur ^ 1);
        va.W = W; va.H = E; va.Wp = Wp; va.k = k;
        va.row_begin = o0 - kBandVarianceExt > 0 ? o0 - kBandVarianceExt : 0;
        va.row_end = o1 + kBandVarianceExt < E ? o1 + kBandVarianceExt : E;
        va.dense_min = c->var_dense_min; va.threads = c->var_threads; va.reverse = c->var_reverse;
        rc = launch_variance(va, s, (c->pdl & 4) != 0); if (rc) return rc;
        c->band_launches += 1;
    }
    const int l = stage - 1;  // a-trous level of this stage
    rc = join_push(); if (rc) return rc;  // the previous stage's push (issued before its interior launch) is long done
    if (l >= 1 && !(c->band_unpacked & (1u << (l + 1)))) {  // its input halo: the neighbours' output of level l-1
        const int in = level_in(l);
        rc = unpack(l + 1, c->c4[in], 16, R.rows_c4[l + 1], c->v[in], 4, R.rows_v[l + 1]);
        if (rc) return rc;
    }
    c->band_unpacked &= ~(1u << (l + 1));
    const bool last = l == k.depth - 1;
    AtrousArgs aa{};
    aa.in_c4 = c->c4[level_in(l)]; aa.in_v = c->v[level_in(l)];
    aa.g4 = c->g4[cur]; aa.dz = c->dz;
    const bool write_planes = !last || l == 0;
    aa.out_c4 = write_planes ? c->c4[level_out(l)] : nullptr;
    aa.out_v = write_planes ? c->v[level_out(l)] : nullptr;
    aa.final_out = last ? (float4*)f->out : nullptr;
    aa.final_rgba8 = last ? (uchar4*)f->out_rgba8 : nullptr;
    aa.albedo = (const uchar4*)f->albedo;
    aa.W = W; aa.H = E; aa.Wp = Wp; aa.Hp = c->Hp; aa.row0 = o0; aa.rows = c->band_rows;
    aa.sigma_z = k.sigma_z; aa.sigma_l = k.sigma_l; aa.sigma_n = k.sigma_n; aa.afloor = k.afloor;
    aa.use_tma = c->use_tma; aa.prefetch_ahead = c->atrous_prefetch;
    if (!last) {
        // Boundary rows first: the rows the neighbours' next level reads are produced and pushed before the
        // interior is computed, so the transfer (and a neighbour that is late by up to the interior's compute
        // time) is hidden behind the interior launch.
        const int out = level_out(l);
        const int nb = R.rows_c4[l + 2] > R.rows_v[l + 2] ? R.rows_c4[l + 2] : R.rows_v[l + 2];
        const bool split = (has[0] || has[1]) && 2 * nb + 8 * (1 << l) < c->band_rows;
        if (split) {
            // The level is split by TILES: first ONE launch over the tiles that hold a boundary row of either edge (a
            // few tile rows, not the whole plane), the push on the high-priority side stream, then the launch over all
            // other tiles.  No tile is evaluated twice.
            aa.row0 = o0; aa.rows = c->band_rows;
            aa.edge0[0] = o0; aa.edgeN[0] = has[0] ? nb : 0;
            aa.edge0[1] = o1 - nb; aa.edgeN[1] = has[1] ? nb : 0;
            aa.split = 1;
            if (c->band_edge_stream) {
                // the boundary tiles are a fraction of a wave of CTAs: on the (highest-priority) push stream they share
                // the GPU with the interior launch instead of having it to themselves; the next stage joins the push
                // stream before it reads or overwrites anything (join_push)
                RMD_CUDA_TRY(cudaEventRecord(c->ev_boundary, s));
                RMD_CUDA_TRY(cudaStreamWaitEvent(c->s_push, c->ev_boundary, 0));
                rc = launch_atrous(l, aa, c->maps[l][cur], c->s_push, c->variant[l], false); if (rc) return rc;
                rc = push(l + 2, c->c4[out], 16, R.rows_c4[l + 2], c->v[out], 4, R.rows_v[l + 2], true); if (rc) return rc;
                if (c->band_early_unpack) {
                    // the halo rows of this level's OUTPUT plane (the next level's input) are written by nobody else
                    rc = unpack(l + 2, c->c4[out], 16, R.rows_c4[l + 2], c->v[out], 4, R.rows_v[l + 2], true); if (rc) return rc;
                    if (l == 0) {
                        // next frame's history (the neighbours' moments / history length of this frame, pushed right
                        // after their temporal pass): nothing reads those halo rows once the variance pass is done
                        rc = unpack(1, c->m[cur], 8, kBandHistoryRows, c->n[cur], 1, kBandHistoryRows, true); if (rc) return rc;
                    }
                }
            } else {
                rc = launch_atrous(l, aa, c->maps[l][cur], s, c->variant[l], (c->pdl & 1) != 0); if (rc) return rc;
                rc = push(l + 2, c->c4[out], 16, R.rows_c4[l + 2], c->v[out], 4, R.rows_v[l + 2]); if (rc) return rc;
            }
            aa.split = 2;
            aa.reverse = c->band_serpentine && (l & 1) == 0;
            rc = launch_atrous(l, aa, c->maps[l][cur], s, c->variant[l], (c->pdl & 1) != 0); if (rc) return rc;
            c->band_launches += 2;
        } else {  // band too short to split (or no neighbours): one launch, then push
            aa.row0 = o0; aa.rows = c->band_rows;
            aa.reverse = c->band_serpentine && (l & 1) == 0;
            rc = launch_atrous(l, aa, c->maps[l][cur], s, c->variant[l], (c->pdl & 1) != 0); if (rc) return rc;
            c->band_launches += 1;
            if (has[0] || has[1]) {
                rc = push(l + 2, c->c4[out], 16, R.rows_c4[l + 2], c->v[out], 4, R.rows_v[l + 2]); if (rc) return rc;
            }
        }
    } else {
        aa.reverse = c->band_serpentine && (l & 1) == 0;
        rc = launch_atrous(l, aa, c->maps[l][cur], s, c->variant[l], (c->pdl & 1) != 0);
        if (rc) return rc;
        c->band_launches += 1;
        // history for the next frame: the neighbours' moments / history length of this frame
        if (!(c->band_unpacked & 2u)) {
            rc = unpack(1, c->m[cur], 8, kBandHistoryRows, c->n[cur], 1, kBandHistoryRows);
            if (rc) return rc;
        }
        c->band_unpacked = 0;
        c->band_frame++;
    }
    return 0;
}

extern "C" int rmd_svgf_band_frame(rmd_svgf_ctx* c, const RmdSvgfFrame* f, const RmdFilterParams* fp,
                                   const RmdSvgfParams* sp, const RmdBandLink* link, void* stream) {
    if (!fp) return RMD_E_NULL;
    for (int stage = 0; stage <= fp->depth; ++stage) {
        const int rc = rmd_svgf_band_stage(c, f, fp, sp, link, stage, stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int rmd_debug_level_cover(int width, int height, int level, int row0, int rows, int split, int edge0_a,
                                     int edge_n_a, int edge0_b, int edge_n_b, int reverse, int variant, int* cover,
                                     int* column_blocks, int* tiles) {
    if (!cover) return RMD_E_NULL;
    if (width <= 0 || height <= 0 || row0 < 0 || rows < 0 || row0 + rows > height) return RMD_E_SHAPE;
    if (level < 0 || level >= kMaxLevels || split < 0 || split > 2) return RMD_E_PARAM;
    if (variant < 0) variant = kAtrousDefaultVariant[level];
    if (!atrous_variant_exists(variant)) return RMD_E_PARAM;
    AtrousArgs aa{};
    aa.W = width; aa.H = height; aa.Wp = (width + 31) & ~31; aa.Hp = (height + 15) & ~15;
    aa.row0 = row0; aa.rows = rows; aa.split = split;
    aa.edge0[0] = edge0_a; aa.edgeN[0] = edge_n_a; aa.edge0[1] = edge0_b; aa.edgeN[1] = edge_n_b;
    aa.reverse = reverse;
    return atrous_cover(level, aa, variant, cover, column_blocks, tiles);
}

extern "C" const char* rmd_error_string(int code) {
    switch (code) {
        case RMD_OK: return "ok";
        case RMD_E_NULL: return "rmd: required pointer is null";
        case RMD_E_SHAPE: return "rmd: width/height out of range or not matching the context";
        case RMD_E_PARAM: return "rmd: parameter out of range";
        case RMD_E_ALIGN: return "rmd: plane base address not aligned";
        case RMD_E_UNSUPPORTED: return "rmd: FilterParams::type not provided by this entry point";
        case RMD_E_NOMEM: return "rmd: host allocation failed";
        case RMD_E_STATE: return "rmd: call not valid in this state";
        case RMD_E_DRIVER: return "rmd: cuTensorMapEncodeTiled unavailable or failed";
        case RMD_E_TIMEOUT: return "rmd: a neighbour's halo rows did not arrive (flag wait gave up); the band context is unusable";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rmd: unknown error";
    }
}
extern "C" int rmd_version(void) { return RMD_VERSION; }
extern "C" size_t rmd_sizeof_gbuffer(void) { return sizeof(RmdGBuffer); }
extern "C" size_t rmd_sizeof_filter_params(void) { return sizeof(RmdFilterParams); }
