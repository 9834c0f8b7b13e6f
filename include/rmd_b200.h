/*
 * rmd_b200.h — C ABI of the B200-native denoise path (librmd_b200.so).
 *
 * Drop-in boundary for the denoise path of VictorHerbert/RaymarchDenoiserCuda.
 * The reference has no FFI layer: its "API" is two __global__ functions that the
 * caller launches itself (reference include/filter.cuh:25-26, call sites
 * src/test.cu:73-75 and 85-87) over two POD structs passed by value
 * (include/filter.cuh:11-23, include/gbuffer.h:6-14).  This header declares the
 * extern "C" entry points a maintainer binds instead of those launches; every
 * struct that crosses the boundary is plain C, bit-compatible with the reference
 * struct it mirrors (static_asserts in csrc/svgf_ctx.cu, probes in tests/test_abi.py).
 * Test and inspection hooks (stage stop, plane read-back, launch counters) are declared in
 * rmd_b200_debug.h, not here.
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t value (> 0) for CUDA
 *     failures, or a negative RMD_E_* code for argument errors; nothing throws.
 *   - `stream` is a cudaStream_t passed as void*; all device work is enqueued on
 *     it and nothing synchronises the device unless the name says "_sync"/"_host".
 *   - the library never takes ownership of caller planes (reference convention:
 *     GBuffer is a raw-pointer view, include/gbuffer.h:6-14).
 *   - threading: the stateless entry points (rmd_filter_*) are re-entrant; an
 *     rmd_svgf_ctx holds one sequence's history and must be driven by one thread at a
 *     time, different contexts (same or different devices) may be used concurrently
 *     from different threads.  The reference is single-threaded throughout
 *     (src/test.cu:17-48).
 */
#ifndef RMD_B200_H
#define RMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMD_VERSION 100

/* ---- error codes (negative = argument/host errors, positive = cudaError_t) ---- */
#define RMD_OK 0
#define RMD_E_NULL (-1)        /* required pointer is null                          */
#define RMD_E_SHAPE (-2)       /* width/height out of range or mismatching the ctx  */
#define RMD_E_PARAM (-3)       /* radius/depth/type/sigma out of range              */
#define RMD_E_ALIGN (-4)       /* plane base not aligned as documented              */
#define RMD_E_UNSUPPORTED (-5) /* valid in the reference API but not provided here  */
#define RMD_E_NOMEM (-6)       /* host allocation failed                            */
#define RMD_E_STATE (-7)       /* call order (e.g. band exchange without peers)     */
#define RMD_E_DRIVER (-8)      /* cuTensorMapEncodeTiled unavailable / failed       */
#define RMD_E_TIMEOUT (-9)     /* band mode: a neighbour's rows never arrived (sticky) */

/* ------------------------------------------------------------------------------
 * Reference-mirroring PODs
 * ---------------------------------------------------------------------------- */

/* Mirrors `struct GBuffer` (reference include/gbuffer.h:6-14): int2 shape followed
 * by six device pointers to row-major uchar4 planes with pitch = 4*W bytes
 * (include/extended_math.h:66-68).  sizeof == 56. */
typedef struct RmdGBuffer {
    int32_t width, height; /* int2 shape {x = W, y = H} */
    void* render;          /* uchar4[W*H], read-only input           */
    void* denoised;        /* uchar4[W*H], output                    */
    void* normal;          /* uchar4[W*H], unused by the box path    */
    void* albedo;          /* uchar4[W*H], unused by the box path    */
    void* buffer[2];       /* uchar4[W*H] ping-pong, needed iff depth > 1 */
} RmdGBuffer;

/* Mirrors `struct FilterParams` (reference include/filter.cuh:11-23). sizeof == 36. */
enum { RMD_FILTER_AVERAGE = 0, RMD_FILTER_GAUSSIAN = 1, RMD_FILTER_CROSS = 2, RMD_FILTER_WAVELET = 3 };
typedef struct RmdFilterParams {
    int32_t type;      /* RMD_FILTER_*; AVERAGE for the box path, WAVELET for SVGF    */
    int32_t depth;     /* number of levels                                             */
    int32_t level;     /* unused by the reference; ignored                             */
    int32_t radius;    /* tap radius; box path: 1..RMD_BOX_MAX_RADIUS, SVGF: must be 2 */
    float sigmaSpace;  /* SVGF: sigma_z (0 => 1)                                       */
    float sigmaColor;  /* SVGF: sigma_l (0 => 4)                                       */
    float sigmaAlbedo; /* unused (albedo is demodulated, not an edge-stopping term)    */
    float sigmaNormal; /* SVGF: sigma_n (0 => 128)                                     */
    uint8_t cacheInput;  /* reference default true; accepted, never changes results   */
    uint8_t cacheBuffer; /* reference default true; accepted, never changes results   */
} RmdFilterParams;

#define RMD_BOX_MAX_RADIUS 32
#define RMD_SVGF_MAX_LEVELS 5

/* ------------------------------------------------------------------------------
 * Legacy box path — replaces the caller-side launches of
 *   filterKernelBaseline (reference src/filter.cu:13-58,  call src/test.cu:73-75)
 *   filterKernelTiled    (reference src/filter.cu:87-158, call src/test.cu:85-87)
 * Semantics are those of the reference with levels iterated as separate launches
 * (the reference's in-kernel depth>1 loop races across blocks, src/filter.cu:56):
 *   level l reads  (l == 0 ? render : buffer[l % 2])
 *   level l writes (l == depth-1 ? denoised : buffer[(l + 1) % 2])
 *   out = trunc(sum over in-image taps / count of in-image taps)
 * baseline: the R channel is replicated into R,G,B (src/filter.cu:51-53); the
 *           reference leaves .w undefined, this library writes 0.
 * tiled:    R,G,B averaged independently, .w = 0 (src/filter.cu:151-155).  The
 *           reference's cacheInput=true shared-memory path mis-strides its tile
 *           (src/filter.cu:66-67 vs 97,130); this library always returns what the
 *           cacheInput=false path returns.
 * ---------------------------------------------------------------------------- */
int rmd_filter_baseline(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream);
int rmd_filter_tiled(const RmdGBuffer* frame, const RmdFilterParams* params, void* stream);

/* ------------------------------------------------------------------------------
 * SVGF path (temporal accumulation -> variance -> depth-level a-trous), the path
 * the reference announces (README.md:3-10; FilterParams::WAVELET, level, sigma*
 * at include/filter.cuh:12-19; waveletSpline at src/filter.cu:10) but does not
 * implement.  Normative arithmetic: DESIGN.md "SVGF specification".
 *
 * Input planes (device memory, row-major, pitch = W * texel size, base 16-B aligned):
 *   color   RGBA16F  (8 B)  linear radiance, 1 spp; .a ignored
 *   albedo  RGBA8    (4 B)  albedo = rgb / 255
 *   guide   2 x u32  (8 B)  word0 = octahedral unit normal, snorm16 x | snorm16 y << 16
 *                           word1 = linear view depth z as fp32 bits; z <= 0 or
 *                           non-finite marks sky (passed through unfiltered)
 *   motion  RG16F    (4 B)  screen-space motion in pixels, prev = p + motion
 * Output plane:
 *   out     RGBA32F (16 B)  rgb = denoised linear radiance, .a = filtered variance of
 *                           the demodulated luminance after the last level
 *   out_rgba8 (optional, may be null) uchar4: the reference's `denoised` format
 *                           (include/gbuffer.h:10): clamp(rgb,0,1)*255 truncated,
 *                           .w = 255
 * ---------------------------------------------------------------------------- */
typedef struct RmdSvgfFrame {
    int32_t width, height;
    const void* color;
    const void* albedo;
    const void* guide;
    const void* motion;
    void* out;
    void* out_rgba8;
} RmdSvgfFrame;

/* Constants the reference API has no field for.  Zero in any field selects the
 * default in brackets (same rule as the sigma fields, which the reference
 * zero-initialises). */
typedef struct RmdSvgfParams {
    float alpha_color;      /* [0.05] floor of the colour blend factor               */
    float alpha_moments;    /* [0.2]  floor of the moment blend factor               */
    int32_t history_cap;    /* [32]   maximum history length N'                      */
    int32_t short_history;  /* [4]    N' below this takes the 7x7 spatial variance   */
    float depth_tolerance;  /* [0.1]  reprojection: |dz| <= tol*z + 2*slope          */
    float normal_threshold; /* [0.9]  reprojection: n_prev . n >= threshold          */
    float albedo_floor;     /* [1e-3] demodulation floor                             */
    float variance_lum_scale; /* [10] fixed luminance scale of the 7x7 variance pass */
} RmdSvgfParams;

typedef struct rmd_svgf_ctx rmd_svgf_ctx; /* opaque per-sequence state (history planes) */

/* Creates the per-sequence context on `device` (cudaSetDevice is called inside and
 * the previous device restored).  Allocates every internal plane, so rmd_svgf_frame
 * never allocates.  The staging buffers of the two convenience entry points
 * (rmd_svgf_frame_host: 2 x 44 B/px + streams; rmd_svgf_frame_gbuffer: 36 B/px) are
 * allocated by their first call, or ahead of time by rmd_svgf_prepare_host /
 * rmd_svgf_prepare_gbuffer.
 * One context serves ONE entry-point family: the device-pointer calls
 * (rmd_svgf_frame, _frame_gbuffer, _band_*) run on the caller's stream, the host-frame
 * call on the context's own streams, and both update the same history planes; mixing
 * them on one context returns RMD_E_STATE. */
int rmd_svgf_create(rmd_svgf_ctx** ctx, int width, int height, int device);
int rmd_svgf_prepare_host(rmd_svgf_ctx* ctx);
int rmd_svgf_prepare_gbuffer(rmd_svgf_ctx* ctx, void* stream);
int rmd_svgf_destroy(rmd_svgf_ctx* ctx);
/* Drops the temporal history (next frame is treated as fully disoccluded). */
int rmd_svgf_reset(rmd_svgf_ctx* ctx);
/* One frame: temporal + variance + params->depth a-trous levels, all on `stream`.  No allocation, no
 * synchronisation: the call may be captured into a CUDA graph (stream capture).  The context alternates its
 * ping-pong planes by frame parity, so capture an EVEN number of consecutive frames per graph, after at least one
 * eager frame (the history flag is baked in at capture time); tests/test_gpu_svgf.py shows the pattern. */
int rmd_svgf_frame(rmd_svgf_ctx* ctx, const RmdSvgfFrame* frame, const RmdFilterParams* params,
                   const RmdSvgfParams* svgf, void* stream);
/* Same frame with HOST planes (pinned or pageable): H2D of the four inputs, the
 * frame, D2H of `out` (and out_rgba8 when non-null).  `out` may be null when
 * out_rgba8 is given: only the reference's `denoised` format (4 B/px instead of 16)
 * crosses PCIe on the way back.  Copies run on the context's
 * copy streams so that the upload of frame f+1 and the download of frame f-1
 * overlap the kernels of frame f; rmd_svgf_host_wait blocks until every
 * submitted frame has landed in its host destination. */
int rmd_svgf_frame_host(rmd_svgf_ctx* ctx, const RmdSvgfFrame* host_frame, const RmdFilterParams* params,
                        const RmdSvgfParams* svgf);
int rmd_svgf_host_wait(rmd_svgf_ctx* ctx);

/* The same frame on the reference's own G-buffer format (BASELINE configs[0]): `frame` is the reference's
 * `struct GBuffer` (include/gbuffer.h:6-14) — RGBA8 render / albedo / normal planes in, RGBA8 `denoised` out;
 * this is the call `filterKernel*(GBuffer, FilterParams{.type = WAVELET, ...})` was reserved for
 * (include/filter.cuh:12-19).  One kernel converts on the device, with individually rounded fp32 operations:
 *   radiance = render.rgb / 255 (linear, as stored);  albedo as is;
 *   normal   = normalise(normal.rgb / 255), a zero vector becoming (0,0,1) — the reference's fixture stores
 *              un-biased normals whose negative components are clamped away;
 *   depth    = 1 everywhere and motion = 0: the reference's GBuffer has neither plane.
 * With zero motion the history accumulates over repeated calls (static camera); rmd_svgf_reset() starts over.
 * `out_rgba32f` (optional) receives the linear fp32 result as well.  frame->buffer[] is not used. */
int rmd_svgf_frame_gbuffer(rmd_svgf_ctx* ctx, const RmdGBuffer* frame, const RmdFilterParams* params,
                           const RmdSvgfParams* svgf, void* out_rgba32f, void* stream);

/* Per-pass GPU timing with CUDA events recorded on the frame's own stream between the
 * passes (the reference times whole tests with std::chrono, src/test.cu:33-38).
 * rmd_svgf_get_pass_times synchronises on the last frame's final event and writes
 * up to `capacity` durations in milliseconds, in launch order:
 *   [0] temporal, [1] variance (estimate + patch), [2 + l] a-trous level l
 * (or [2] = remodulate when depth == 0); returns the number of entries, < 0 on error. */
int rmd_svgf_set_profiling(rmd_svgf_ctx* ctx, int enable);
int rmd_svgf_get_pass_times(rmd_svgf_ctx* ctx, float* ms, int capacity);

/* ------------------------------------------------------------------------------
 * History rows in/out — the state a sequence carries from frame to frame (colour
 * history = level-0 output, luminance moments, history length, decoded guide),
 * for rows [row_begin, row_begin + nrows) of the context, packed tightly (pitch W):
 *   [float4 colour][float4 guide][float2 moments][uint8 length]  = 41 bytes per pixel
 * Used for row-band partitioning of one large frame over several GPUs (no
 * reference counterpart: the reference is single-GPU): every rank
 * runs an ordinary context over its band plus a halo and, after each frame,
 * overwrites the history rows of its halo with the owning neighbour's rows
 * (raymarchdenoisercuda_b200/shard.py).  Also usable as a
 * checkpoint/restore of a sequence.  `buf` is DEVICE memory on the context's device.
 * ---------------------------------------------------------------------------- */
size_t rmd_svgf_history_bytes(const rmd_svgf_ctx* ctx, int nrows);
int rmd_svgf_history_pack(rmd_svgf_ctx* ctx, int row_begin, int nrows, void* buf, void* stream);
int rmd_svgf_history_unpack(rmd_svgf_ctx* ctx, int row_begin, int nrows, const void* buf, void* stream);

/* ------------------------------------------------------------------------------
 * NVLink peer-to-peer plumbing between ranks (one process per GPU) for the banded
 * mode: device buffers that other processes can map (CUDA IPC) and stream-ordered
 * flags.  A rank packs its boundary history rows straight into the neighbour's
 * buffer (rmd_svgf_history_pack with a peer-mapped `buf`), then rmd_p2p_signal()s
 * the neighbour's flag; the neighbour's stream rmd_p2p_wait()s on its own flag
 * before unpacking.  No collective, no host synchronisation per frame.
 * ---------------------------------------------------------------------------- */
#define RMD_IPC_HANDLE_BYTES 64
int rmd_p2p_alloc(void** dev_ptr, size_t bytes);              /* zero-initialised, IPC-exportable device memory */
int rmd_p2p_free(void* dev_ptr);
int rmd_p2p_export(void* dev_ptr, void* handle_out);          /* RMD_IPC_HANDLE_BYTES bytes                      */
int rmd_p2p_open(const void* handle, void** peer_ptr);        /* maps another process's buffer (enables peer access) */
int rmd_p2p_close(void* peer_ptr);
/* `flag` points to an 8-byte word.  signal: after all prior work of `stream`, store `value` (system scope).
 * wait: hold `stream` until the word is >= value (bounded: gives up after ~2 s, bumps rmd_p2p_timeouts(), and every later
 * rmd_p2p_wait of the process returns RMD_E_TIMEOUT instead of letting the caller unpack rows that never arrived). */
int rmd_p2p_signal(void* flag, unsigned long long value, void* stream);
int rmd_p2p_wait(const void* flag, unsigned long long value, void* stream);  /* RMD_E_TIMEOUT once any earlier wait gave up (sticky) */
int rmd_p2p_timeouts(void);                                   /* number of waits that gave up so far (host-visible word, no device sync) */

/* ------------------------------------------------------------------------------
 * Row bands with per-level halo exchange (north-star item 4; no reference counterpart).
 * The context covers the band's own rows plus RMD_BAND_HALO halo rows on each interior side.  The a-trous levels
 * produce ONLY the own rows; after each level the rank pushes the boundary rows the neighbour's next level reads
 * (2*2^(l+1)+1 rows of colour+variance) straight into the neighbour's receive buffer over NVLink and bumps its
 * flag; the neighbour's stream waits on the flag, copies the rows into its halo and runs the next level.  The
 * temporal and variance passes (15 % of the frame) simply run 6 / 3 rows into the halo instead of exchanging.
 * History for the next frame (moments, history length, level-0 colour: 21 rows) travels the same way.
 * A frame is depth+1 stages; rmd_svgf_band_frame runs them in order.  Inside a stage the boundary tiles of the level,
 * the push and the unpack of the neighbours' rows of the same level run on the context's own highest-priority stream
 * beside the interior launch on `stream`; the next stage joins that stream before it reads or overwrites anything.
 * When several bands live in one process on one GPU (tests), call rmd_svgf_band_stage s for every band before stage
 * s+1: a band's unpack (148 CTAs) spins on the device until the neighbour's stage s has been enqueued and has run,
 * so the neighbour's stage s must follow without a host-side wait in between.
 * Requires width % 16 == 0, depth >= 2, own_rows >= 33.  A context is configured once.
 *
 * Motion limit.  The history rows a band can reproject from are its own rows plus the 21 rows either neighbour
 * refreshes after every frame.  Results equal the single-GPU frame bit for bit while |motion_y| <=
 * RMD_BAND_MAX_MOTION_Y for every pixel within 27 rows of a band edge; a reprojection tap that falls beyond the
 * refreshed rows is treated as outside the image (that pixel is disoccluded: history length restarts at 1) — defined,
 * but no longer what one GPU would compute.  Horizontal motion is unrestricted.
 *
 * Failure.  Every wait on a neighbour's flag is bounded (~2 s).  A wait that gives up bumps a host-visible word;
 * from then on every rmd_svgf_band_* call on the context returns RMD_E_TIMEOUT (rmd_svgf_band_timeouts() reads the
 * count without synchronising the device).
 * ---------------------------------------------------------------------------- */
#define RMD_BAND_HALO 40
#define RMD_BAND_MAX_MOTION_Y 13
typedef struct RmdBandLink {
    void* peer_recv[2]; /* [0] upper / [1] lower neighbour's receive buffer base (peer-mapped); null at the image border */
    void* peer_flag[2]; /* the flag word in that neighbour that THIS rank bumps                                      */
    void* recv;         /* this rank's receive buffer, rmd_svgf_band_recv_bytes() bytes (rmd_p2p_alloc)              */
    void* flags;        /* this rank's two flag words: [0] bumped by the upper, [1] by the lower neighbour           */
} RmdBandLink;
int rmd_svgf_band_configure(rmd_svgf_ctx* ctx, int own_row0, int own_rows); /* rows of the context that are owned */
size_t rmd_svgf_band_recv_bytes(const rmd_svgf_ctx* ctx);
int rmd_svgf_band_stage(rmd_svgf_ctx* ctx, const RmdSvgfFrame* frame, const RmdFilterParams* params,
                        const RmdSvgfParams* svgf, const RmdBandLink* link, int stage, void* stream);
int rmd_svgf_band_frame(rmd_svgf_ctx* ctx, const RmdSvgfFrame* frame, const RmdFilterParams* params,
                        const RmdSvgfParams* svgf, const RmdBandLink* link, void* stream); /* stages 0..depth */
int rmd_svgf_band_timeouts(const rmd_svgf_ctx* ctx); /* flag waits that gave up so far (no device sync) */

/* ------------------------------------------------------------------------------ */
const char* rmd_error_string(int code);
int rmd_version(void);
/* sizeof probes used by the layout tests */
size_t rmd_sizeof_gbuffer(void);
size_t rmd_sizeof_filter_params(void);

#ifdef __cplusplus
}
#endif
#endif /* RMD_B200_H */
