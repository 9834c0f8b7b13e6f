// svgf.cuh — internal plane layout and pass launchers of the SVGF path.
//
// HBM layout (DESIGN.md "Data layout"): every context-owned plane is fp32 SoA with a
// row pitch of Wp = round_up(W, 32) texels and Hp = round_up(H, 16) rows; the
// padding rows/columns are zeroed once at creation and never written, so that a
// TMA box that runs past the image (zero-filled by hardware beyond the tensor,
// zero by construction inside the padding) always yields "invalid texel"
// (normal = 0 => normal weight 0) — the reference's border rule "skip the tap and
// renormalise" (reference src/filter.cu:38-39) with no branch in the tap loop.
#pragma once
#include "common.cuh"

namespace rmd {

constexpr int kMaxLevels = RMD_SVGF_MAX_LEVELS;
constexpr int kTemporalBx = 32, kTemporalBy = 8;  // temporal/variance CTA tile == flag tile

// resolved numeric parameters (FilterParams + SvgfParams with defaults applied)
struct SvgfConsts {
    float sigma_z, sigma_l, sigma_n;
    float alpha_c, alpha_m;
    int cap, short_hist;
    float dtol, nthr, afloor, lscale;
    int depth;
};

// a-trous tile geometry: a CTA produces WT columns x TY lattice rows of ONE phase
// (rows y = phase + step * k).  Each thread owns one column and 4 consecutive
// lattice rows, so a staged texel is reused for up to 4 outputs from registers.
constexpr int kAtrousWT = 128;  // output columns per CTA (= threads in x)
constexpr int kAtrousTR = 1;    // thread rows per CTA
constexpr int kAtrousOPT = 4;   // outputs per thread (consecutive lattice rows)
constexpr int kAtrousTY = kAtrousTR * kAtrousOPT;
// tile-kernel variants compiled into the library (svgf_atrous_tile.cu, -DRMD_VARIANT=n; build.py compiles the same list)
#ifndef RMD_ATROUS_VARIANTS
#define RMD_ATROUS_VARIANTS(X) X(0) X(1) X(3) X(6) X(7) X(8) X(11) X(12) X(15) X(16)
#endif
constexpr int kAtrousDefaultVariant[RMD_SVGF_MAX_LEVELS] = {16, 16, 16, 16, 16};  // per level, measured (csrc/svgf_atrous_tile.cu, profiles/r2_notes.md)

struct AtrousMaps {  // one set per (level, guide parity)
    // tile kernel: WT = atrous_variant_tile_width(variant, level), TW = WT + 2*max(2*step,4)
    CUtensorMap c4;  // 3-D {2W (8-byte elements), step, Hp/step}, box {TW (= TW/2 texels), 1, TY+4}; 2 boxes per tile
    CUtensorMap g4;  // same geometry on the decoded guide plane
    CUtensorMap v;   // 3-D {W, step, Hp/step} fp32, box {WT+2*max(2*step,4), 1, TY+4}
    CUtensorMap vn;  // same variance plane, box {WT+8, 1, TY}: the image rows y-1 / y+1 of the TY output rows (neighbour phases)
    CUtensorMap dzm; // slope plane, box {WT, 1, TY}
};

struct AtrousArgs {
    const float4* in_c4;
    const float* in_v;
    const float4* g4;
    const float* dz;
    float4* out_c4;  // may be null (last level, level > 0)
    float* out_v;
    float4* final_out;           // non-null on the last level
    uchar4* final_rgba8;         // optional
    const uchar4* albedo;        // caller plane, pitch W (last level only)
    int W, H, Wp, Hp;
    int row0;                    // first row this launch produces (band mode), else 0
    int rows;                    // number of rows produced
    // band mode splits a level by TILES: the tiles that hold a row of either edge range are one launch (split = 1,
    // produced first and pushed to the neighbours), all other tiles a second launch (split = 2); every launch stores
    // all rows of [row0, row0 + rows) that its tiles hold, so no tile is evaluated twice.  split = 0: every tile.
    int split;
    int edge0[2], edgeN[2];      // the two edge row ranges (edgeN = 0: none)
    int kt_lo[2], kt_cnt[2];     // lattice-tile index ranges the launch enumerates (filled by launch_atrous)
    float sigma_z, sigma_l, sigma_n, afloor;
    int use_tma;
    int prefetch_ahead;          // tiles of look-ahead of the L2 prefetch (variants with that mode bit; 0 = off, the default)
    int pf_dx, pf_dlo, pf_dhi;   // the same look-ahead as digits of the tile index (column block, low and high row digit)
    int reverse;                 // walk the tiles last to first (variants with the serpentine mode bit): the previous pass's
                                 // last-written rows are still in L2
};

struct TemporalArgs {
    const uint2* color;    // RGBA16F
    const uint32_t* albedo;
    const uint2* guide;
    const uint32_t* motion;  // RG16F
    const float4* hist_c4;
    const float2* hist_m;
    const uint8_t* hist_n;
    const float4* prev_g4;
    float4* out_c4;
    float* out_v;
    float2* out_m;
    uint8_t* out_n;
    float4* out_g4;
    float* out_dz;
    float4* side_c4;  // copy of out_c4 for short-history pixels (read by the variance pass)
    uint32_t* tile_list;   // compact list of the tiles that have short-history pixels: (tile x << 16) | first row
    uint32_t* tile_count;  // its length (appended with atomics; zeroed by the previous frame's variance pass)
    uint32_t tile_capacity;
    int W, H, Wp;
    int row_begin, row_end;  // rows this launch produces (whole plane unless the context is one band of a frame)
    int full_begin, full_end;      // rows that get the full temporal pass; the other rows of the launch only the guide decode
    int hist_row_lo, hist_row_hi;  // rows of the history planes that are valid ([0, H) unless the context is a band)
    int have_history;
    int prefetch_ctas;  // look-ahead of the L2 prefetch in CTAs (0 = off; set by launch_temporal)
    SvgfConsts k;
};

struct VarianceArgs {
    const float4* c4;  // temporal output (read for long-history neighbours, written in place for short ones)
    const float2* m;
    const uint8_t* n;
    const float4* g4;
    const float* dz;
    const float4* side_c4;  // untouched temporal colour of every short-history pixel
    float4* patch_c4;       // == c4
    float* patch_v;
    const uint32_t* tile_list;   // written by the temporal pass of this frame
    const uint32_t* tile_count;
    uint32_t tile_capacity;
    uint32_t* next_count;        // the counter the NEXT frame's temporal pass appends to: zeroed here
    int W, H, Wp;
    int row_begin, row_end;            // rows whose short-history pixels are re-estimated
    int dense_min;                     // qualifying pixels from which a tile is walked by position (two rows per thread)
    int threads;                       // CTA size: 128 (default) or 256
    int reverse;                       // walk the list last entry first: the tiles the temporal pass finished last are in L2
    SvgfConsts k;
};

int launch_temporal(const TemporalArgs& a, cudaStream_t s, bool pdl);
int launch_variance(const VarianceArgs& a, cudaStream_t s, bool pdl);
// independent tiles; `variant` selects one of the compiled kernel variants (RMD_ATROUS_VARIANTS), `pdl` launches with
// programmatic stream serialisation (the prologue overlaps the previous kernel's tail)
int launch_atrous(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s, int variant, bool pdl);
bool atrous_variant_exists(int variant);
// host enumeration of the tiles a launch would run (no device work): cover[y * nbx + bx] += 1 per stored (row, column block)
int atrous_cover(int level, const AtrousArgs& a, int variant, int* cover, int* nbx_out, int* tiles_with_work);
int atrous_variant_tile_width(int variant, int level);  // output columns per CTA (the TMA boxes are built for it)
int launch_atrous_ring(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s);  // persistent ring (4-row boxes)
int launch_guide_rows(const uint2* guide, float4* out_g4, int W, int H, int Wp, int row_begin, int row_end, cudaStream_t s);
int launch_remodulate(const float4* c4, const float* v, const float4* g4, const uchar4* albedo, float4* out,
                      uchar4* out8, int W, int H, int Wp, float afloor, cudaStream_t s);
int atrous_configure();  // opt-in dynamic shared memory, once per process/device

}  // namespace rmd
