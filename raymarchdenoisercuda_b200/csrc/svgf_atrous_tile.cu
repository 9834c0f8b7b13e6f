// svgf_atrous_tile.cu — passes 3..7 of the SVGF path, independent-tile kernel: one edge-avoiding
// a-trous wavelet level (5x5 B3-spline taps dilated by step = 2^level; normal, depth and luminance
// edge-stopping weights; variance propagation), DESIGN.md spec S4-S5, checked against
// oracle/oracle_svgf.c:pass_atrous.
//
// Reference hooks: the global->shared halo tile that the reference fills with a strided
// cooperative copy (`cacheTile`, src/filter.cu:60-85) is filled here by TMA tensor loads; taps and
// border rule: svgf_atrous.cuh.
//
// Tiling (DESIGN.md "A-trous kernel"):
//   * Polyphase rows.  At step S a pixel only ever reads rows with the same (y mod S), so a CTA
//     works on ONE row phase: its tile is WT dense columns x TY lattice rows (y = phase + S*k).
//     The planes are described to TMA as {x, phase, k} tensors (strides pitch, S*pitch), so one
//     box fetches the (WT + 2*max(2S,4)) x (TY + 4) texels the tile needs: the vertical halo is
//     2 lattice rows at every level instead of 2*S image rows.
//   * TMA zero-fills texels outside the image (and the planes' padding rows are zero), which
//     decodes to "normal = 0": the normal weight max(0, n.n')^sigma is then exactly 0, i.e. the
//     tap is skipped and the sum renormalised, with no bounds test in the tap loop.
//   * Centre terms from shared memory (MODE & 1).  The 3x3 variance pre-filter of an output needs
//     the image rows y-1 and y+1, which belong to the neighbouring row phases and are therefore
//     not in the tile: two more TMA boxes on the variance map (phase -/+ 1, wrapping to the
//     previous / next lattice row at the phase ends; at step 1 the rows are already in the tile)
//     and one on the slope map deliver them, so the prologue has no global load, no 64-bit address
//     arithmetic and a single wait.  (Round 1 issued 40 `ld.global.nc` per thread here: a quarter
//     of the kernel's instructions and half of its warp-stall samples, profiles/r2_notes.md.)
//   * Register blocking.  Each thread owns one column and 4 consecutive lattice rows; the 8x5
//     texels it stages through registers feed 100 taps (2.5 taps per shared-memory load).
//     Lanes are consecutive in x, so every LDS.128 is conflict-free.
//   * MODE & 2: the 5 tap columns are walked grouped by |dx| (2 x 20 + 2 x 20 + 16 taps in three
//     unrolled bodies) instead of as one 100-tap body: 43 KB of code becomes ~25 KB.
//   * MODE & 8 / & 64: tile order.  blockIdx enumerates (lattice tile row, row phase, column block), so
//     the S phases of 4*S consecutive image rows are resident together and the centre-term loads of a
//     tile (rows y-1 / y+1 = the neighbouring phases) hit L2; launches flagged `reverse` walk the same
//     order backwards (levels alternate: each starts on the rows its predecessor wrote last).
//   * MODE & 32: in the centre column the taps between two outputs of the same thread share their
//     dot product, lg2, |dz| and |dL| (tile_centre_column_pairs); bit-identical to the plain walk.
//
// This file is compiled once per variant (-DRMD_VARIANT=n, build.py); svgf_atrous.cu dispatches.
#include "svgf_atrous.cuh"

#ifndef RMD_VARIANT
#define RMD_VARIANT 1
#endif

namespace rmd {
namespace {

// kMode: bit 0 = centre terms staged by TMA, bit 1 = |dx|-grouped tap body, bit 2 = persistent CTAs,
//        bit 3 = tiles enumerated lattice-row-major (the S row phases of a lattice tile row are consecutive CTAs),
//        bit 4 = thread 0 prefetches the boxes of a look-ahead tile into L2, bit 5 = centre column with the
//        pair terms of two outputs of the same thread evaluated once, bit 6 = serpentine: a launch with a.reverse
//        walks its tiles last to first (levels alternate, so each starts on the rows the previous pass wrote last)
// kMinB: resident CTAs per SM the register allocation is bounded for (4 -> 128 registers, 5 -> 96)
// Measured on B200 (profiles/r2_notes.md, us per level at 1080p / 4K, steps 1,2,4,8,16):
//   0 legacy                      55.2 55.1 55.5 55.5 62.4 / 184 186 187 187 213
//   1 TMA centre terms            53.2 57.6 58.2 59.1 71.5 / 177 197 197 199 249   (helps at step 1 only: no extra boxes there)
//   3 grouped                     54.5 53.4 54.1 54.4 62.1 / 181 182 182 183 211
//   6 grouped, 5 CTAs/SM          52.4 52.6 52.5 54.5 63.6 / 176 178 178 185 213   (5 CTAs fit for steps <= 4)
//   7 TMA centre, grouped, 5 CTAs 51.7 55.3 54.3 54.9 73.4 / 174 185 182 183 254
//   8 grouped, persistent CTAs    56.6 56.7 56.6 56.9 65.8 / 196 197 196 198 228   (kept as the measured negative result)
//   (7 was 2 us ahead at step 1 until the band-mode arguments were added to the kernel; at the 96-register cap it
//   now spills 8 bytes there and measures equal: 179.8 vs 179.6 us at 4K.)
// Later in round 2 (profiles/r2_atrous_variants_c.jsonl / _d.jsonl, 4K, frame us and level us; all bit-identical to 6):
//   6                                        1195-1200   178.5 178.7 179.0 187.3 214.3
//   11 lattice-row-major, prefetch off / 740 1197 / 1207 (the L2 tensor prefetch costs TMA requests: worse at any look-ahead)
//   12 pair terms                            1190        177.6 175.8 175.9 185.0 212.3
//   15 row-major + serpentine, prefetch off  1183        177.2 175.1 175.1 181.3 214.0
//   16 row-major + serpentine + pair terms   1166        176.0 171.5 171.1 179.0 212.0   (prefetch off)
//   the same mode compiled WITHOUT the (unused) prefetch code: 1211 us, step 16 at 230 us — ptxas keeps the tile
//   position in vector instead of uniform registers, spills 12 bytes and schedules the loop worse; a 128-register
//   build for steps 8 / 16 (4 CTAs/SM fit anyway): 1182 us.  The 96-register allocation is that fragile.
// Shipped: 16 at every step, look-ahead 0 (RMD_ATROUS_PREFETCH=n switches the prefetch on).
#if RMD_VARIANT == 0
constexpr int kMode = 0, kMinB = 4;
#elif RMD_VARIANT == 1
constexpr int kMode = 1, kMinB = 4;
#elif RMD_VARIANT == 3
constexpr int kMode = 2, kMinB = 4;
#elif RMD_VARIANT == 6
constexpr int kMode = 2, kMinB = 5;
#elif RMD_VARIANT == 7
constexpr int kMode = 3, kMinB = 5;
#elif RMD_VARIANT == 8
constexpr int kMode = 6, kMinB = 4;
#elif RMD_VARIANT == 11
constexpr int kMode = 2 | 8 | 16, kMinB = 5;
#elif RMD_VARIANT == 12
constexpr int kMode = 2 | 32, kMinB = 5;
#elif RMD_VARIANT == 15
constexpr int kMode = 2 | 8 | 16 | 64, kMinB = 5;
#elif RMD_VARIANT == 16
constexpr int kMode = 2 | 8 | 16 | 32 | 64, kMinB = 5;
#else
#error "unknown RMD_VARIANT"
#endif
// output columns per CTA (= threads) at step S.  192-column tiles at steps 8 / 16 (x-halo amplification 1.17 / 1.33
// instead of 1.25 / 1.5, 3 CTAs of 6 warps per SM) were measured within 1-2 % of 128 columns and dropped.
constexpr int tile_wt(int) { return kAtrousWT; }
constexpr int tile_minb(int) { return kMinB; }

int g_sms = 148, g_smem_per_sm = 233472;  // set by atrous_tile_configure (same for every GPU of the box)

template <int S, int MODE>
struct Tile {
    static constexpr int WT = tile_wt(S);
    static constexpr bool PRO = (MODE & 1) != 0;
    static constexpr bool GROUPED = (MODE & 2) != 0;
    static constexpr bool PERSIST = (MODE & 4) != 0;
    static constexpr bool KT_MAJOR = (MODE & 8) != 0;
    static constexpr bool PREFETCH = (MODE & 16) != 0;
    static constexpr bool PAIRS = (MODE & 32) != 0;
    static constexpr bool SERPENTINE = (MODE & 64) != 0;
    static_assert(!(PERSIST && SERPENTINE), "the persistent walk is forward only");
    // x halo: 2*S texels are needed; TMA wants every box row to start on a 16-byte
    // boundary, and the variance plane has 4-byte texels, so the halo is a multiple of 4.
    static constexpr int HX = 2 * S < 4 ? 4 : 2 * S;
    static constexpr int TW = WT + 2 * HX;
    static constexpr int TH = kAtrousTY + 4;
    // float4 planes are staged as two half-width column blocks [2][TH][TW/2]: a TMA box
    // dimension holds at most 256 elements, so one box of 8-byte elements covers TW/2
    // texels (<= 96) per row with 1-1.5 KB rows (16-byte inner rows made TMA request-bound).
    static constexpr int HW2 = TW / 2;
    static constexpr int HALF_BYTES = HW2 * TH * 16;
    static constexpr int C4_BYTES = 2 * HALF_BYTES;
    static constexpr int V_BYTES = TW * TH * 4;
    // neighbour-phase variance rows (y-1 / y+1 of the TY output rows), columns x0-4 .. x0+WT+3
    static constexpr int NBW = WT + 8;
    static constexpr int NB_BYTES = (PRO && S > 1) ? NBW * kAtrousTY * 4 : 0;
    static constexpr int DZ_BYTES = PRO ? WT * kAtrousTY * 4 : 0;
    static constexpr int OFF_C4 = 0;
    static constexpr int OFF_G4 = align128(C4_BYTES);
    static constexpr int OFF_V = OFF_G4 + align128(C4_BYTES);
    static constexpr int OFF_VM = OFF_V + align128(V_BYTES);
    static constexpr int OFF_VP = OFF_VM + align128(NB_BYTES);
    static constexpr int OFF_DZ = OFF_VP + align128(NB_BYTES);
    static constexpr int OFF_BAR = OFF_DZ + align128(DZ_BYTES);
    static constexpr int SMEM = OFF_BAR + 16 + 128;  // + slack to align the dynamic base to 128 B
    static constexpr uint32_t TX_BYTES = 2u * C4_BYTES + V_BYTES + 2u * NB_BYTES + DZ_BYTES;
    static_assert(HALF_BYTES % 128 == 0, "second column block must stay 128-B aligned for TMA");
    static_assert(TW % 2 == 0 && 2 * HW2 <= 256, "box limit");
    static_assert((NBW * 4) % 16 == 0 && (WT * 4) % 16 == 0 && WT % 32 == 0, "TMA inner box bytes");
    // texel offset of column `col` inside a float4 plane (row 0)
    __device__ static __forceinline__ int coloff(int col) { return col < HW2 ? col : TH * HW2 + col - HW2; }
};

// One column of the 5x5 footprint for the kAtrousOPT outputs of a thread: TH staged texels, all
// addresses [register + immediate]; loads run one texel ahead of the arithmetic that consumes them.
template <class T, int ADX>
__device__ __forceinline__ void tile_column(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const uint32_t b,
                                            const uint32_t vb, const float sigma_n) {
    static_assert(T::TH == 8, "unrolled for 4 outputs per thread");
#define RMD_TILE_LOAD(JR, Q, G, VV)                                   \
    const float4 Q = lds128<T::OFF_C4 + (JR)*T::HW2 * 16>(b);         \
    const float4 G = lds128<T::OFF_G4 + (JR)*T::HW2 * 16>(b);         \
    const float VV = lds32<T::OFF_V + (JR)*T::TW * 4>(vb);
    RMD_TILE_LOAD(0, q0, g0, w0)
    RMD_TILE_LOAD(1, q1, g1, w1)
    taps_of_texel_adx<ADX, 0>(acc, ctr, q0, g0, w0, sigma_n);
    RMD_TILE_LOAD(2, q2, g2, w2)
    taps_of_texel_adx<ADX, 1>(acc, ctr, q1, g1, w1, sigma_n);
    RMD_TILE_LOAD(3, q3, g3, w3)
    taps_of_texel_adx<ADX, 2>(acc, ctr, q2, g2, w2, sigma_n);
    RMD_TILE_LOAD(4, q4, g4, w4)
    taps_of_texel_adx<ADX, 3>(acc, ctr, q3, g3, w3, sigma_n);
    RMD_TILE_LOAD(5, q5, g5, w5)
    taps_of_texel_adx<ADX, 4>(acc, ctr, q4, g4, w4, sigma_n);
    RMD_TILE_LOAD(6, q6, g6, w6)
    taps_of_texel_adx<ADX, 5>(acc, ctr, q5, g5, w5, sigma_n);
    RMD_TILE_LOAD(7, q7, g7, w7)
    taps_of_texel_adx<ADX, 6>(acc, ctr, q6, g6, w6, sigma_n);
    taps_of_texel_adx<ADX, 7>(acc, ctr, q7, g7, w7, sigma_n);
#undef RMD_TILE_LOAD
}

// The centre column (dx = 0) with the taps between two of the thread's own outputs evaluated in pairs (MODE bit 5).
// Tile rows 2..5 are the centres of outputs 0..3: their normal, depth and luminance are in ctr[], so only colour and
// variance are loaded, and each of the 5 pairs (0,1) (0,2) (1,2) (1,3) (2,3) costs one dot product, one lg2 and one
// |dz| / |dL| instead of two.  The weight of the later tap of a pair waits in a register until its texel's row is
// reached: every accumulator receives its taps in tile-row order, i.e. the result has the bits of tile_column<T, 0>.
template <class T>
__device__ __forceinline__ void tile_centre_column_pairs(Acc (&acc)[kAtrousOPT], const Centre (&ctr)[kAtrousOPT], const uint32_t b,
                                                         const uint32_t vb, const float sigma_n) {
    static_assert(T::TH == 8 && kAtrousOPT == 4, "written out for 4 outputs per thread");
#define RMD_TILE_LOAD(JR, Q, G, VV)                                   \
    const float4 Q = lds128<T::OFF_C4 + (JR)*T::HW2 * 16>(b);         \
    const float4 G = lds128<T::OFF_G4 + (JR)*T::HW2 * 16>(b);         \
    const float VV = lds32<T::OFF_V + (JR)*T::TW * 4>(vb);
#define RMD_TILE_LOAD_QV(JR, Q, VV)                                   \
    const float4 Q = lds128<T::OFF_C4 + (JR)*T::HW2 * 16>(b);         \
    const float VV = lds32<T::OFF_V + (JR)*T::TW * 4>(vb);
    RMD_TILE_LOAD(0, q0, g0, w0)
    RMD_TILE_LOAD(1, q1, g1, w1)
    taps_of_texel_adx<0, 0>(acc, ctr, q0, g0, w0, sigma_n);
    RMD_TILE_LOAD_QV(2, q2, w2)
    taps_of_texel_adx<0, 1>(acc, ctr, q1, g1, w1, sigma_n);
    RMD_TILE_LOAD_QV(3, q3, w3)
    // row 2 = centre of output 0: read by outputs 1 (dy = -1) and 2 (dy = -2)
    float h01, h10, h02, h20;
    pair_weights<1>(ctr[0], ctr[1], sigma_n, h01, h10);
    tap_accumulate(acc[1], h10, q2, w2);
    pair_weights<2>(ctr[0], ctr[2], sigma_n, h02, h20);
    tap_accumulate(acc[2], h20, q2, w2);
    RMD_TILE_LOAD_QV(4, q4, w4)
    // row 3 = centre of output 1: read by outputs 0 (dy = +1), 2 (dy = -1), 3 (dy = -2)
    float h12, h21, h13, h31;
    tap_accumulate(acc[0], h01, q3, w3);
    pair_weights<1>(ctr[1], ctr[2], sigma_n, h12, h21);
    tap_accumulate(acc[2], h21, q3, w3);
    pair_weights<2>(ctr[1], ctr[3], sigma_n, h13, h31);
    tap_accumulate(acc[3], h31, q3, w3);
    RMD_TILE_LOAD_QV(5, q5, w5)
    // row 4 = centre of output 2: read by outputs 0 (dy = +2), 1 (dy = +1), 3 (dy = -1)
    float h23, h32;
    tap_accumulate(acc[0], h02, q4, w4);
    tap_accumulate(acc[1], h12, q4, w4);
    pair_weights<1>(ctr[2], ctr[3], sigma_n, h23, h32);
    tap_accumulate(acc[3], h32, q4, w4);
    RMD_TILE_LOAD(6, q6, g6, w6)
    // row 5 = centre of output 3: read by outputs 1 (dy = +2), 2 (dy = +1)
    tap_accumulate(acc[1], h13, q5, w5);
    tap_accumulate(acc[2], h23, q5, w5);
    RMD_TILE_LOAD(7, q7, g7, w7)
    taps_of_texel_adx<0, 6>(acc, ctr, q6, g6, w6, sigma_n);
    taps_of_texel_adx<0, 7>(acc, ctr, q7, g7, w7, sigma_n);
#undef RMD_TILE_LOAD
#undef RMD_TILE_LOAD_QV
}

struct TilePos {
    int x0, phase, k0;
    int bx, hi, lo;  // digits of the tile index: t = (hi * by_div + lo) * nbx + bx
};

__host__ __device__ __forceinline__ bool row_in_launch(const AtrousArgs& a, int y) { return y >= a.row0 && y < a.row0 + a.rows; }

// tile index -> (first column, row phase, first lattice row); false when this launch produces none of its rows
// KT_MAJOR: (lattice tile, phase) instead of (phase, lattice tile) — the S row phases of one lattice tile row
// (4*S consecutive image rows) run as neighbouring CTAs, so the rows y-1 / y+1 a tile's centre terms read are the tile
// rows of CTAs resident at the same time (L2 hits instead of a second and third DRAM read of the variance plane)
// (__host__ too: rmd_debug_level_cover enumerates a launch's tiles on the CPU with this very function)
template <int S, bool KT_MAJOR>
__host__ __device__ __forceinline__ bool tile_pos(const AtrousArgs& a, int t, int nbx, int by_div, TilePos& p) {
    // by enumerates (phase, lattice tile of the launch's row ranges) with by_div = lattice tiles per phase, or
    // (lattice tile, phase) with by_div = number of phases when KT_MAJOR
    const int by = t / nbx;
    p.bx = t - by * nbx;
    p.x0 = p.bx * tile_wt(S);
    const int hi = by / by_div, lo = by - hi * by_div;
    p.hi = hi; p.lo = lo;
    p.phase = KT_MAJOR ? lo : hi;
    const int kt = KT_MAJOR ? hi : lo;
    p.k0 = (kt < a.kt_cnt[0] ? a.kt_lo[0] + kt : a.kt_lo[1] + kt - a.kt_cnt[0]) * kAtrousTY;
    const int y_first = p.phase + S * p.k0;
    if (y_first >= a.H) return false;  // this phase has fewer lattice rows
    const int y_last = y_first + S * (kAtrousTY - 1);
    if (y_last < a.row0 || y_first >= a.row0 + a.rows) return false;  // band mode: none of the launch's rows
    if (a.split == 0) return true;
    // a tile holds a row of range [e, e+n) iff one of its rows y_first + S*j (j < TY) lies in it
    bool edge = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (a.edgeN[i] <= 0) continue;
#pragma unroll
        for (int j = 0; j < kAtrousTY; ++j) {
            const int y = y_first + S * j;
            edge |= y >= a.edge0[i] && y < a.edge0[i] + a.edgeN[i];
        }
    }
    return a.split == 1 ? edge : !edge;
}

// thread 0: one mbarrier phase = every box of the tile
template <class T, int S>
__device__ __forceinline__ void issue_tile(uint8_t* smem, uint64_t* bar, const AtrousMaps& maps, const TilePos& p) {
    mbar_arrive_expect_tx(bar, T::TX_BYTES);
    const int cx = 2 * (p.x0 - T::HX);  // 8-byte elements: 2 per texel
    tma_load_3d(smem + T::OFF_C4, &maps.c4, bar, cx, p.phase, p.k0 - 2);
    tma_load_3d(smem + T::OFF_C4 + T::HALF_BYTES, &maps.c4, bar, cx + 2 * T::HW2, p.phase, p.k0 - 2);
    tma_load_3d(smem + T::OFF_G4, &maps.g4, bar, cx, p.phase, p.k0 - 2);
    tma_load_3d(smem + T::OFF_G4 + T::HALF_BYTES, &maps.g4, bar, cx + 2 * T::HW2, p.phase, p.k0 - 2);
    tma_load_3d(smem + T::OFF_V, &maps.v, bar, p.x0 - T::HX, p.phase, p.k0 - 2);
    if constexpr (T::PRO) {
        if constexpr (S > 1) {
            // neighbour-phase rows: image row y-1 of lattice row k is (phase-1, k), or (S-1, k-1) when phase == 0;
            // image row y+1 is (phase+1, k), or (0, k+1) when phase == S-1
            const int pm = p.phase > 0 ? p.phase - 1 : S - 1, km = p.phase > 0 ? p.k0 : p.k0 - 1;
            const int pp = p.phase < S - 1 ? p.phase + 1 : 0, kp = p.phase < S - 1 ? p.k0 : p.k0 + 1;
            tma_load_3d(smem + T::OFF_VM, &maps.vn, bar, p.x0 - 4, pm, km);
            tma_load_3d(smem + T::OFF_VP, &maps.vn, bar, p.x0 - 4, pp, kp);
        }
        tma_load_3d(smem + T::OFF_DZ, &maps.dzm, bar, p.x0, p.phase, p.k0);
    }
}

// thread 0: pull the boxes of a tile that will be staged about one wave of CTAs from now into L2, so that its TMA
// loads (and the centre-term loads of its rows) are L2 hits when it starts
template <class T, int S>
__device__ __forceinline__ void prefetch_tile(const AtrousMaps& maps, const TilePos& p) {
    const int cx = 2 * (p.x0 - T::HX);
    tma_prefetch_l2_3d(&maps.c4, cx, p.phase, p.k0 - 2);
    tma_prefetch_l2_3d(&maps.c4, cx + 2 * T::HW2, p.phase, p.k0 - 2);
    tma_prefetch_l2_3d(&maps.g4, cx, p.phase, p.k0 - 2);
    tma_prefetch_l2_3d(&maps.g4, cx + 2 * T::HW2, p.phase, p.k0 - 2);
    tma_prefetch_l2_3d(&maps.v, p.x0 - T::HX, p.phase, p.k0 - 2);
}

// position of the tile `a.prefetch_ahead` tiles further along the walk, from the digits of this tile's index and the
// look-ahead's (host-computed) digits: carries instead of two more integer divisions.  No launch-membership test:
// prefetching a tile another launch produces is harmless.
template <int S, bool KT_MAJOR, bool BACKWARD>
__device__ __forceinline__ void tile_ahead(const AtrousArgs& a, const TilePos& p, int nbx, int by_div, TilePos& q) {
    int bx, lo, hi;
    if (BACKWARD) {
        bx = p.bx - a.pf_dx;
        const int c0 = bx < 0;
        bx += c0 ? nbx : 0;
        lo = p.lo - a.pf_dlo - c0;
        const int c1 = lo < 0;
        lo += c1 ? by_div : 0;
        hi = p.hi - a.pf_dhi - c1;
    } else {
        bx = p.bx + a.pf_dx;
        const int c0 = bx >= nbx;
        bx -= c0 ? nbx : 0;
        lo = p.lo + a.pf_dlo + c0;
        const int c1 = lo >= by_div;
        lo -= c1 ? by_div : 0;
        hi = p.hi + a.pf_dhi + c1;
    }
    q.x0 = bx * tile_wt(S);
    q.phase = KT_MAJOR ? lo : hi;
    const int kt = KT_MAJOR ? hi : lo;
    q.k0 = (kt < a.kt_cnt[0] ? a.kt_lo[0] + kt : a.kt_lo[1] + kt - a.kt_cnt[0]) * kAtrousTY;
}

// RMD_NO_TMA=1: the same boxes with plain coalesced loads and explicit zero fill (cross-check of the TMA path)
template <class T, int S>
__device__ __forceinline__ void stage_tile_plain(uint8_t* smem, const AtrousArgs& a, const TilePos& p, int tx) {
    const int W = a.W, H = a.H, Wp = a.Wp;
    float4* wC4 = reinterpret_cast<float4*>(smem + T::OFF_C4);
    float4* wG4 = reinterpret_cast<float4*>(smem + T::OFF_G4);
    float* wV = reinterpret_cast<float*>(smem + T::OFF_V);
    for (int i = tx; i < T::TW * T::TH; i += T::WT) {
        const int row = i / T::TW, col = i - row * T::TW;
        const int gx = p.x0 - T::HX + col, k = p.k0 - 2 + row;
        const int gy = p.phase + S * k;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f), g = c;
        float v = 0.f;
        if (gx >= 0 && gx < W && k >= 0 && gy < H) {
            const size_t q = (size_t)gy * Wp + gx;
            c = a.in_c4[q];
            g = a.g4[q];
            v = a.in_v[q];
        }
        wC4[T::coloff(col) + row * T::HW2] = c;
        wG4[T::coloff(col) + row * T::HW2] = g;
        wV[i] = v;
    }
    if constexpr (T::PRO) {
        float* wDZ = reinterpret_cast<float*>(smem + T::OFF_DZ);
        for (int i = tx; i < T::WT * kAtrousTY; i += T::WT) {
            const int row = i / T::WT, col = i - row * T::WT;
            const int gx = p.x0 + col, gy = p.phase + S * (p.k0 + row);
            wDZ[i] = (gx < W && gy < H) ? a.dz[(size_t)gy * Wp + gx] : 0.f;
        }
        if constexpr (S > 1) {
            const int pm = p.phase > 0 ? p.phase - 1 : S - 1, km = p.phase > 0 ? p.k0 : p.k0 - 1;
            const int pp = p.phase < S - 1 ? p.phase + 1 : 0, kp = p.phase < S - 1 ? p.k0 : p.k0 + 1;
            float* wVM = reinterpret_cast<float*>(smem + T::OFF_VM);
            float* wVP = reinterpret_cast<float*>(smem + T::OFF_VP);
            for (int i = tx; i < T::NBW * kAtrousTY; i += T::WT) {
                const int row = i / T::NBW, col = i - row * T::NBW;
                const int gx = p.x0 - 4 + col;
                const int ym = pm + S * (km + row), yp = pp + S * (kp + row);
                const bool xin = gx >= 0 && gx < W;
                wVM[i] = (xin && ym >= 0 && ym < H) ? a.in_v[(size_t)ym * Wp + gx] : 0.f;
                wVP[i] = (xin && yp >= 0 && yp < H) ? a.in_v[(size_t)yp * Wp + gx] : 0.f;
            }
        }
    }
}

template <int S, int MODE, int MINB>
__global__ void __launch_bounds__(tile_wt(S), MINB)
    atrous_kernel(const AtrousArgs a, const __grid_constant__ AtrousMaps maps, const int nbx, const int by_div,
                  const int total_tiles) {
    using T = Tile<S, MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);

    // PDL: nothing below may run ahead of the previous kernel in the stream.  The wait sits before the early
    // exits so that this grid can never complete before its predecessor has (a grid whose CTAs all returned
    // early would otherwise release ITS successor too soon).
    pdl_wait();
    const int W = a.W, H = a.H, Wp = a.Wp;
    const int tx = threadIdx.x;
    // tiles of this CTA: t = blockIdx.x (+ k * gridDim.x when persistent)
    int t = blockIdx.x;
    if constexpr (T::SERPENTINE) {
        if (a.reverse) t = total_tiles - 1 - t;
    }
    TilePos p;
    if constexpr (T::PERSIST) {
        while (t < total_tiles && !tile_pos<S, T::KT_MAJOR>(a, t, nbx, by_div, p)) t += gridDim.x;
        if (t >= total_tiles) return;
    } else {
        if (!tile_pos<S, T::KT_MAJOR>(a, t, nbx, by_div, p)) return;  // uniform per CTA
    }
    if (a.use_tma) {
        if (tx == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();  // the barrier must be initialised before any thread polls it
        if (tx == 0) {
            issue_tile<T, S>(smem, bar, maps, p);
            if constexpr (T::PREFETCH) {
                const bool back = T::SERPENTINE && a.reverse;
                const int ta = back ? t - a.prefetch_ahead : t + a.prefetch_ahead;
                if (a.prefetch_ahead > 0 && ta >= 0 && ta < total_tiles) {
                    TilePos pa;
                    if (back) tile_ahead<S, T::KT_MAJOR, true>(a, p, nbx, by_div, pa);
                    else tile_ahead<S, T::KT_MAJOR, false>(a, p, nbx, by_div, pa);
                    prefetch_tile<T, S>(maps, pa);
                }
            }
        }
    }
    pdl_launch_dependents();

    const uint32_t sbase = smem_u32(smem);
    const uint32_t ccol = sbase + 16u * (uint32_t)T::coloff(tx + T::HX);   // centre column, float4 planes
    const uint32_t vcol = sbase + T::OFF_V + 4u * (uint32_t)(tx + T::HX);  // centre column, variance
    // per-thread column bases of the 5 tap columns (shared-window byte addresses); every tap load is
    // [register + compile-time immediate]
    uint32_t cb[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) cb[c] = sbase + 16u * (uint32_t)T::coloff(tx + (T::HX - 2 * S) + c * S);
    const uint32_t vb = sbase + 4u * (uint32_t)(tx + (T::HX - 2 * S));
    const float sigma_n = a.sigma_n;
    uint32_t parity = 0;

    while (true) {
        const int x = p.x0 + tx;
        if (!a.use_tma) stage_tile_plain<T, S>(smem, a, p, tx);
        float vbar[kAtrousOPT], dzv[kAtrousOPT];
        if constexpr (!T::PRO) {
            // ---- per-output centre terms from global memory (L1/L2), issued before the tile wait ----
            // clamp-to-edge in x by selection, not by address: all 36 loads are [row pointer + immediate] and
            // independent; at x = 0 / x = W-1 the neighbour load reads the adjacent padding element (the planes
            // carry a guard at either end, svgf_ctx.cu) and its value is replaced by the centre column's
            const int xc = min(x, W - 1);
            const bool has_l = xc > 0, has_r = xc < W - 1;
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                const int y = min(p.phase + S * (p.k0 + j), H - 1);
                const int ym = max(y - 1, 0), yp = min(y + 1, H - 1);
                const float* r0 = a.in_v + ((size_t)ym * Wp + xc);
                const float* r1 = a.in_v + ((size_t)y * Wp + xc);
                const float* r2 = a.in_v + ((size_t)yp * Wp + xc);
                const float t0 = __ldg(r0 - 1), t1 = __ldg(r0), t2 = __ldg(r0 + 1);
                const float m0 = __ldg(r1 - 1), m1 = __ldg(r1), m2 = __ldg(r1 + 1);
                const float b0 = __ldg(r2 - 1), b1 = __ldg(r2), b2 = __ldg(r2 + 1);
                vbar[j] = vbar3x3(has_l ? t0 : t1, t1, has_r ? t2 : t1, has_l ? m0 : m1, m1, has_r ? m2 : m1,
                                  has_l ? b0 : b1, b1, has_r ? b2 : b1);
                dzv[j] = __ldg(a.dz + ((size_t)y * Wp + xc));
            }
        }

        if (a.use_tma) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            __syncthreads();
        }

        if constexpr (T::PRO) {
            // ---- per-output centre terms from the staged rows: 3x3 Gaussian of the variance with clamped
            //      coordinates (spec S4) and the depth slope ----
            const bool has_l = x > 0, has_r = x < W - 1;
            const uint32_t nb = sbase + 4u * (uint32_t)(tx + 4);
            const uint32_t dzb = sbase + T::OFF_DZ + 4u * (uint32_t)tx;
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                const int y = p.phase + S * (p.k0 + j);
                float t0, t1, t2, b0, b1, b2;
                const float m0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW - 1));
                const float m1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW));
                const float m2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW + 1));
                if constexpr (S > 1) {
                    t0 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW - 1));
                    t1 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW));
                    t2 = lds32_dyn(nb + T::OFF_VM + 4u * (uint32_t)(j * T::NBW + 1));
                    b0 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW - 1));
                    b1 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW));
                    b2 = lds32_dyn(nb + T::OFF_VP + 4u * (uint32_t)(j * T::NBW + 1));
                } else {  // step 1: the neighbouring image rows are the neighbouring tile rows
                    t0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW - 1));
                    t1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW));
                    t2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 1) * T::TW + 1));
                    b0 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW - 1));
                    b1 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW));
                    b2 = lds32_dyn(vcol + 4u * (uint32_t)((j + 3) * T::TW + 1));
                }
                if (y <= 0) { t0 = m0; t1 = m1; t2 = m2; }        // row y-1 clamps to row y at the top edge
                if (y >= H - 1) { b0 = m0; b1 = m1; b2 = m2; }    // row y+1 clamps to row y at the bottom edge
                vbar[j] = vbar3x3(has_l ? t0 : t1, t1, has_r ? t2 : t1, has_l ? m0 : m1, m1, has_r ? m2 : m1,
                                  has_l ? b0 : b1, b1, has_r ? b2 : b1);
                dzv[j] = lds32_dyn(dzb + 4u * (uint32_t)(j * T::WT));
            }
        }

        // ---- centre set-up ---------------------------------------------------------------
        Centre ctr[kAtrousOPT];
        Acc acc[kAtrousOPT];
#pragma unroll
        for (int j = 0; j < kAtrousOPT; ++j) {
            const float4 c = lds128_dyn(ccol + T::OFF_C4 + 16u * (uint32_t)((j + 2) * T::HW2));
            const float4 g = lds128_dyn(ccol + T::OFF_G4 + 16u * (uint32_t)((j + 2) * T::HW2));
            const float v = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW));
            centre_setup<S>(ctr[j], acc[j], c, g, v, vbar[j], dzv[j], a);
        }

        // ---- 100 taps from 40 staged texels ---------------------------------------------
        if constexpr (T::GROUPED) {
#pragma unroll 1
            for (int it = 0; it < 2; ++it)  // |dx| = 2: columns 0 and 4
                tile_column<T, 2>(acc, ctr, it ? cb[4] : cb[0], vb + (it ? 16u * S : 0u), sigma_n);
#pragma unroll 1
            for (int it = 0; it < 2; ++it)  // |dx| = 1: columns 1 and 3
                tile_column<T, 1>(acc, ctr, it ? cb[3] : cb[1], vb + (it ? 12u * S : 4u * S), sigma_n);
            if constexpr (T::PAIRS) tile_centre_column_pairs<T>(acc, ctr, cb[2], vb + 8u * S, sigma_n);
            else tile_column<T, 0>(acc, ctr, cb[2], vb + 8u * S, sigma_n);
        } else {
            tile_column<T, 2>(acc, ctr, cb[0], vb, sigma_n);
            tile_column<T, 1>(acc, ctr, cb[1], vb + 4u * S, sigma_n);
            tile_column<T, 0>(acc, ctr, cb[2], vb + 8u * S, sigma_n);
            tile_column<T, 1>(acc, ctr, cb[3], vb + 12u * S, sigma_n);
            tile_column<T, 2>(acc, ctr, cb[4], vb + 16u * S, sigma_n);
        }

        if constexpr (!T::PERSIST) {
            // ---- epilogue -----------------------------------------------------------------
            if (x >= W) return;
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                const int y = p.phase + S * (p.k0 + j);
                if (row_in_launch(a, y))
                    store_output(a, acc[j], ctr[j], ccol + T::OFF_C4 + 16u * (uint32_t)((j + 2) * T::HW2),
                                 vcol + 4u * (uint32_t)((j + 2) * T::TW), x, y);
            }
            return;
        } else {
            // ---- persistent: the next tile's load is issued before this tile's epilogue ------------
            // sky outputs pass their input through: fetch it while the tile is still in shared memory
#pragma unroll
            for (int j = 0; j < kAtrousOPT; ++j) {
                if (ctr[j].z == 0.0f) {
                    const float4 cC = lds128_dyn(ccol + T::OFF_C4 + 16u * (uint32_t)((j + 2) * T::HW2));
                    acc[j].r = cC.x; acc[j].g = cC.y; acc[j].b = cC.z; acc[j].w = 1.0f;
                    acc[j].v = lds32_dyn(vcol + 4u * (uint32_t)((j + 2) * T::TW));
                    ctr[j].L = cC.w;
                }
            }
            int tn = t + gridDim.x;
            TilePos pn;
            while (tn < total_tiles && !tile_pos<S, T::KT_MAJOR>(a, tn, nbx, by_div, pn)) tn += gridDim.x;
            const bool more = tn < total_tiles;
            __syncthreads();  // every thread has finished reading the tile
            if (more && a.use_tma && tx == 0) issue_tile<T, S>(smem, bar, maps, pn);
            if (x < W) {
#pragma unroll
                for (int j = 0; j < kAtrousOPT; ++j) {
                    const int y = p.phase + S * (p.k0 + j);
                    if (row_in_launch(a, y)) store_output_regs(a, acc[j], ctr[j], x, y);
                }
            }
            if (!more) return;
            t = tn;
            p = pn;
        }
    }
}

// Tile geometry of one launch: fills a.kt_lo / a.kt_cnt and returns false when the launch has no tile.
struct LevelGrid {
    int nbx, phases, tiles_per_phase, total, by_div;
};
template <int S>
bool level_grid(AtrousArgs& a, LevelGrid& g) {
    using T = Tile<S, kMode>;
    // lattice tiles that can hold rows of the launch's (one or two) row ranges: row y of phase (y mod S) is lattice
    // row y / S, so rows [r0, r1) live in lattice tiles [(r0/S)/TY, ((r1-1)/S)/TY] of every phase.  A band's boundary
    // launch therefore enumerates a few tile rows instead of the whole plane.
    const int lat_tiles = ((a.H + S - 1) / S + kAtrousTY - 1) / kAtrousTY;
    int lo[2] = {0, 0}, hi[2] = {0, 0};
    // split == 1 enumerates only the tile rows around the two edge ranges, otherwise those of the stored rows
    const int r0[2] = {a.split == 1 ? a.edge0[0] : a.row0, a.split == 1 ? a.edge0[1] : 0};
    const int rn[2] = {a.split == 1 ? a.edgeN[0] : a.rows, a.split == 1 ? a.edgeN[1] : 0};
    for (int i = 0; i < 2; ++i) {
        if (rn[i] <= 0) continue;
        lo[i] = (r0[i] / S) / kAtrousTY;
        hi[i] = ((r0[i] + rn[i] - 1) / S) / kAtrousTY + 1;
        if (hi[i] > lat_tiles) hi[i] = lat_tiles;
    }
    if (rn[1] > 0 && rn[0] > 0 && lo[1] < hi[0] && lo[0] < hi[1]) {  // overlapping: one merged range
        lo[0] = lo[0] < lo[1] ? lo[0] : lo[1];
        hi[0] = hi[0] > hi[1] ? hi[0] : hi[1];
        lo[1] = hi[1] = 0;
    }
    a.kt_lo[0] = lo[0]; a.kt_cnt[0] = hi[0] - lo[0];
    a.kt_lo[1] = lo[1]; a.kt_cnt[1] = hi[1] - lo[1];
    g.tiles_per_phase = a.kt_cnt[0] + a.kt_cnt[1];
    if (g.tiles_per_phase <= 0) return false;
    g.phases = S < a.H ? S : a.H;
    g.nbx = (a.W + T::WT - 1) / T::WT;
    g.total = g.nbx * g.phases * g.tiles_per_phase;
    g.by_div = T::KT_MAJOR ? g.phases : g.tiles_per_phase;
    return true;
}

template <int S>
int launch_level(const AtrousArgs& a_in, const AtrousMaps& maps, cudaStream_t s, bool pdl) {
    using T = Tile<S, kMode>;
    AtrousArgs a = a_in;
    LevelGrid lg;
    if (!level_grid<S>(a, lg)) return 0;
    const int nbx = lg.nbx, total = lg.total;
    int grid = total;
    if (T::PERSIST) {  // one CTA per resident slot; each walks the tile list with a grid stride
        const int smem_ctas = (int)((size_t)g_smem_per_sm / (size_t)(T::SMEM + 1024));
        const int per_sm = tile_minb(S) < smem_ctas ? tile_minb(S) : smem_ctas;
        const int slots = g_sms * (per_sm > 0 ? per_sm : 1);
        if (grid > slots) grid = slots;
    }
    if (T::PREFETCH && a.prefetch_ahead > 0) {  // the look-ahead as digits of the tile index (tile_ahead)
        const int dby = a.prefetch_ahead / nbx;
        a.pf_dx = a.prefetch_ahead - dby * nbx;
        a.pf_dhi = dby / lg.by_div;
        a.pf_dlo = dby - a.pf_dhi * lg.by_div;
    } else {
        a.prefetch_ahead = 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(T::WT, 1);
    cfg.dynamicSmemBytes = T::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, atrous_kernel<S, kMode, tile_minb(S)>, a, maps, nbx, lg.by_div, total);
}

// Host enumeration of the same launch (debug / CPU tests): cover[y * nbx + bx] += 1 for every (output row, column
// block) a CTA of the launch would store, walking the tiles in the order the grid does.  Returns the tile count.
template <int S>
int cover_level(const AtrousArgs& a_in, int* cover, int* nbx_out, int* tiles_with_work) {
    using T = Tile<S, kMode>;
    AtrousArgs a = a_in;
    LevelGrid lg;
    if (nbx_out) *nbx_out = (a.W + T::WT - 1) / T::WT;
    if (tiles_with_work) *tiles_with_work = 0;
    if (!level_grid<S>(a, lg)) return 0;
    for (int b = 0; b < lg.total; ++b) {
        const int t = (T::SERPENTINE && a.reverse) ? lg.total - 1 - b : b;
        TilePos p;
        if (!tile_pos<S, T::KT_MAJOR>(a, t, lg.nbx, lg.by_div, p)) continue;
        if (tiles_with_work) ++*tiles_with_work;
        for (int j = 0; j < kAtrousOPT; ++j) {
            const int y = p.phase + S * (p.k0 + j);
            if (y < a.H && row_in_launch(a, y)) cover[(size_t)y * lg.nbx + p.bx] += 1;
        }
    }
    return lg.total;
}

template <int S>
int configure_level() {
    return (int)cudaFuncSetAttribute(atrous_kernel<S, kMode, tile_minb(S)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Tile<S, kMode>::SMEM);
}

}  // namespace

#define RMD_CAT2(a, b) a##b
#define RMD_CAT(a, b) RMD_CAT2(a, b)

int RMD_CAT(atrous_tile_configure_v, RMD_VARIANT)() {
    int dev = 0;
    RMD_CUDA_TRY(cudaGetDevice(&dev));
    RMD_CUDA_TRY(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    RMD_CUDA_TRY(cudaDeviceGetAttribute(&g_smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    int rc = configure_level<1>(); if (rc) return rc;
    rc = configure_level<2>(); if (rc) return rc;
    rc = configure_level<4>(); if (rc) return rc;
    rc = configure_level<8>(); if (rc) return rc;
    return configure_level<16>();
}

int RMD_CAT(atrous_tile_width_v, RMD_VARIANT)(int level) { return tile_wt(1 << level); }

int RMD_CAT(atrous_tile_cover_v, RMD_VARIANT)(int level, const AtrousArgs& a, int* cover, int* nbx_out, int* tiles_with_work) {
    switch (level) {
        case 0: return cover_level<1>(a, cover, nbx_out, tiles_with_work);
        case 1: return cover_level<2>(a, cover, nbx_out, tiles_with_work);
        case 2: return cover_level<4>(a, cover, nbx_out, tiles_with_work);
        case 3: return cover_level<8>(a, cover, nbx_out, tiles_with_work);
        case 4: return cover_level<16>(a, cover, nbx_out, tiles_with_work);
        default: return RMD_E_PARAM;
    }
}

int RMD_CAT(launch_atrous_tile_v, RMD_VARIANT)(int level, const AtrousArgs& a, const AtrousMaps& maps, cudaStream_t s,
                                               bool pdl) {
    switch (level) {
        case 0: return launch_level<1>(a, maps, s, pdl);
        case 1: return launch_level<2>(a, maps, s, pdl);
        case 2: return launch_level<4>(a, maps, s, pdl);
        case 3: return launch_level<8>(a, maps, s, pdl);
        case 4: return launch_level<16>(a, maps, s, pdl);
        default: return RMD_E_PARAM;
    }
}

}  // namespace rmd
