/*
 * oracle.h — C interface of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  See oracle_box.c / oracle_svgf.c for the parity status.
 */
#ifndef RMD_ORACLE_H
#define RMD_ORACLE_H
#include <stdint.h>
#include "../include/rmd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* legacy box path (reference src/filter.cu:13-58, 87-158) */
void oracle_box_baseline_level(const uint8_t* in, uint8_t* out, int W, int H, int radius);
void oracle_box_tiled_level(const uint8_t* in, uint8_t* out, int W, int H, int radius);
int oracle_box_filter(const uint8_t* render, uint8_t* denoised, uint8_t* buf0, uint8_t* buf1, int W, int H,
                      int radius, int depth, int variant);

/* FilterParams::GAUSSIAN / CROSS on the RGBA8 planes (oracle_weighted.c; no reference implementation exists) */
int oracle_weighted_scales(const RmdFilterParams* p, float out[4]);
int oracle_weighted_filter(const uint8_t* render, uint8_t* denoised, uint8_t* buf0, uint8_t* buf1, const uint8_t* albedo,
                           const uint8_t* normal, int W, int H, const RmdFilterParams* p);

/* SVGF (published algorithm) */
typedef struct oracle_svgf oracle_svgf;
oracle_svgf* oracle_svgf_create(int W, int H);
void oracle_svgf_destroy(oracle_svgf* s);
void oracle_svgf_reset(oracle_svgf* s);
/* all planes of `frame` are HOST pointers in the storage formats of RmdSvgfFrame */
int oracle_svgf_frame(oracle_svgf* s, const RmdSvgfFrame* frame, const RmdFilterParams* fp, const RmdSvgfParams* sp);
/* plane ids: RMD_PLANE_* plus the two below (temporal output before the variance pass) */
#define ORACLE_PLANE_TEMPORAL_COLOR_PRE 100
#define ORACLE_PLANE_TEMPORAL_VAR_PRE 101
const void* oracle_svgf_plane(const oracle_svgf* s, int plane);
/* OpenMP team size used by the oracle (set explicitly by bench.py; reported in cpu_baseline.cores) */
void oracle_set_threads(int n);
int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
