// common.cuh — shared device/host helpers of librmd_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rmd_b200_debug.h"

#define RMD_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

namespace rmd {

// ---- luminance weights (DESIGN.md spec S0) ------------------------------------------
__device__ __forceinline__ float luminance(float r, float g, float b) {
    return fmaf(0.0722f, b, fmaf(0.7152f, g, 0.2126f * r));
}

// ---- guide decode: fp32, operation order fixed by the spec ----------
// Every operation is an explicitly rounded, un-fused IEEE op so that the decoded
// normal/z (and hence every reprojection predicate that consumes them) is bit
// identical to the oracle's (-ffp-contract=off) evaluation.
__device__ __forceinline__ float4 decode_guide(uint2 g) {
    const float z = __uint_as_float(g.y);
    if (!(z > 0.0f) || !isfinite(z)) return make_float4(0.f, 0.f, 0.f, 0.f);  // sky
    int sx = (int)(short)(g.x & 0xFFFFu), sy = (int)(short)(g.x >> 16);
    sx = max(sx, -32767);
    sy = max(sy, -32767);
    const float c = 1.0f / 32767.0f;
    float fx = __fmul_rn((float)sx, c), fy = __fmul_rn((float)sy, c);
    const float fz = __fsub_rn(__fsub_rn(1.0f, fabsf(fx)), fabsf(fy));
    if (fz < 0.0f) {
        const float ox = __fmul_rn(__fsub_rn(1.0f, fabsf(fy)), fx >= 0.0f ? 1.0f : -1.0f);
        const float oy = __fmul_rn(__fsub_rn(1.0f, fabsf(fx)), fy >= 0.0f ? 1.0f : -1.0f);
        fx = ox;
        fy = oy;
    }
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(len2));
    return make_float4(__fmul_rn(fx, inv), __fmul_rn(fy, inv), __fmul_rn(fz, inv), z);
}

// un-fused dot product, same order as the oracle's dot3f
__device__ __forceinline__ float dot3_rn(float4 a, float4 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

// ---- cache-hinted global accesses -----------------------------------------------------
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) {  // write-once streaming output
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- mbarrier + TMA (cp.async.bulk.tensor) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(m), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3)
                 : "memory");
}

// ---- Ampere-style async copies (per-thread, used for small clamped gathers) ----------
__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- fast transcendental wrappers (MUFU) ---------------------------------------------
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still draining; pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible (no-op for an ordinary launch), pdl_launch_dependents()
// lets the successor's CTAs take SM slots as this grid's CTAs retire.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace rmd
