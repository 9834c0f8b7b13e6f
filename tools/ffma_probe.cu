// ffma_probe.cu — FP32 issue-rate probe: scalar FFMA vs packed FFMA2 (debug tool, not product).
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
constexpr int ITER = 2048, CH = 8;

__global__ void k_scalar(float* out, float a, float b) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 0.001f + i;
    float y = a + threadIdx.x, z = b;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], y, z);   // 3 distinct registers per FFMA
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_2src(float* out, float a, float b) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 0.001f + i;
    float y = a + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], y, x[i]);   // 2 distinct registers
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
    float2 x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    float2 y = make_float2(a + threadIdx.x, a - threadIdx.x), z = make_float2(b, b + 1);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = __ffma2_rn(x[i], y, z);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_distinct(float* out, float a, float b) {
    float x[CH], p[CH], q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 0.001f + i; p[i] = a + i * 1e-4f + threadIdx.x * 1e-6f; q[i] = b + i * 1e-3f; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(p[i], q[(i + 3) % CH], x[i]);   // 3 distinct, none shared with the neighbouring FFMAs
#pragma unroll
        for (int i = 0; i < CH; ++i) p[i] = fmaf(q[i], x[(i + 5) % CH], p[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_distinct(float* out, float a, float b) {
    float2 x[CH], p[CH], q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = make_float2(threadIdx.x * 0.001f + i, i); p[i] = make_float2(a + i * 1e-4f + threadIdx.x * 1e-6f, a); q[i] = make_float2(b + i * 1e-3f, b); }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = __ffma2_rn(p[i], q[(i + 3) % CH], x[i]);
#pragma unroll
        for (int i = 0; i < CH; ++i) p[i] = __ffma2_rn(q[i], x[(i + 5) % CH], p[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float* out, float a, float b) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 0.001f + i + a;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> float time_it(K k, float* out, int blocks) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, 256>>>(out, 1.0001f, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<blocks, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int blocks = p.multiProcessorCount * 8;
    float* out; CK(cudaMalloc(&out, (size_t)blocks * 256 * 4));
    double clk = p.clockRate * 1e3;
    double winst = (double)blocks * 8 * ITER * CH;  // warp-instructions
    float t;
    t = time_it(k_scalar, out, blocks);      printf("scalar FFMA 3-reg : %.3f ms  %.2f warp-inst/clk/SMSP\n", t, winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    t = time_it(k_scalar_2src, out, blocks); printf("scalar FFMA 2-reg : %.3f ms  %.2f warp-inst/clk/SMSP\n", t, winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    t = time_it(k_packed, out, blocks);      printf("packed FFMA2      : %.3f ms  %.2f warp-inst/clk/SMSP (x2 FMAs each)\n", t, winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    t = time_it(k_scalar_distinct, out, blocks); printf("scalar FFMA distinct: %.3f ms  %.2f warp-inst/clk/SMSP\n", t, 2 * winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    t = time_it(k_packed_distinct, out, blocks); printf("packed FFMA2 distinct: %.3f ms  %.2f warp-inst/clk/SMSP (x2 FMAs each)\n", t, 2 * winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    t = time_it(k_mufu, out, blocks);        printf("MUFU.EX2          : %.3f ms  %.3f warp-inst/clk/SMSP\n", t, winst / (t * 1e-3 * clk) / (p.multiProcessorCount * 4));
    return 0;
}
