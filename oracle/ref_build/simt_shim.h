/*
 * simt_shim.h — lets the reference's UNMODIFIED src/filter.cu compile as host C++.
 * Force-included (-include) ahead of the reference headers.  It pre-defines the
 * reference's include guard UTILS_H (include/utils.h:2) so that the reference's
 * macro block (KERNEL = __global__, CUDA_FUNC = __forceinline__ __device__, ...;
 * include/utils.h:9-12) is replaced by host equivalents, and declares the SIMT
 * built-ins as thread-local variables that ref_cpu_driver.cpp sets before calling
 * a kernel body once per emulated thread.  __syncthreads() is a no-op, which is
 * exact for depth == 1 with cacheInput == false (no shared-memory traffic and no
 * cross-thread dependency inside one level; src/filter.cu:23-57, 103-157).
 * TEST INFRASTRUCTURE ONLY.
 */
#pragma once
#define UTILS_H 1
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#define KERNEL
#define CUDA_FUNC inline
#define CUDA_CPU_FUNC inline
#define LAUNCHER
typedef unsigned char byte;
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static inline void __syncthreads() {}
void printGPUProperties();
