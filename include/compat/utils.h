// include/compat/utils.h — source-compatible stand-in for the reference's include/utils.h (macros :9-12, byte :15,
// printGPUProperties :26).  Written from the interface, not copied: only what code on the denoise path uses.
#pragma once
#ifndef RMD_COMPAT_UTILS_H
#define RMD_COMPAT_UTILS_H

#include <cuda_runtime.h>

#include <chrono>
#include <iostream>
#include <stdexcept>
#include <string>

#define KERNEL __global__
#define CUDA_FUNC __forceinline__ __device__
#define CUDA_CPU_FUNC __forceinline__ __device__ __host__
#define LAUNCHER

typedef unsigned char byte;

// The reference's CHECK_CUDA wraps the DRIVER api and is never used (include/utils.h:17-24).  This one checks the
// runtime api and throws what the reference's harness catches (src/test.cu:40-42).
#define RMD_CHECK_CUDA(call)                                                                            \
    do {                                                                                                \
        cudaError_t rmd_err_ = (call);                                                                  \
        if (rmd_err_ != cudaSuccess)                                                                    \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(rmd_err_) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                       \
    } while (0)

void printGPUProperties();

#endif
