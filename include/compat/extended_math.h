// include/compat/extended_math.h — the three index helpers of the reference's include/extended_math.h that define
// the memory layout of the denoise path (totalSize :54-60, inRange :62-64, flattenIndex :66-68: row-major, pitch = W,
// no padding) plus the int2 arithmetic the kernels use from helper_math.  The reference's uchar3 helpers are dead
// code there (and its operator-(uchar3, uchar3) adds, :6-8); they are not reproduced.
#pragma once
#ifndef RMD_COMPAT_EXTENDED_MATH_H
#define RMD_COMPAT_EXTENDED_MATH_H

#include "utils.h"

// int2 / int3 arithmetic: the reference gets these from NVIDIA's helper_math.h.  When that header is on the include
// path as well (building the reference's own sources against this directory) it wins and these are skipped.
#if !defined(HELPER_MATH_H) && !defined(RMD_COMPAT_NO_VECTOR_OPS)
CUDA_CPU_FUNC int2 operator+(int2 a, int2 b) { return make_int2(a.x + b.x, a.y + b.y); }
CUDA_CPU_FUNC int2 operator-(int2 a, int2 b) { return make_int2(a.x - b.x, a.y - b.y); }
CUDA_CPU_FUNC int2 operator*(int2 a, int2 b) { return make_int2(a.x * b.x, a.y * b.y); }
CUDA_CPU_FUNC int2 operator*(int2 a, int b) { return make_int2(a.x * b, a.y * b); }
CUDA_CPU_FUNC int2 operator*(int b, int2 a) { return make_int2(a.x * b, a.y * b); }
#endif

CUDA_CPU_FUNC int totalSize(int2 shape) { return shape.x * shape.y; }
CUDA_CPU_FUNC int totalSize(int3 shape) { return shape.x * shape.y * shape.z; }
CUDA_CPU_FUNC int inRange(int2 pos, int2 shape) { return pos.x >= 0 && pos.x < shape.x && pos.y >= 0 && pos.y < shape.y; }
CUDA_CPU_FUNC int flattenIndex(int2 p, int2 shape) { return p.y * shape.x + p.x; }

#endif
