// include/compat/vector.h — source-compatible CudaVector<T> / CpuVector<T> (reference include/vector.h): RAII device
// array with synchronous and stream-ordered copies.  Same member names and argument meaning; unlike the reference
// (which checks no CUDA call, include/vector.h:119-126) every failed copy throws std::runtime_error, and the type is
// movable and non-copyable instead of double-freeing on copy.  A failed ALLOCATION does not throw from the constructor
// (the reference's harness constructs its planes during static initialisation, src/test.cu:65-66, where an exception
// cannot be caught): the vector stays empty (data() == nullptr, size() == 0), error() names the reason and the first
// copy throws it.
#pragma once
#ifndef RMD_COMPAT_VECTOR_H
#define RMD_COMPAT_VECTOR_H

#include <cuda_runtime.h>

#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

template <typename T>
using CpuVector = std::vector<T>;

template <typename T>
struct CudaVector {
    CudaVector() = default;
    explicit CudaVector(size_t size) : size_p(size) { alloc(); }
    CudaVector(T* v, size_t size) : size_p(size) {
        alloc();
        copyFrom(v, size);
    }
    explicit CudaVector(CpuVector<T>& v) : CudaVector(v.data(), v.size()) {}
    CudaVector(const CudaVector&) = delete;
    CudaVector& operator=(const CudaVector&) = delete;
    CudaVector(CudaVector&& o) noexcept : size_p(o.size_p), data_p(o.data_p), error_p(o.error_p) { o.size_p = 0; o.data_p = nullptr; }
    CudaVector& operator=(CudaVector&& o) noexcept {
        if (this != &o) {
            release();
            size_p = o.size_p; data_p = o.data_p; error_p = o.error_p;
            o.size_p = 0; o.data_p = nullptr;
        }
        return *this;
    }
    ~CudaVector() { release(); }

    T* data() { return data_p; }
    const T* data() const { return data_p; }
    size_t size() const { return size_p; }
    cudaError_t error() const { return error_p; }  // cudaSuccess unless the allocation failed

    void copyFrom(T* v, size_t size) {
        check(error_p, "cudaMalloc");
        if (size > size_p) throw std::runtime_error("Size mismatch");
        check(cudaMemcpy(data_p, v, size * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy H2D");
    }
    void copyFromAsync(T* v, size_t size, cudaStream_t stream) {
        check(error_p, "cudaMalloc");
        if (size > size_p) throw std::runtime_error("Size mismatch");
        check(cudaMemcpyAsync(data_p, v, size * sizeof(T), cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D");
    }
    void copyTo(T* v) {
        check(error_p, "cudaMalloc");
        check(cudaMemcpy(v, data_p, size_p * sizeof(T), cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
    }
    void copyToAsync(T* v, cudaStream_t stream) {
        check(error_p, "cudaMalloc");
        check(cudaMemcpyAsync(v, data_p, size_p * sizeof(T), cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H");
    }

   private:
    size_t size_p = 0;
    T* data_p = nullptr;
    cudaError_t error_p = cudaSuccess;
    static void check(cudaError_t e, const char* what) {
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    }
    void alloc() {
        if (!size_p) return;
        error_p = cudaMalloc(reinterpret_cast<void**>(&data_p), size_p * sizeof(T));
        if (error_p != cudaSuccess) { data_p = nullptr; size_p = 0; }
    }
    void release() {
        if (data_p) cudaFree(data_p);
        data_p = nullptr;
        size_p = 0;
    }
};

#endif
