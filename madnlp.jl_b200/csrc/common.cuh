// Shared helpers for the b200kkt CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/b200kkt.h"

namespace b2 {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define B2_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ::b2::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define B2_CUDA_THROW(call)                                                   \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) { ::b2::cuda_fail(e__, #call, __FILE__, __LINE__); throw std::runtime_error("cuda"); } \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const T* h, size_t count) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    size_t bytes() const { return n * sizeof(T); }
};

inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// number of SMs of the current device (cached)
int sm_count();

}  // namespace b2
