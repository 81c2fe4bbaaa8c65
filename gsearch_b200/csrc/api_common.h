// api_common.h -- host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsearch_b200.h"

namespace gsb {

void set_error(const char *fmt, ...);

#define GSB_CUDA_TRY(expr)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            gsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                         \
            return e__ == cudaErrorMemoryAllocation ? GSB_ERR_OOM : GSB_ERR_CUDA;             \
        }                                                                                     \
    } while (0)

// growable device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, bool zero_new = false) {
        if (bytes <= cap) return GSB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return GSB_ERR_OOM;
        }
        cap = want;
        if (zero_new) cudaMemset(p, 0, want);
        return GSB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// growable pinned host buffer
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return GSB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
            return GSB_ERR_OOM;
        }
        cap = want;
        return GSB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

int check_device(int device);

}  // namespace gsb
