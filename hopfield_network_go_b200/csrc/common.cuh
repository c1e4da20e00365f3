// common.cuh — shared helpers for libhopfield_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/hopfield_b200.h"

namespace hn {

// thread-local last error message (hn_last_error)
void set_error(const char* fmt, ...);

#define HN_CUDA_TRY(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            hn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                       \
            return _e == cudaErrorMemoryAllocation ? HN_E_OOM : HN_E_CUDA;                 \
        }                                                                                  \
    } while (0)

#define HN_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc != HN_OK) return _rc; \
    } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Simple growable device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return HN_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return HN_E_OOM;
        }
        cap = want;
        return HN_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31; return x;
}

}  // namespace hn
