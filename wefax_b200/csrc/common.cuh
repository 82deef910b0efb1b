// Shared declarations of the wefax_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/wefax_b200.h"

namespace wefax {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};

#define WEFAX_THROW(code_, ...)                                  \
    do {                                                         \
        char buf_[512];                                          \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);               \
        throw ::wefax::Error{(code_), std::string(buf_)};        \
    } while (0)

#define CUDA_CHECK(expr)                                                          \
    do {                                                                          \
        cudaError_t e_ = (expr);                                                  \
        if (e_ != cudaSuccess)                                                    \
            WEFAX_THROW(e_ == cudaErrorMemoryAllocation ? WEFAX_ERR_NOMEM         \
                                                        : WEFAX_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                      \
    } while (0)

// ---------------------------------------------------------------------------
// exact unsigned 32-bit division by a run-time constant (Granlund-Montgomery)
// ---------------------------------------------------------------------------
struct FastDiv {
    uint32_t d, m, s1, s2;
    void init(uint32_t d_) {
        d = d_;
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
        s1 = l < 1 ? l : 1;
        s2 = l ? l - 1 : 0;
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        uint32_t t = __umulhi(m, n);
#else
        uint32_t t = (uint32_t)(((uint64_t)m * n) >> 32);
#endif
        return (t + ((n - t) >> s1)) >> s2;
    }
};

// ---------------------------------------------------------------------------
// device buffer that grows on demand (context-owned scratch)
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    void *reserve(size_t bytes) {
        if (bytes > cap) {
            release();
            size_t want = bytes + (bytes >> 3);
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                p = nullptr;
                (void)cudaGetLastError();
                WEFAX_THROW(WEFAX_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", want,
                            cudaGetErrorString(e));
            }
            cap = want;
        }
        return p;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

}  // namespace wefax
