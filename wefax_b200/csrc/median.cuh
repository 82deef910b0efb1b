// median-of-5 loaders shared by the grey-map stages (stages.cu) and the segment-mode kernels (segment.cu)
#pragma once

#include <cstdint>

namespace wefax {

// ===========================================================================
// median of 5 with zero padding == scipy.signal.medfilt(v, 5)   (wefax.py:175)
// ===========================================================================
__device__ __forceinline__ float med3(float a, float b, float c) {
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}
__device__ __forceinline__ float med5(float a, float b, float c, float d, float e) {
    float f = fmaxf(fminf(a, b), fminf(c, d));   // the two middle values of {a,b,c,d}
    float g = fminf(fmaxf(a, b), fmaxf(c, d));
    return med3(e, f, g);
}

// medians (window MED = 5, or 3 for the live path's packets: data_packet.py:440) of elements i0..i0+3 of one
// recording (zero outside [0, n))
template <int MED = 5>
__device__ __forceinline__ void load_med4(const float *e, long long i0, long long n, float out[4]) {
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        long long i = i0 - 2 + j;
        w[j] = (i >= 0 && i < n) ? __ldg(e + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        out[j] = MED == 5 ? med5(w[j], w[j + 1], w[j + 2], w[j + 3], w[j + 4]) : med3(w[j + 1], w[j + 2], w[j + 3]);
}

// medians of elements i0..i0+7 (i0 a multiple of 8); interior 16-byte aligned runs
// come in as four 128-bit loads
template <int MED = 5>
__device__ __forceinline__ void load_med8(const float *e, long long i0, long long n, float out[8]) {
    float w[12];   // elements i0-2 .. i0+9
    if (i0 >= 4 && i0 + 12 <= n && (reinterpret_cast<uintptr_t>(e + i0) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(e + i0 - 4));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(e + i0));
        const float4 c = __ldg(reinterpret_cast<const float4 *>(e + i0 + 4));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(e + i0 + 8));
        w[0] = a.z; w[1] = a.w;
        w[2] = b.x; w[3] = b.y; w[4] = b.z; w[5] = b.w;
        w[6] = c.x; w[7] = c.y; w[8] = c.z; w[9] = c.w;
        w[10] = d.x; w[11] = d.y;
    } else {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const long long i = i0 - 2 + j;
            w[j] = (i >= 0 && i < n) ? __ldg(e + i) : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        out[j] = MED == 5 ? med5(w[j], w[j + 1], w[j + 2], w[j + 3], w[j + 4]) : med3(w[j + 1], w[j + 2], w[j + 3]);
}

// medians of x[OFF + 2 .. OFF + 9] given x[OFF .. OFF + 11]: every adjacent pair is ordered once and shared
// by the two windows it belongs to (same min / max expression tree as med5)
template <int OFF>
__device__ __forceinline__ void med8_from16(const float (&f)[16], float (&m)[8]) {
    float lo[10], hi[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        lo[j] = fminf(f[OFF + j], f[OFF + j + 1]);
        hi[j] = fmaxf(f[OFF + j], f[OFF + j + 1]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float a = fmaxf(lo[k], lo[k + 2]);
        const float b = fminf(hi[k], hi[k + 2]);
        m[k] = med3(f[OFF + k + 4], a, b);
    }
}

// medians of x[OFF + 2 .. OFF + 5] given x[OFF .. OFF + 7] (OFF <= 4, 12 values loaded)
template <int OFF>
__device__ __forceinline__ void med4_from12(const float (&f)[12], float (&m)[4]) {
    float lo[6], hi[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        lo[j] = fminf(f[OFF + j], f[OFF + j + 1]);
        hi[j] = fmaxf(f[OFF + j], f[OFF + j + 1]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = fmaxf(lo[k], lo[k + 2]);
        const float b = fminf(hi[k], hi[k + 2]);
        m[k] = med3(f[OFF + k + 4], a, b);
    }
}

}  // namespace wefax
