// Grey-level helpers shared by the fused grey map + raster kernel (greyraster.cu) and the sequential phasing
// scan (stages.cu), which may need grey levels that do not exist yet.
#pragma once

#include "median.cuh"
#include "stages.cuh"

namespace wefax {

// exact level from the table alone (T is non-decreasing): largest k with T[k] <= bits
template <class TT>
__device__ __forceinline__ int grey_from_table(const TT *T, uint32_t bits) {
    int k = 0;
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1)
        if (T[k + step] <= bits) k += step;
    return k;
}

// grey level of sample i of one recording straight from its envelope (median-5 with zero padding, wefax.py:175)
__device__ __forceinline__ int grey_at_from_envelope(const float *e, long long i, long long n, const GreyTable *tab) {
    float w[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const long long k = i - 2 + j;
        w[j] = (k >= 0 && k < n) ? __ldg(e + k) : 0.f;
    }
    return grey_from_table(tab->T, __float_as_uint(med5(w[0], w[1], w[2], w[3], w[4])));
}

}  // namespace wefax
