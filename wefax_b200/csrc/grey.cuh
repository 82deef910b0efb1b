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

// exact grey level of a median value (same operations, in the same order, as numpy: wefax.py:197-200, 216)
__device__ __forceinline__ int grey_level_exact(float m, double low, double delta) {
    float v = (float)rint(__ddiv_rn(__dmul_rn(255.0, __dsub_rn((double)m, low)), delta));
    v = fminf(fmaxf(v, 0.f), 255.f);
    return (int)v;
}

__device__ __forceinline__ int grey_estimate(float m, float scale, float off) {
    const int k = __float2int_rn(fmaf(m, scale, off));
    return min(max(k, 0), 255);
}

// The threshold table of one recording, by all threads of a CTA (>= 256; thread k < 256 finds the threshold of level k).
__device__ __forceinline__ void build_grey_table(const RecResult *res, GreyTable *tab) {
    const double low = res->low, delta = __dsub_rn(res->high, low);
    const int k = threadIdx.x;
    if (k >= 256) {                      // (only votes)
        __syncthreads_and(1);
        return;
    }
    const uint32_t kInf = 0x7F800000u;
    uint32_t t = 0u;
    if (k >= 1) {
        if (grey_level_exact(__uint_as_float(kInf), low, delta) < k) {
            t = 0xFFFFFFFFu;   // no float reaches this level
        } else {
            // the step lies within a few ulps of low + (k - 1/2) * delta / 255: gallop away from that guess until the
            // step is bracketed, then bisect (about 6 evaluations of the exact level instead of 31)
            auto level = [&](uint32_t bits) { return grey_level_exact(__uint_as_float(bits), low, delta); };
            const float guess = (float)(low + ((double)k - 0.5) * delta / 255.0);
            uint32_t b = guess >= 0.f ? __float_as_uint(guess) : 0u;     // (NaN compares false)
            if (b > kInf) b = kInf;
            uint32_t lo, hi, step = 1u;
            if (level(b) >= k) {
                hi = b;
                lo = 0u;
                while (hi > 0u) {
                    const uint32_t c = hi > step ? hi - step : 0u;
                    if (level(c) >= k) {
                        hi = c;
                        step <<= 1;
                    } else {
                        lo = c + 1u;
                        break;
                    }
                }
            } else {
                lo = b + 1u;
                hi = kInf;
                while (true) {
                    const uint32_t c = (kInf - b > step) ? b + step : kInf;
                    if (c == kInf || level(c) >= k) {
                        hi = c;
                        break;
                    }
                    lo = c + 1u;
                    step <<= 1;
                }
            }
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (grey_level_exact(__uint_as_float(mid), low, delta) >= k) hi = mid;
                else lo = mid + 1u;
            }
            t = lo;
        }
    }
    tab->T[k] = t;
    if (k == 0) {
        tab->T[256] = 0xFFFFFFFFu;
        tab->T[257] = tab->T[258] = tab->T[259] = 0xFFFFFFFFu;
    }
    // fp32 estimate of the level and the proof that it is never off by more than one: the estimate is monotone
    // in m, the level is a step function, so it is enough to look at both sides of every step (and at +inf)
    const float scale = (float)(255.0 / delta), off = (float)(-low * (255.0 / delta));
    bool ok = scale >= 0.f && isfinite(scale) && isfinite(off);
    auto near = [&](uint32_t bits) {
        const float m = __uint_as_float(bits);
        const int d = grey_estimate(m, scale, off) - grey_level_exact(m, low, delta);
        return d >= -1 && d <= 1;
    };
    if (k >= 1) {
        if (t == 0xFFFFFFFFu) ok = false;
        else {
            ok = ok && near(t);
            if (t > 0u) ok = ok && near(t - 1u);
        }
    } else {
        ok = ok && near(kInf) && near(0u);
    }
    const int all_ok = __syncthreads_and(ok ? 1 : 0);
    if (k == 0) {
        tab->scale = scale;
        tab->off = off;
        tab->est_ok = all_ok;
        tab->pad = 0;
    }
}


}  // namespace wefax
