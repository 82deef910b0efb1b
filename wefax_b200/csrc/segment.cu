// Segment mode (BASELINE.json configs[2]): one long recording decoded as
// overlapping segments, one per GPU.  The only stage that needs more than the existing kernels is
// the GLOBAL percentile pair of wefax.py:196: every rank histograms the median-filtered envelope of
// its core samples by radix digit of the float bit pattern, the host sums the histograms of all
// segments and narrows the four target order statistics digit by digit (11 + 11 + 10 bits), so the
// result is the exact order statistic of the union of the cores with three tiny exchanges.
#include <algorithm>

#include "median.cuh"
#include "stages.cuh"

namespace wefax {

constexpr int kSegHistThreads = 256;

// level 0: every core sample, bin = key >> 21 (into hist[0]).
// level 1: samples with key >> 21 == prefix[t], bin = (key >> 10) & 0x7FF (into hist[t]).
// level 2: samples with key >> 10 == prefix[t], bin = key & 0x3FF (into hist[t]).
// Envelope values are >= 0, so the bit pattern orders like the value.
template <int LEVEL>
__global__ void __launch_bounds__(kSegHistThreads)
seg_hist_kernel(const float *env, long long n, long long core_lo, long long core_hi, uint4 prefix4, uint32_t *hist) {
    constexpr int NH = LEVEL == 0 ? 1 : 4;
    __shared__ uint32_t s_hist[NH][2048];
    for (int i = threadIdx.x; i < NH * 2048; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    const uint32_t prefix[4] = {prefix4.x, prefix4.y, prefix4.z, prefix4.w};
    __syncthreads();

    const long long first = core_lo & ~3ll;
    const long long stride = 4ll * blockDim.x * gridDim.x;
    const long long span = core_hi - first;
    const long long rounds = (span + stride - 1) / stride;   // every thread runs the same number of rounds (match_any)
    for (long long k = 0; k < rounds; ++k) {
        const long long i0 = first + k * stride + 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
        float m[4] = {0.f, 0.f, 0.f, 0.f};
        if (i0 < core_hi) load_med4(env, i0, n, m);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool valid = i0 + j >= core_lo && i0 + j < core_hi;
            const uint32_t key = __float_as_uint(m[j]);
            if (LEVEL == 0) {
                // warp-aggregated: the envelope of an FM signal sits in very few exponent bins
                const uint32_t bin = valid ? (key >> 21) : 0xFFFFFFFFu;
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin);
                if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[0][bin], __popc(peers));
            } else if (valid) {
                const uint32_t hi = key >> (LEVEL == 1 ? 21 : 10);
                const uint32_t bin = LEVEL == 1 ? ((key >> 10) & 0x7FFu) : (key & 0x3FFu);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (hi == prefix[t]) atomicAdd(&s_hist[t % NH][bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NH * 2048; i += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

void launch_segment_hist(wefax_ctx *ctx, const float *env, long long n, long long core_lo, long long core_hi, int level,
                         const uint32_t prefix[4], uint32_t *hist) {
    StageTimer timer(ctx, "seg_hist");
    CUDA_CHECK(cudaMemsetAsync(hist, 0, 4 * 2048 * sizeof(uint32_t), ctx->stream));
    if (core_hi <= core_lo) return;
    const long long quads = (core_hi - (core_lo & ~3ll) + 3) / 4;
    const int blocks = (int)std::max<long long>(
        1, std::min<long long>((quads + kSegHistThreads - 1) / kSegHistThreads, 4ll * ctx->sm_count));
    const uint4 p4 = make_uint4(prefix[0], prefix[1], prefix[2], prefix[3]);
    if (level == 0)
        seg_hist_kernel<0><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, hist);
    else if (level == 1)
        seg_hist_kernel<1><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, hist);
    else
        seg_hist_kernel<2><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, hist);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

// medfilt(env, 5) of the core samples only (demodulated_data of this segment's share)
__global__ void seg_median_kernel(const float *env, long long n, long long core_lo, long long core_hi, float *out) {
    const long long i0 = core_lo + 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= core_hi) return;
    float m[4];
    load_med4(env, i0, n, m);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (i0 + j < core_hi) out[i0 - core_lo + j] = m[j];
}

void launch_segment_median(wefax_ctx *ctx, const float *env, long long n, long long core_lo, long long core_hi,
                           float *out) {
    if (core_hi <= core_lo) return;
    StageTimer timer(ctx, "seg_median");
    const unsigned blocks = (unsigned)((core_hi - core_lo + 1023) / 1024);
    seg_median_kernel<<<blocks, 256, 0, ctx->stream>>>(env, n, core_lo, core_hi, out);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

}  // namespace wefax
