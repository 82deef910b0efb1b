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
seg_hist_kernel(const float *env, long long n, long long core_lo, long long core_hi, uint4 prefix4,
                const uint32_t *prefix_dev, uint32_t *hist) {
    constexpr int NH = LEVEL == 0 ? 1 : 4;
    __shared__ uint32_t s_hist[NH][2048];
    for (int i = threadIdx.x; i < NH * 2048; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    uint32_t prefix[4] = {prefix4.x, prefix4.y, prefix4.z, prefix4.w};
    if (prefix_dev) {   // device-resident exchange: the prefixes never visit the host
#pragma unroll
        for (int t = 0; t < 4; ++t) prefix[t] = prefix_dev[t];
    }
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
                         const uint32_t prefix[4], uint32_t *hist, const uint32_t *prefix_dev) {
    StageTimer timer(ctx, "seg_hist");
    CUDA_CHECK(cudaMemsetAsync(hist, 0, 4 * 2048 * sizeof(uint32_t), ctx->stream));
    if (core_hi <= core_lo) return;
    const long long quads = (core_hi - (core_lo & ~3ll) + 3) / 4;
    const int blocks = (int)std::max<long long>(
        1, std::min<long long>((quads + kSegHistThreads - 1) / kSegHistThreads, 4ll * ctx->sm_count));
    const uint4 p4 = prefix ? make_uint4(prefix[0], prefix[1], prefix[2], prefix[3]) : make_uint4(0, 0, 0, 0);
    if (level == 0)
        seg_hist_kernel<0><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, prefix_dev, hist);
    else if (level == 1)
        seg_hist_kernel<1><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, prefix_dev, hist);
    else
        seg_hist_kernel<2><<<blocks, kSegHistThreads, 0, ctx->stream>>>(env, n, core_lo, core_hi, p4, prefix_dev, hist);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

// ---- device-resident exchange state (include/wefax_b200.h: WEFAX_SEG_STATE_*) ------------------------------------
__global__ void seg_state_init_kernel(uint32_t *state, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, double t_lo,
                                      double t_hi) {
    if (threadIdx.x == 0) {
        state[0] = r0; state[1] = r1; state[2] = r2; state[3] = r3;
        state[4] = state[5] = state[6] = state[7] = 0u;
        double *d = reinterpret_cast<double *>(state + 8);
        d[0] = t_lo; d[1] = t_hi; d[2] = 0.0; d[3] = 0.0;
        state[16] = 0u;
    }
}

// One CTA of 256 threads: locates, for each of the 4 target ranks, the bin of the SUMMED histogram that holds it
// (thread t owns bins 8t .. 8t+7), narrows rank and prefix; after the last level numpy's lerp gives low / high.
__global__ void __launch_bounds__(256) seg_select_dev_kernel(uint32_t *state, int level) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_bin[4], s_before[4];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int bits = level == 2 ? 10 : 11;
    const uint32_t *hist_all = state + WEFAX_SEG_STATE_HIST;
    for (int t = 0; t < 4; ++t) {
        const uint32_t *hist = hist_all + (level == 0 ? 0 : t) * 2048;
        const uint32_t rank = state[t];
        uint32_t c[8], local = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = hist[8 * tid + j];
            local += c[j];
        }
        uint32_t incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += s_warp[w];
        uint32_t before = woff + incl - local;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (rank >= before && rank < before + c[j]) {
                s_bin[t] = 8 * tid + j;
                s_before[t] = before;
            }
            before += c[j];
        }
        __syncthreads();
    }
    if (tid < 4) {
        state[4 + tid] = (state[4 + tid] << bits) | s_bin[tid];
        state[tid] -= s_before[tid];
    }
    __syncthreads();
    if (level == 2 && tid == 0) {
        double *d = reinterpret_cast<double *>(state + 8);
        const double t_lo = d[0], t_hi = d[1];
        double v[4];
        for (int t = 0; t < 4; ++t) v[t] = (double)__uint_as_float(state[4 + t]);
        const double d0 = __dsub_rn(v[1], v[0]), d1 = __dsub_rn(v[3], v[2]);
        const double low = t_lo >= 0.5 ? __dsub_rn(v[1], __dmul_rn(d0, __dsub_rn(1.0, t_lo))) : __dadd_rn(v[0], __dmul_rn(d0, t_lo));
        const double high = t_hi >= 0.5 ? __dsub_rn(v[3], __dmul_rn(d1, __dsub_rn(1.0, t_hi))) : __dadd_rn(v[2], __dmul_rn(d1, t_hi));
        d[2] = low;
        d[3] = high;
        const double delta = __dsub_rn(high, low);
        if (!(delta > 0.0) || isinf(delta)) state[16] |= WEFAX_REC_NAN;
    }
}

__global__ void seg_state_to_result_kernel(const uint32_t *state, RecResult *res) {
    if (threadIdx.x == 0) {
        const double *d = reinterpret_cast<const double *>(state + 8);
        res->low = d[2];
        res->high = d[3];
        res->status = (int32_t)state[16];
    }
}

void launch_segment_state_init(wefax_ctx *ctx, uint32_t *state, const uint32_t ranks[4], double t_lo, double t_hi) {
    seg_state_init_kernel<<<1, 32, 0, ctx->stream>>>(state, ranks[0], ranks[1], ranks[2], ranks[3], t_lo, t_hi);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void launch_segment_select_dev(wefax_ctx *ctx, uint32_t *state, int level) {
    seg_select_dev_kernel<<<1, 256, 0, ctx->stream>>>(state, level);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void launch_segment_state_to_result(wefax_ctx *ctx, const uint32_t *state, RecResult *res) {
    seg_state_to_result_kernel<<<1, 32, 0, ctx->stream>>>(state, res);
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
