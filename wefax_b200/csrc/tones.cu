// Start / stop tone test of 1-s packets.
//
// Reference semantics: data_packet.py:345-406.  A packet "contains the tone" iff
// scipy.signal.find_peaks(amplitude, distance, height, prominence) of its one-sided,
// max-normalised amplitude spectrum returns min_amount..max_amount peaks, all of them
// between min_frequency and max_frequency.  The spectrum comes from the FFT engine
// (fft_exec.cu: spectrum_natural); this file is find_peaks itself, one CTA per packet:
//
//   1. amplitude |X[k]| / half, block maximum, x = amp / (max + 1e-4)       (data_packet.py:396-404)
//   2. local maxima with plateau midpoints (scipy _local_maxima_1d), height filter
//   3. distance filter (scipy _select_by_peak_distance) as "take the highest remaining
//      candidate, drop every candidate closer than ceil(distance)" — identical to scipy's
//      priority-ordered sweep because removal is symmetric
//   4. prominence with an unbounded window (scipy _peak_prominences), one warp per kept peak
//   5. count / frequency-range decision                                      (data_packet.py:378-385)
//
// Steps 3-5 run twice on the same candidates: start-tone distance, stop-tone distance.
#include "stages.cuh"

namespace wefax {

constexpr int kToneThreads = 256;

struct ToneDev {
    int half;            // bins kept = packet_len / 2
    int cap_cand;        // capacity of the candidate list
    int cap_kept;        // capacity of the kept list
    int dist[2];         // ceil(distance) for start / stop
    double height, prominence, fmin, fmax;
    double bin_hz_div;   // frequency of bin k = k / bin_hz_div   (packet_len / sample_rate)
    int min_amount, max_amount;
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// minimum of x over the run of elements <= h that starts at p and extends in direction dir,
// cooperatively by one warp (scipy: `while 0 <= i <= n-1 and x[i] <= h: track min; i += dir`)
__device__ float run_min(const float *x, int n, int p, int dir, float h, int lane) {
    float m = h;
    int i = p;
    while (true) {
        const int idx = i + dir * lane;
        const bool in = idx >= 0 && idx < n;
        const float v = in ? x[idx] : 0.f;
        const bool stop = !in || v > h;
        const unsigned stops = __ballot_sync(0xffffffffu, stop);
        const int first = stops ? __ffs(stops) - 1 : 32;
        const float mine = (lane < first) ? v : h;
        m = fminf(m, warp_min(mine));
        if (stops) break;
        i += dir * 32;
    }
    return m;
}

__global__ void __launch_bounds__(kToneThreads)
tone_peaks_kernel(const float2 *X, size_t xs, ToneDev t, uint8_t *flags, int32_t *counts) {
    extern __shared__ __align__(16) unsigned char tone_smem[];
    float *x = reinterpret_cast<float *>(tone_smem);
    int *cand = reinterpret_cast<int *>(x + t.half);        // position | removed bit 31
    int *kept = cand + t.cap_cand;
    __shared__ float s_red[kToneThreads / 32];
    __shared__ unsigned long long s_best[kToneThreads / 32];
    __shared__ int s_ncand, s_nkept, s_count, s_bad;
    __shared__ float s_max;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = t.half;
    const float2 *Xp = X + (size_t)blockIdx.x * xs;

    // 1. amplitude and its maximum
    float mx = 0.f;
    const float inv_half = 1.0f / (float)n;
    for (int k = tid; k < n; k += kToneThreads) {
        const float2 v = Xp[k];
        const float a = sqrtf(fmaf(v.x, v.x, v.y * v.y)) * inv_half;
        x[k] = a;
        mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int w = 1; w < kToneThreads / 32; ++w) m = fmaxf(m, s_red[w]);
        s_max = m + 1e-4f;
    }
    __syncthreads();
    const float denom = s_max;
    for (int k = tid; k < n; k += kToneThreads) x[k] = __fdiv_rn(x[k], denom);
    __syncthreads();

    // 2. local maxima (plateau midpoint), height filter
    const int i_max = n - 1;
    for (int i = tid + 1; i < i_max; i += kToneThreads) {
        const float xi = x[i];
        if (x[i - 1] < xi) {
            int ahead = i + 1;
            while (ahead < i_max && x[ahead] == xi) ++ahead;
            if (x[ahead] < xi) {
                const int p = (i + ahead - 1) >> 1;
                if ((double)x[p] >= t.height) {
                    const int slot = atomicAdd(&s_ncand, 1);
                    if (slot < t.cap_cand) cand[slot] = p;
                }
            }
        }
    }
    __syncthreads();
    const int ncand = min(s_ncand, t.cap_cand);

    for (int which = 0; which < 2; ++which) {
        const int dist = t.dist[which];
        for (int c = tid; c < ncand; c += kToneThreads) cand[c] &= 0x7fffffff;
        if (tid == 0) {
            s_nkept = 0;
            s_count = 0;
            s_bad = 0;
        }
        __syncthreads();
        // 3. distance filter: highest remaining candidate first (ties: the later position,
        //    as a stable argsort walked from its end would)
        while (true) {
            unsigned long long best = 0ull;   // (value bits << 32) | (position + 1); values are >= 0
            for (int c = tid; c < ncand; c += kToneThreads) {
                const int e = cand[c];
                if (e >= 0) {
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(x[e]) << 32) | (unsigned)(e + 1);
                    best = key > best ? key : best;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (lane == 0) s_best[warp] = best;
            __syncthreads();
            best = s_best[0];
#pragma unroll
            for (int w = 1; w < kToneThreads / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
            if (best == 0ull) break;   // uniform: nothing left
            const int p = (int)(unsigned)(best & 0xffffffffull) - 1;
            for (int c = tid; c < ncand; c += kToneThreads) {
                const int e = cand[c];
                if (e >= 0 && abs(e - p) < dist) cand[c] = e | (int)0x80000000;   // includes p itself
            }
            if (tid == 0) {
                if (s_nkept < t.cap_kept) kept[s_nkept] = p;
                s_nkept++;
            }
            __syncthreads();
        }
        __syncthreads();
        const int nkept = min(s_nkept, t.cap_kept);
        // 4./5. prominence, count, frequency range
        for (int j = warp; j < nkept; j += kToneThreads / 32) {
            const int p = kept[j];
            const float h = x[p];
            const float lmin = run_min(x, n, p, -1, h, lane);
            const float rmin = run_min(x, n, p, +1, h, lane);
            if (lane == 0) {
                const double prom = (double)h - (double)fmaxf(lmin, rmin);
                if (prom >= t.prominence) {
                    atomicAdd(&s_count, 1);
                    const double f = (double)p / t.bin_hz_div;
                    if (!(t.fmin <= f && f <= t.fmax)) atomicOr(&s_bad, 1);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int cnt = s_count;
            const bool overflow = s_ncand > t.cap_cand || s_nkept > t.cap_kept;   // cannot happen by construction
            flags[(size_t)blockIdx.x * 2 + which] =
                (!overflow && !s_bad && cnt >= t.min_amount && cnt <= t.max_amount) ? 1 : 0;
            counts[(size_t)blockIdx.x * 2 + which] = overflow ? -1 : cnt;
        }
        __syncthreads();
    }
}

// flags / counts: 2 per packet (start, stop), device pointers
void launch_tone_peaks(wefax_ctx *ctx, const float2 *X, size_t xs, long long packet_len, int sample_rate,
                       int n_packets, const wefax_tone_settings &s, uint8_t *flags, int32_t *counts) {
    ToneDev t{};
    t.half = (int)(packet_len / 2);
    t.dist[0] = (int)ceil(s.start_distance);
    t.dist[1] = (int)ceil(s.stop_distance);
    if (t.dist[0] < 1 || t.dist[1] < 1) WEFAX_THROW(WEFAX_ERR_INVALID, "tone peak distance must be >= 1");
    t.cap_cand = t.half / 2 + 1;
    const int dmin = t.dist[0] < t.dist[1] ? t.dist[0] : t.dist[1];
    t.cap_kept = t.half / dmin + 2;
    if (t.cap_kept > t.cap_cand) t.cap_kept = t.cap_cand;
    t.height = s.height;
    t.prominence = s.prominence;
    t.fmin = s.min_frequency;
    t.fmax = s.max_frequency;
    t.bin_hz_div = (double)packet_len / (double)sample_rate;
    t.min_amount = s.min_amount;
    t.max_amount = s.max_amount;
    const size_t smem = (size_t)t.half * sizeof(float) + ((size_t)t.cap_cand + t.cap_kept) * sizeof(int);
    if (smem > 220 * 1024)
        WEFAX_THROW(WEFAX_ERR_UNSUPPORTED, "tone packet of %lld frames needs %zu bytes of shared memory (max 220 KiB)",
                    packet_len, smem);
    const void *fn = (const void *)tone_peaks_kernel;
    if (!ctx->smem_configured.count(fn)) {
        CUDA_CHECK(cudaFuncSetAttribute(tone_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        ctx->smem_configured[fn] = 1;
    }
    StageTimer timer(ctx, "tone_peaks");
    tone_peaks_kernel<<<n_packets, kToneThreads, smem, ctx->stream>>>(X, xs, t, flags, counts);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

}  // namespace wefax
