// N4: a true WEFAX demodulator as an EXTENSION of the reference path.
//
// The reference decodes by slope detection (notch + envelope, wefax.py:63-74) and rasters
// int(samples-per-line) columns.  Its README (README.md:85-101) describes what a radiofax receiver
// really does: the grey level is the instantaneous frequency (black 1500 Hz, white 2300 Hz), a
// line lasts 60/LPM seconds exactly (5512.5 samples at 120 LPM), and a line has pi*IOC pixels.
// None of that exists in the reference's code, so there is no reference parity here: this mode is
// off by default, is checked against a float64 restatement (oracle/fm_oracle.py) and against the
// ground truth of the synthetic generator, and never touches the reference-exact path.
//
//   band-limit   zero-phase FIR band-pass around the 1900 Hz +- 400 Hz carrier: the notch kernel
//                (stages.cu: filtfilt_kernel) run with windowed-sinc taps instead of the notch's
//   analytic     y = Hilbert(x) from the same real-input transform (fft_exec.cu), y stored
//   FM demod     phase difference of consecutive analytic samples -> Hz -> grey in [0, 1]
//   phasing      fold `fold_lines` lines of grey modulo the exact line length, boxcar of the 5 %
//                white pulse, arg max = line start (block-wide reduction)
//   image        pixel (row, col) = box average of grey over its exact fractional sample span,
//                one thread per pixel: demod -> grey map -> line resample in one pass over x, y
#include <cmath>
#include <vector>

#include "stages.cuh"

namespace wefax {

// windowed-sinc (Hamming) band-pass, unit gain at the band centre; applied forwards and backwards
FirParams make_bandpass_fir(double lo_hz, double hi_hz, double fs, int taps) {
    if (taps < 3 || taps > kMaxFirTaps - 1 || !(lo_hz > 0.0) || !(hi_hz > lo_hz) || !(hi_hz < fs / 2))
        WEFAX_THROW(WEFAX_ERR_INVALID, "bad band-pass parameters");
    if ((taps & 1) == 0) taps -= 1;
    std::vector<double> h(taps);
    const double fl = lo_hz / fs, fh = hi_hz / fs, mid = (taps - 1) / 2.0, fc = 0.5 * (fl + fh);
    double gr = 0.0, gi = 0.0;
    for (int i = 0; i < taps; ++i) {
        const double t = i - mid;
        const double ideal = t == 0.0 ? 2.0 * (fh - fl) : (sin(2.0 * M_PI * fh * t) - sin(2.0 * M_PI * fl * t)) / (M_PI * t);
        const double win = 0.54 - 0.46 * cos(2.0 * M_PI * i / (taps - 1));
        h[i] = ideal * win;
        gr += h[i] * cos(2.0 * M_PI * fc * t);
        gi += h[i] * sin(2.0 * M_PI * fc * t);
    }
    const double gain = sqrt(gr * gr + gi * gi);
    FirParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.K = taps;
    const int sizes[] = {16, 20, 24, 32, 48, 64};
    for (int s : sizes)
        if (taps <= s) {
            fp.KP = s;
            break;
        }
    for (int i = 0; i < taps; ++i) fp.h[i] = (float)(h[i] / gain);
    for (int j = 0; j < fp.KP; ++j) fp.hr[j] = fp.h[fp.KP - 1 - j];
    return fp;
}

// grey[i] = (f_inst - black) / (white - black), f_inst from the phase step z[i] * conj(z[i-1])
__global__ void __launch_bounds__(256)
fm_grey_kernel(const float *x, const float *y, float *g, long long n, float hz_per_rad, float black, float inv_span) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long j = i > 0 ? i - 1 : 0, k = i > 0 ? i : 1;   // sample 0 repeats the first step
    const float xr = __ldg(x + k), yr = __ldg(y + k), xp = __ldg(x + j), yp = __ldg(y + j);
    const float re = fmaf(xr, xp, yr * yp), im = fmaf(yr, xp, -xr * yp);
    const float f = atan2f(im, re) * hz_per_rad;
    g[i] = (f - black) * inv_span;
}

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// P[o] = sum over lines of clip(grey[from + floor(l * Ls + o)]), o < Lc = ceil(Ls)
__global__ void __launch_bounds__(256)
fm_fold_kernel(const float *g, long long n, long long from, double Ls, int lines, int Lc, float *P) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Lc) return;
    float acc = 0.f;
    for (int l = 0; l < lines; ++l) {
        const long long i = from + (long long)floor((double)l * Ls + (double)o);
        if (i < n) acc += clip01(__ldg(g + i));
    }
    P[o] = acc;
}

// line start = from + arg max_o sum_{j < wb} P[(o + j) mod Lc]  (smallest o on ties); one CTA
__global__ void __launch_bounds__(1024)
fm_phase_kernel(const float *P, int Lc, int wb, long long from, long long *line_start) {
    extern __shared__ double pre[];    // pre[i] = sum_{j < i} P[j mod Lc], i <= 2*Lc (two periods: circular boxcar)
    __shared__ double s_part[1024];
    __shared__ unsigned long long s_best[32];
    const int tid = threadIdx.x, nt = blockDim.x, total = 2 * Lc;
    // chunked scan: each thread sums a contiguous chunk, thread 0 scans the chunk totals
    const int chunk = (total + nt - 1) / nt;
    const int b = tid * chunk, e = min(b + chunk, total);
    double sum = 0.0;
    for (int i = b; i < e; ++i) sum += (double)P[i % Lc];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int t = 0; t < nt; ++t) {
            const double v = s_part[t];
            s_part[t] = run;
            run += v;
        }
    }
    __syncthreads();
    double run = s_part[tid];
    for (int i = b; i < e; ++i) {
        pre[i] = run;
        run += (double)P[i % Lc];
    }
    if (tid == nt - 1) pre[total] = run;
    __syncthreads();
    // key: score as ordered bits in the high word, (0x7fffffff - o) in the low word
    unsigned long long best = 0ull;
    for (int o = tid; o < Lc; o += nt) {
        const float score = (float)(pre[o + wb] - pre[o]);   // >= 0
        const unsigned long long key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)(0x7fffffff - o);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, d);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) s_best[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nt / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
        *line_start = from + (long long)(0x7fffffff - (int)(unsigned)(best & 0xffffffffull));
    }
}

// pixel (r, p): box average of clip(grey) over [a, a + Ls / W), a = line_start + r * Ls + p * Ls / W
__global__ void __launch_bounds__(256)
fm_image_kernel(const float *g, long long n, const long long *line_start, double Ls, int W, long long image_end,
                uint8_t *img) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const double ls = (double)*line_start;
    // rows = floor((image_end - line_start) / Ls): the host sizes the grid for the largest possible count
    if (p >= W || ls + (double)(r + 1) * Ls > (double)image_end) return;
    const double step = Ls / (double)W;
    const double a = ls + (double)r * Ls + (double)p * step, b = a + step;
    const long long i0 = (long long)floor(a), i1 = (long long)ceil(b);
    float acc = 0.f;
    for (long long i = i0; i < i1; ++i) {
        const double lo = fmax(a, (double)i), hi = fmin(b, (double)(i + 1));
        const float wgt = (float)(hi - lo);
        if (i >= 0 && i < n && wgt > 0.f) acc = fmaf(wgt, clip01(__ldg(g + i)), acc);
    }
    const float v = rintf(255.f * acc / (float)step);
    img[(size_t)r * W + p] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
}

void launch_fm_grey(wefax_ctx *ctx, const float *x, const float *y, float *g, long long n, double black_hz,
                    double white_hz) {
    StageTimer timer(ctx, "fm_grey");
    const float hz_per_rad = (float)((double)WEFAX_TARGET_RATE / (2.0 * M_PI));
    fm_grey_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, y, g, n, hz_per_rad, (float)black_hz,
                                                                       (float)(1.0 / (white_hz - black_hz)));
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void launch_fm_phasing(wefax_ctx *ctx, const float *g, long long n, long long from, double Ls, int lines, float *P,
                       long long *d_line_start) {
    StageTimer timer(ctx, "fm_phasing");
    const int Lc = (int)ceil(Ls);
    const int wb = std::max(1, (int)llround(0.05 * Ls));
    fm_fold_kernel<<<(Lc + 255) / 256, 256, 0, ctx->stream>>>(g, n, from, Ls, lines, Lc, P);
    const size_t smem = (size_t)(2 * Lc + 1) * sizeof(double);
    if (smem > 200 * 1024) WEFAX_THROW(WEFAX_ERR_UNSUPPORTED, "line of %d samples is too long for the phasing search", Lc);
    const void *fn = (const void *)fm_phase_kernel;
    if (!ctx->smem_configured.count(fn)) {
        CUDA_CHECK(cudaFuncSetAttribute(fm_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->smem_configured[fn] = 1;
    }
    fm_phase_kernel<<<1, 1024, smem, ctx->stream>>>(P, Lc, wb, from, d_line_start);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches += 2;
}

void launch_fm_image(wefax_ctx *ctx, const float *g, long long n, const long long *d_line_start, double Ls, int W,
                     int rows_max, long long image_end, uint8_t *img) {
    if (rows_max <= 0) return;
    StageTimer timer(ctx, "fm_image");
    dim3 grid((W + 255) / 256, rows_max);
    fm_image_kernel<<<grid, 256, 0, ctx->stream>>>(g, n, d_line_start, Ls, W, image_end, img);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

}  // namespace wefax
