// Host-side FFT planning: factor n into pass lengths, size the shared-memory
// tiles, build the twiddle / permutation tables on the device.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "fft.cuh"

namespace wefax {

namespace {

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// tile budget in complex elements (64 KiB of float2 => 3 CTAs / SM)
int tile_max() { return env_int("WEFAX_FFT_TILE", 8192); }
// strided passes keep at least this many adjacent columns per tile (coalescing)
int min_cols() { return env_int("WEFAX_FFT_MINC", 8); }

bool smooth13(long long n) {
    for (int p : {2, 3, 5, 7, 11, 13})
        while (n % p == 0) n /= p;
    return n == 1;
}

std::vector<long long> divisors_upto(long long n, long long cap) {
    std::vector<long long> small, out;
    for (long long d = 1; d * d <= n; ++d)
        if (n % d == 0) {
            if (d <= cap) out.push_back(d);
            long long e = n / d;
            if (e != d && e <= cap) out.push_back(e);
        }
    std::sort(out.begin(), out.end());
    return out;
}

struct Search {
    long long cap_s, cap_l;
    double best_cost = 1e300;
    std::vector<int> best;
    std::vector<int> cur;
    // strided factors are chosen non-decreasing; the remainder is the last pass
    void go(long long rem, int left, long long min_r) {
        if (left == 0) {
            if (rem > cap_l || rem < 2) return;
            double cost = (double)rem;
            for (int r : cur) cost = std::max(cost, 4.0 * r);
            if (cost < best_cost) {
                best_cost = cost;
                best = cur;
                best.push_back((int)rem);
            }
            return;
        }
        for (long long d : divisors_upto(rem, cap_s)) {
            if (d < min_r || d < 2) continue;
            cur.push_back((int)d);
            go(rem / d, left - 1, d);
            cur.pop_back();
        }
    }
};

std::vector<int> stage_radices(int R) {
    std::vector<int> r;
    for (int p : {8, 4, 2, 3, 5, 7, 11, 13})
        while (R % p == 0) {
            r.push_back(p);
            R /= p;
        }
    return r;
}

__global__ void twiddle_table_kernel(float2 *dst, int count, double L, double mult) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // w_L^(i*mult) = exp(-2*pi*i * (i*mult mod L) / L), evaluated in double
    double e = fmod((double)i * mult, L);
    double s, c;
    sincospi(-2.0 * e / L, &s, &c);
    dst[i] = make_float2((float)c, (float)s);
}

}  // namespace

bool plan_factors(long long n, std::vector<int> &Rs) {
    Rs.clear();
    if (n < 1) return false;
    if (n == 1) {
        Rs.push_back(1);
        return true;
    }
    if (!smooth13(n)) return false;
    const int tmax = tile_max();
    const long long cap_l = tmax;
    const long long cap_s = std::max(2, tmax / min_cols());
    int forced = env_int("WEFAX_FFT_PASSES", 0);
    for (int P = 1; P <= kMaxPasses; ++P) {
        if (forced && P != forced) continue;
        Search s;
        s.cap_s = cap_s;
        s.cap_l = cap_l;
        s.go(n, P - 1, 2);
        if (!s.best.empty()) {
            Rs = s.best;
            return true;
        }
    }
    return false;
}

long long next_smooth_length(long long m) {
    long long best = -1;
    // enumerate 2^a 3^b 5^c 7^d in [m, 2m)
    for (long long p7 = 1; p7 < 2 * m; p7 *= 7)
        for (long long p5 = p7; p5 < 2 * m; p5 *= 5)
            for (long long p3 = p5; p3 < 2 * m; p3 *= 3) {
                long long v = p3;
                while (v < m) v *= 2;
                if (best < 0 || v < best) {
                    std::vector<int> tmp;
                    if (plan_factors(v, tmp)) best = v;
                }
            }
    return best;
}

std::unique_ptr<FftPlan> make_plan(long long n, cudaStream_t stream) {
    std::vector<int> Rs;
    if (!plan_factors(n, Rs)) return nullptr;
    auto plan = std::make_unique<FftPlan>();
    plan->n = n;
    plan->npass = (int)Rs.size();
    const int P = plan->npass;
    for (int i = 0; i < P; ++i) plan->Rs[i] = Rs[i];
    for (int i = P - 1; i >= 0; --i) plan->S[i] = (i == P - 1) ? 1 : plan->S[i + 1] * Rs[i + 1];

    // table layout
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~size_t(255);
        return o;
    };
    size_t o_twR[kMaxPasses], o_perm[kMaxPasses], o_lo[kMaxPasses], o_hi[kMaxPasses];
    long long n_hi[kMaxPasses];
    for (int i = 0; i < P; ++i) {
        o_twR[i] = take((size_t)Rs[i] * sizeof(float2));
        o_perm[i] = take((size_t)Rs[i] * sizeof(uint16_t));
        long long L = (long long)Rs[i] * plan->S[i];
        n_hi[i] = (L >> kTwLoBits) + 1;
        o_lo[i] = take(sizeof(float2) << kTwLoBits);
        o_hi[i] = take((size_t)n_hi[i] * sizeof(float2));
    }
    char *base = (char *)plan->tables.reserve(off);

    const int tmax = tile_max();
    for (int i = 0; i < P; ++i) {
        PassDev d;
        memset(&d, 0, sizeof(d));
        const int R = Rs[i];
        d.R = R;
        d.S = (int)plan->S[i];
        d.ncols = (int)(n / R);
        d.contiguous = (i == P - 1);
        int C = 1;
        while (C * 2 * R <= tmax && C * 2 <= 64) C *= 2;
        while (C > 1 && C / 2 >= d.ncols) C /= 2;
        d.C = C;
        d.log2C = 0;
        while ((1 << d.log2C) < C) ++d.log2C;
        std::vector<int> rad = stage_radices(R);
        if ((int)rad.size() > kMaxStages) return nullptr;
        d.nstages = (int)rad.size();
        int L = R;
        for (int s = 0; s < d.nstages; ++s) {
            d.radix[s] = rad[s];
            d.divM[s].init(L / rad[s]);
            d.divNbf[s].init(R / rad[s]);
            L /= rad[s];
        }
        d.divR.init(R);
        d.divS.init(d.S);
        d.twR = (const float2 *)(base + o_twR[i]);
        d.perm = (const uint16_t *)(base + o_perm[i]);
        d.tw_lo = (const float2 *)(base + o_lo[i]);
        d.tw_hi = (const float2 *)(base + o_hi[i]);
        d.smem_bytes = (int)((size_t)C * R * sizeof(float2) + (size_t)R * sizeof(float2) +
                             (size_t)((R + 1) & ~1) * sizeof(uint16_t) + (size_t)C * sizeof(int));
        d.ntiles = (d.ncols + C - 1) / C;

        // digit-reversal: smem position of output k after the in-place DIF stages
        std::vector<uint16_t> perm(R);
        for (int k = 0; k < R; ++k) {
            int pos = 0, kk = k, span = R;
            for (int r : rad) {
                int dg = kk % r;
                kk /= r;
                span /= r;
                pos += dg * span;
            }
            perm[k] = (uint16_t)pos;
        }
        CUDA_CHECK(cudaMemcpyAsync(base + o_perm[i], perm.data(), R * sizeof(uint16_t),
                                   cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));   // perm is a stack temporary

        const double Lp = (double)R * (double)plan->S[i];
        twiddle_table_kernel<<<(R + 255) / 256, 256, 0, stream>>>((float2 *)(base + o_twR[i]), R, (double)R, 1.0);
        twiddle_table_kernel<<<((1 << kTwLoBits) + 255) / 256, 256, 0, stream>>>(
            (float2 *)(base + o_lo[i]), 1 << kTwLoBits, Lp, 1.0);
        twiddle_table_kernel<<<(unsigned)((n_hi[i] + 255) / 256), 256, 0, stream>>>(
            (float2 *)(base + o_hi[i]), (int)n_hi[i], Lp, (double)(1 << kTwLoBits));
        CUDA_CHECK(cudaGetLastError());

        d.tw_mode = d.contiguous ? 0 : 1;
        snprintf(d.tag, sizeof(d.tag), "fft_fwd_%d", i);
        plan->fwd[i] = d;
        d.tw_mode = d.contiguous ? 0 : 2;
        snprintf(d.tag, sizeof(d.tag), "fft_inv_%d", i);
        plan->inv[i] = d;
    }
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return plan;
}

}  // namespace wefax
