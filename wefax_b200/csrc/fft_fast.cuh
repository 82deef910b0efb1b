// Specialised two-stage strided FFT pass: compile-time pass length R = R1 * R2.
//
// The generic pass kernel (fft.cuh) is instruction-bound: ~120 thread instructions per
// point per pass (run-time radices, strides and divisions; three shared-memory round
// trips; a separate read-out loop).  For the strided passes of the long transforms on the
// decode path this kernel does the same arithmetic with everything but the global stride
// known at compile time:
//
//   stage 1  thread (q, cc): R1 rows  q + R2*t  of column cc straight from global memory
//            into registers (two 128-byte row segments per warp), length-R1 DFT in
//            registers (prime-factor / constant-twiddle codelets, no table look-ups),
//            stage twiddles w_R^(q*u) from a [q][u] table at immediate offsets, one
//            shared-memory write;
//   stage 2  thread (u, cc): the R2 contiguous rows of block u from shared memory, length-R2
//            DFT in registers, inter-pass twiddle, and the store functor straight from
//            registers (row k = u + R1*k2, again 128-byte segments per row).
//
// One shared-memory round trip, one __syncthreads per tile (double-buffered tile), persistent
// CTAs striding over the tiles.  The inter-pass twiddle w^(e0 + k*de) of row k = u + R1*k2
// is A[u] * P[k2] with A = w^(e0 + u*de) (one exact two-level look-up per thread) and
// P[k2] = w^(k2*R1*de) (one look-up per thread, shared through the tile buffer).
//
// Output convention, tables and twiddle exponents are those of fft_pass_kernel with stage
// radices {R1, R2}; a plan may mix both kernels freely.
#pragma once

#include "fft.cuh"

namespace wefax {
namespace fast {

// ----------------------------- codelets -------------------------------------
template <int r> struct Cs {
    __device__ __forceinline__ static void run(float2 *v) { Bfly<r>::run(v, nullptr); }   // 2, 4, 8
};
template <> struct Cs<3> {
    __device__ __forceinline__ static void run(float2 *v) {
        const float2 cs[1] = {make_float2(-0.5f, 0.86602540378443865f)};
        Bfly<3>::run(v, cs);
    }
};
template <> struct Cs<5> {
    __device__ __forceinline__ static void run(float2 *v) {
        const float2 cs[2] = {make_float2(0.30901699437494742f, 0.95105651629515357f),
                              make_float2(-0.80901699437494742f, 0.58778525229247313f)};
        Bfly<5>::run(v, cs);
    }
};
template <> struct Cs<7> {
    __device__ __forceinline__ static void run(float2 *v) {
        const float2 cs[3] = {make_float2(0.62348980185873353f, 0.78183148246802981f),
                              make_float2(-0.22252093395631440f, 0.97492791218182361f),
                              make_float2(-0.90096886790241913f, 0.43388373911755812f)};
        Bfly<7>::run(v, cs);
    }
};

__host__ __device__ constexpr int mod_inverse(int a, int m) {
    for (int i = 1; i < m; ++i)
        if ((a * i) % m == 1) return i;
    return 1;
}

// Good-Thomas prime-factor DFT of length A*B (gcd(A, B) = 1): no inner twiddles, index
// maps resolved at compile time.  Natural order in, natural order out.
template <int A, int B> struct Pfa {
    static constexpr int N = A * B;
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t[A][B];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1)
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) t[n1][n2] = v[(B * n1 + A * n2) % N];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) {
            float2 x[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) x[n1] = t[n1][n2];
            Cs<A>::run(x);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) t[k1][n2] = x[k1];
        }
        static_assert(A > 1 && B > 1, "coprime factors");
        constexpr int ia = B * mod_inverse(B % A, A), ib = A * mod_inverse(A % B, B);
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            Cs<B>::run(t[k1]);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) v[(ia * k1 + ib * k2) % N] = t[k1][k2];
        }
    }
};

template <int N> struct Dft;
template <> struct Dft<12> {
    __device__ __forceinline__ static void run(float2 *v) { Pfa<3, 4>::run(v); }
};
template <> struct Dft<14> {
    __device__ __forceinline__ static void run(float2 *v) { Pfa<2, 7>::run(v); }
};
template <> struct Dft<15> {
    __device__ __forceinline__ static void run(float2 *v) { Pfa<3, 5>::run(v); }
};
// 16 = 4 x 4 Cooley-Tukey with literal twiddles w16^e = exp(-2*pi*i*e/16)
template <> struct Dft<16> {
    __device__ __forceinline__ static void run(float2 *v) {
        const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
        float2 t[4][4];   // t[ua][tb]
#pragma unroll
        for (int tb = 0; tb < 4; ++tb) {
            float2 x[4] = {v[tb], v[4 + tb], v[8 + tb], v[12 + tb]};
            Bfly<4>::run(x, nullptr);
#pragma unroll
            for (int ua = 0; ua < 4; ++ua) t[ua][tb] = x[ua];
        }
        // inner twiddles w16^(ua*tb)
        const float2 w1 = make_float2(c1, -s1), w2 = make_float2(h, -h), w3 = make_float2(s1, -c1);
        const float2 w6 = make_float2(-h, -h), w9 = make_float2(-c1, s1);
        t[1][1] = cmul(t[1][1], w1);
        t[1][2] = cmul(t[1][2], w2);
        t[1][3] = cmul(t[1][3], w3);
        t[2][1] = cmul(t[2][1], w2);
        t[2][2] = make_float2(t[2][2].y, -t[2][2].x);   // w16^4 = -i
        t[2][3] = cmul(t[2][3], w6);
        t[3][1] = cmul(t[3][1], w3);
        t[3][2] = cmul(t[3][2], w6);
        t[3][3] = cmul(t[3][3], w9);
#pragma unroll
        for (int ua = 0; ua < 4; ++ua) {
            Bfly<4>::run(t[ua], nullptr);
#pragma unroll
            for (int ub = 0; ub < 4; ++ub) v[ua + 4 * ub] = t[ua][ub];
        }
    }
};

// ----------------------------- the pass -------------------------------------
template <int R1, int R2, int C> struct Cfg {
    static constexpr int R = R1 * R2;
    static constexpr int ROWS = R1 > R2 ? R1 : R2;
    static constexpr int T = ((ROWS * C + 31) / 32) * 32;
    static constexpr int BUF = R * C + R2 * C;                      // tile + P table, in float2
    static constexpr int SMEM = (R + 2 * BUF) * (int)sizeof(float2);
    static constexpr int MINB = (3 * SMEM + 3 * 1024 <= 227 * 1024 && 3 * T <= 768) ? 3 : 2;
};

template <int R1, int R2, int C, class StoreOp>
__global__ void __launch_bounds__((Cfg<R1, R2, C>::T), (Cfg<R1, R2, C>::MINB))
fft_fast_strided_kernel(const PassDev p, const float2 *src, size_t src_bstride, const StoreOp st, int total_tiles) {
    using K = Cfg<R1, R2, C>;
    constexpr int R = K::R;
    extern __shared__ __align__(16) unsigned char fast_smem[];
    float2 *twQ = reinterpret_cast<float2 *>(fast_smem);   // [q][u] stage twiddles w_R^(q*u)
    float2 *buf0 = twQ + R;

    const int tid = threadIdx.x;
    const int cc = tid % C, row = tid / C;                 // row = q in stage 1, = u in stage 2
    for (int i = tid; i < R; i += K::T) {
        const int q = i / R1, u = i - q * R1;
        twQ[i] = __ldg(p.twR + q * u);
    }
    __syncthreads();
    const bool act1 = row < R2, act2 = row < R1;
    const size_t rstride = (size_t)p.S;
    constexpr bool kHasSide = !std::is_empty<typename StoreOp::Side>::value;

    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        float2 *tb = buf0 + (it & 1) * K::BUF;
        float2 *P = tb + R * C;
        const int batch = tile / p.fast_ntiles;
        const int t_in = tile - batch * p.fast_ntiles;
        const int o = p.fast_divTpo.div(t_in);
        const uint32_t m = (uint32_t)((t_in - o * p.fast_tiles_per_o) * C + cc);
        const bool colok = m < (uint32_t)p.S;
        const size_t cbase = (size_t)o * (size_t)R * rstride + m;

        // inter-pass twiddle factors of this thread (exact two-level look-ups, in flight with the data)
        uint32_t e0 = 0, de = 0;
        if (p.tw_mode == 1) {
            de = m;
        } else if (p.tw_mode == 2) {
            const uint32_t ko = (uint32_t)o % (uint32_t)p.ko_R;
            e0 = ko * m;
            de = ko * (uint32_t)p.S;
        }
        if (!colok) e0 = de = 0;
        float2 Pval = make_float2(1.f, 0.f), Aval = make_float2(1.f, 0.f);
        if (p.tw_mode != 0) {
            if (act1) Pval = pass_twiddle(p, (uint32_t)(row * R1) * de);
            if (act2) Aval = pass_twiddle(p, e0 + (uint32_t)row * de);
        }

        if (act1) {
            float2 v[R1];
            const float2 *g = src + (size_t)batch * src_bstride + cbase + (size_t)row * rstride;
#pragma unroll
            for (int t = 0; t < R1; ++t)
                v[t] = colok ? __ldg(g + (size_t)(R2 * t) * rstride) : make_float2(0.f, 0.f);
            if (p.tw_mode != 0) P[row * C + cc] = Pval;
            Dft<R1>::run(v);
            const float2 *tq = twQ + row * R1;
            tb[row * C + cc] = v[0];
#pragma unroll
            for (int u = 1; u < R1; ++u) tb[(row + R2 * u) * C + cc] = cmul(v[u], tq[u]);
        }
        __syncthreads();
        if (act2) {
            float2 y[R2];
            typename StoreOp::Side side[R2];
            if constexpr (kHasSide) {
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2)
                    if (colok) side[k2] = st.side_load(cbase + (size_t)(row + R1 * k2) * rstride, batch);
            }
#pragma unroll
            for (int t = 0; t < R2; ++t) y[t] = tb[(row * R2 + t) * C + cc];
            Dft<R2>::run(y);
            if (colok) {
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    const int k = row + R1 * k2;
                    float2 val = y[k2];
                    if (p.tw_mode != 0) val = cmul(val, k2 == 0 ? Aval : cmul(Aval, P[k2 * C + cc]));
                    st(cbase + (size_t)k * rstride, batch, val, k, 0, side[k2]);
                }
            }
        }
    }
}

// (R1, R2) pairs with a compiled kernel; 0 when R has none
inline bool fast_pair(int R, int *R1, int *R2) {
    static const int pairs[][2] = {{12, 12}, {14, 12}, {15, 12}, {16, 12}, {14, 14},
                                   {15, 14}, {16, 14}, {15, 15}, {16, 15}, {16, 16}};
    for (auto &pr : pairs)
        if (pr[0] * pr[1] == R) {
            if (R1) *R1 = pr[0];
            if (R2) *R2 = pr[1];
            return true;
        }
    return false;
}

constexpr int kFastC = 16;

}  // namespace fast
}  // namespace wefax
