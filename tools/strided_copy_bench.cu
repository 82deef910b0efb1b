// Micro-benchmark: how fast can B200 move data with the access pattern of a strided FFT pass
// (tiles of R rows x C adjacent complex columns, row pitch S complex, in place) when the SMs do no
// arithmetic at all?  Separates "memory pattern" from "instruction" limits of fft_fast_strided_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_copy_bench strided_copy_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int R1, int R2, int C, int MODE>   // MODE 0 read+write, 1 read only, 2 write only
__global__ void __launch_bounds__(R2 *C <= 256 ? 256 : 512)
tile_copy(float2 *z, int S, int tiles_per_o, int total_tiles) {
    const int tid = threadIdx.x, cc = tid % C, row = tid / C;
    if (row >= R2) return;
    float2 acc = make_float2(0.f, 0.f);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int o = tile / tiles_per_o;
        const int m = (tile - o * tiles_per_o) * C + cc;
        if (m >= S) continue;
        float2 *g = z + (size_t)o * (R1 * R2) * (size_t)S + m + (size_t)row * S;
        float2 v[R1];
#pragma unroll
        for (int t = 0; t < R1; ++t) v[t] = MODE == 2 ? make_float2((float)t, (float)tile) : __ldg(g + (size_t)(R2 * t) * S);
        if (MODE == 1) {
#pragma unroll
            for (int t = 0; t < R1; ++t) { acc.x += v[t].x; acc.y += v[t].y; }
        } else {
            // rows written by this thread: row*R1 .. (a different set than it read, as in the FFT pass)
            float2 *w = z + (size_t)o * (R1 * R2) * (size_t)S + m;
#pragma unroll
            for (int t = 0; t < R1; ++t) w[(size_t)(row + R2 * t) * S] = make_float2(v[t].y, v[t].x);
        }
    }
    if (MODE == 1 && acc.x == 123.456f) z[0] = acc;
}

__global__ void flat_copy(float4 *z, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = z[i];
        z[i] = make_float4(v.y, v.x, v.w, v.z);
    }
}

template <class F> float time_it(F f, int reps = 20) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

template <int C, int MODE> void run(float2 *z, int S, long long n, const char *label, int ctas_per_sm) {
    constexpr int R1 = 15, R2 = 15;
    const int tiles_per_o = (S + C - 1) / C;
    const int nouter = (int)(n / ((long long)R1 * R2 * S));
    const int total = nouter * tiles_per_o;
    const int threads = R2 * C <= 256 ? 256 : 512;
    const int grid = std::min(total, 148 * ctas_per_sm);
    float ms = time_it([&] { tile_copy<R1, R2, C, MODE><<<grid, threads>>>(z, S, tiles_per_o, total); });
    const double bytes = (MODE == 0 ? 16.0 : 8.0) * (double)n;
    printf("%-34s S=%-6d C=%-2d ctas/SM=%d  %.1f us  %.0f GB/s\n", label, S, C, ctas_per_sm, ms * 1e3, bytes / ms / 1e6);
}

int main() {
    const long long n = 19845000;   // complex points (60-min recording, half length)
    float2 *z;
    CK(cudaMalloc(&z, n * sizeof(float2)));
    CK(cudaMemset(z, 0, n * sizeof(float2)));
    float ms = time_it([&] { flat_copy<<<148 * 8, 256>>>((float4 *)z, n / 2); });
    printf("%-34s %.1f us  %.0f GB/s\n", "flat in-place float4 copy", ms * 1e3, 16.0 * n / ms / 1e6);
    for (int S : {392, 88200}) {
        for (int cps : {3, 6}) {
            run<16, 0>(z, S, n, "tile read+write (in place)", cps);
            run<32, 0>(z, S, n, "tile read+write (in place)", cps);
            run<8, 0>(z, S, n, "tile read+write (in place)", cps);
        }
        run<16, 1>(z, S, n, "tile read only", 6);
        run<32, 1>(z, S, n, "tile read only", 6);
        run<16, 2>(z, S, n, "tile write only", 6);
        run<32, 2>(z, S, n, "tile write only", 6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
