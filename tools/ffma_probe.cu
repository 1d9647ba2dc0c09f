// FFMA issue-rate probe: 192-thread CTAs, one per SM, register-only FMA chains (not product code)
#include <cuda_runtime.h>
#include <cstdio>
template <int CH, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(float *out, int iters, float x) {
    float w[64], acc[CH];
    for (int i = 0; i < 64; ++i) w[i] = x * (i + threadIdx.x);
    for (int c = 0; c < CH; ++c) acc[c] = 0.f;
    float h = x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i % CH] = fmaf(w[i], h, acc[i % CH]);
        h += 1e-9f;
    }
    float s = 0; for (int c = 0; c < CH; ++c) s += acc[c];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}
template <int CH, int THREADS> void run(float *out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    k<CH, THREADS><<<148, THREADS>>>(out, 100, 1.0f);
    cudaEventRecord(a); k<CH, THREADS><<<148, THREADS>>>(out, iters, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * 1.965e9;
    double warp_ffma_per_smsp = (double)iters * 64 * (THREADS / 32) / 4.0;
    printf("chains=%d threads=%d: %.3f ms, %.2f cycles per warp-FFMA per SMSP (avg warps/SMSP %.1f)\n", CH, THREADS, ms, cyc / warp_ffma_per_smsp, THREADS / 128.0);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    run<1, 192>(out); run<2, 192>(out); run<4, 192>(out); run<8, 192>(out);
    run<4, 128>(out); run<8, 128>(out); run<4, 256>(out); run<8, 256>(out); run<8, 512>(out); run<4, 1024>(out);
    return 0;
}
