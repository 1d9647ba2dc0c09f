// FFMA vs FFMA2 (fma.rn.f32x2) issue-rate probe (not product code): register-only chains, one CTA per SM.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
template <int CH, int THREADS, int PACKED>
__global__ void __launch_bounds__(THREADS, 1) k(float *out, int iters, float x) {
    unsigned long long w2[32], acc2[CH];
    float w[64], acc[CH];
    for (int i = 0; i < 64; ++i) w[i] = x * (i + threadIdx.x);
    for (int i = 0; i < 32; ++i) w2[i] = pack2(w[2 * i], w[2 * i + 1]);
    for (int c = 0; c < CH; ++c) { acc[c] = 0.f; acc2[c] = pack2(0.f, 0.f); }
    float h = x;
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
            const unsigned long long h2 = pack2(h, h + 1.0f);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc2[i % CH] = fma2(w2[i], h2, acc2[i % CH]);
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i % CH] = fmaf(w[i], h, acc[i % CH]);
        }
        h += 1e-9f;
    }
    float s = 0;
    for (int c = 0; c < CH; ++c) s += acc[c] + (float)(acc2[c] & 0xffff);
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}
template <int CH, int THREADS, int PACKED> void run(float *out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    k<CH, THREADS, PACKED><<<148, THREADS>>>(out, 100, 1.0f);
    cudaEventRecord(a); k<CH, THREADS, PACKED><<<148, THREADS>>>(out, iters, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * 1.965e9;
    double fma_lane_ops_per_smsp = (double)iters * 64 * (THREADS / 32) / 4.0;     // scalar-FMA warp-instruction equivalents
    printf("%s chains=%d threads=%d: %.3f ms, %.2f cycles per 32 FMAs per SMSP (%.2f per instruction)\n", PACKED ? "FFMA2" : "FFMA ", CH, THREADS, ms,
           cyc / fma_lane_ops_per_smsp, cyc / fma_lane_ops_per_smsp * (PACKED ? 2 : 1));
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    run<4, 128, 0>(out); run<4, 128, 1>(out); run<8, 128, 1>(out); run<3, 128, 1>(out); run<2, 128, 1>(out);
    run<4, 256, 0>(out); run<4, 256, 1>(out); run<4, 512, 1>(out); run<8, 512, 0>(out);
    return 0;
}
