// Probe 2 (not product code): 384-thread recurrence, K split over thread pairs, one-warp mbarrier wait.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define HID 64
#define G3 192
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *d, const void *s, uint32_t b, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(b), "r"(smem_u32(bar)) : "memory");
}
// thread tid: gate row j = tid >> 1, k-half kh = tid & 1 (32 weights in registers)
template <int RT, int WAITMODE>   // WAITMODE 0: all threads wait on the mbarrier; 1: warp 11 waits one step ahead
__global__ void __launch_bounds__(384, 1) probe(const float *W, const float *gi, float *hout, float *gates, int TT, int R) {
    constexpr int DEPTH = 8;
    __shared__ __align__(128) float gi_s[DEPTH][RT * G3];
    __shared__ __align__(16) float h_s[RT * HID];
    __shared__ float rz_s[RT * 128];
    __shared__ float ghn_s[RT * HID];
    __shared__ __align__(8) uint64_t bars[DEPTH];
    const int tid = threadIdx.x, j = tid >> 1, kh = tid & 1, g = j >> 6, r0 = blockIdx.x * RT, warp = tid >> 5;
    for (int i = tid; i < DEPTH * RT * G3; i += 384) (&gi_s[0][0])[i] = 0.01f * (i % 7);
    for (int i = tid; i < RT * HID; i += 384) h_s[i] = 0.f;
    if (tid == 0) { for (int s = 0; s < DEPTH; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint32_t bytes = RT * G3 * 4;
    if (tid == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int t = 0; t < DEPTH && t < TT; ++t) { mbar_expect_tx(&bars[t], bytes); bulk_g2s(gi_s[t], gi + ((int64_t)t * R + r0) * G3, bytes, &bars[t]); }
    }
    float w[32];
    for (int k = 0; k < 32; ++k) w[k] = W[j * HID + 32 * kh + k];
    if (WAITMODE == 1) mbar_wait(&bars[0], 0);
    __syncthreads();
    for (int t = 0; t < TT; ++t) {
        const int slot = t % DEPTH;
        float acc[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float4 hv[8];
            const float4 *hp = reinterpret_cast<const float4 *>(h_s + r * HID + 32 * kh);
#pragma unroll
            for (int k = 0; k < 8; ++k) hv[k] = hp[k];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { a0 = fmaf(w[4*k], hv[k].x, a0); a1 = fmaf(w[4*k+1], hv[k].y, a1); a2 = fmaf(w[4*k+2], hv[k].z, a2); a3 = fmaf(w[4*k+3], hv[k].w, a3); }
            float s = (a0 + a1) + (a2 + a3);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            acc[r] = s + 0.1f;
        }
        if (WAITMODE == 0) mbar_wait(&bars[slot], (uint32_t)((t / DEPTH) & 1));
        // each thread pair finishes one gate value per row: kh picks the row parity to spread the transcendental work
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if ((r & 1) == kh) {
                if (g < 2) { float x = gi_s[slot][r * G3 + j] + acc[r]; rz_s[r * 128 + j] = __fdividef(1.f, 1.f + expf(-x)); }
                else ghn_s[r * HID + j - 128] = acc[r];
            }
        }
        __syncthreads();
        for (int item = tid; item < RT * HID; item += 384) {
            const int r = item >> 6, i = item & 63, row = r0 + r;
            float rr = rz_s[r * 128 + i], zz = rz_s[r * 128 + 64 + i], ghn = ghn_s[item];
            float nn = tanhf(gi_s[slot][r * G3 + 128 + i] + rr * ghn);
            float hp = h_s[item], hn = nn + zz * (hp - nn);
            h_s[item] = hn;
            if (row < R) {
                int64_t m = (int64_t)t * R + row;
                hout[m * HID + i] = hn;
                float *gp = gates + m * 256;
                gp[i] = rr; gp[64 + i] = zz; gp[128 + i] = nn; gp[192 + i] = ghn;
            }
        }
        if (WAITMODE == 1 && warp == 11 && t + 1 < TT) mbar_wait(&bars[(t + 1) % DEPTH], (uint32_t)(((t + 1) / DEPTH) & 1));
        __syncthreads();
        if (tid == 0 && t + DEPTH < TT) { mbar_expect_tx(&bars[slot], bytes); bulk_g2s(gi_s[slot], gi + ((int64_t)(t + DEPTH) * R + r0) * G3, bytes, &bars[slot]); }
    }
}
template <int RT, int WM>
void run(const float *W, const float *gi, float *hout, float *gates, int TT, int R) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = (R + RT - 1) / RT;
    for (int i = 0; i < 3; ++i) probe<RT, WM><<<grid, 384>>>(W, gi, hout, gates, TT, R);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) probe<RT, WM><<<grid, 384>>>(W, gi, hout, gates, TT, R);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("384 threads K-split RT=%d waitmode=%d grid=%3d  %.1f us/launch  %.0f cycles/step  (%s)\n", RT, WM, grid, ms * 100, ms * 100 / TT * 1965, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int TT = 201, R = 320;
    float *W, *gi, *hout, *gates;
    cudaMalloc(&W, 192 * 64 * 4); cudaMalloc(&gi, (size_t)TT * R * G3 * 4); cudaMalloc(&hout, (size_t)TT * R * 64 * 4); cudaMalloc(&gates, (size_t)TT * R * 256 * 4);
    cudaMemset(W, 0, 192 * 64 * 4); cudaMemset(gi, 0, (size_t)TT * R * G3 * 4);
    run<3, 0>(W, gi, hout, gates, TT, R); run<3, 1>(W, gi, hout, gates, TT, R);
    run<2, 0>(W, gi, hout, gates, TT, R); run<2, 1>(W, gi, hout, gates, TT, R);
    run<4, 1>(W, gi, hout, gates, TT, R);
    return 0;
}
