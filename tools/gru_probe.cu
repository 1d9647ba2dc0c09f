// Standalone probe (not product code): which ingredient of the GRU recurrence step costs what on B200.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cmath>
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
// FLAGS: 1 = global stores, 2 = transcendentals, 4 = TMA ring + mbarrier, 8 = second barrier, 16 = matvec
template <int RT, int FLAGS>
__global__ void __launch_bounds__(192, 1) probe(const float *W, const float *gi, float *hout, float *gates, int TT, int R) {
    constexpr int DEPTH = 8;
    __shared__ __align__(128) float gi_s[DEPTH][RT * G3];
    __shared__ __align__(16) float h_s[RT * HID];
    __shared__ float rz_s[RT * 128];
    __shared__ float ghn_s[RT * HID];
    __shared__ __align__(8) uint64_t bars[DEPTH];
    const int tid = threadIdx.x, g = tid >> 6, r0 = blockIdx.x * RT;
    for (int i = tid; i < DEPTH * RT * G3; i += 192) (&gi_s[0][0])[i] = 0.01f * (i % 7);
    for (int i = tid; i < RT * HID; i += 192) h_s[i] = 0.f;
    if (tid == 0) { for (int s = 0; s < DEPTH; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint32_t bytes = RT * G3 * 4;
    if ((FLAGS & 4) && tid == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int t = 0; t < DEPTH && t < TT; ++t) { mbar_expect_tx(&bars[t], bytes); bulk_g2s(gi_s[t], gi + ((int64_t)t * R + r0) * G3, bytes, &bars[t]); }
    }
    float w[HID];
    for (int k = 0; k < HID; ++k) w[k] = W[tid * HID + k];
    float wt[8][8];   // 8x8 register tile: rows 8*(tid>>3)+jj, k-slice 8*(tid&7)+kk
    for (int jj = 0; jj < 8; ++jj) for (int kk = 0; kk < 8; ++kk) wt[jj][kk] = W[(8 * (tid >> 3) + jj) * HID + 8 * (tid & 7) + kk];
    const int lane = tid & 31;
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    for (int t = 0; t < TT; ++t) {
        const int slot = t % DEPTH;
        float acc[RT];
        if (FLAGS & 64) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const float4 h0 = reinterpret_cast<const float4 *>(h_s + r * HID + 8 * (tid & 7))[0];
                const float4 h1 = reinterpret_cast<const float4 *>(h_s + r * HID + 8 * (tid & 7))[1];
                float p[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    float a = wt[jj][0] * h0.x;
                    a = fmaf(wt[jj][1], h0.y, a); a = fmaf(wt[jj][2], h0.z, a); a = fmaf(wt[jj][3], h0.w, a);
                    a = fmaf(wt[jj][4], h1.x, a); a = fmaf(wt[jj][5], h1.y, a); a = fmaf(wt[jj][6], h1.z, a); a = fmaf(wt[jj][7], h1.w, a);
                    p[jj] = a;
                }
                float q4[4], q2[2];
#pragma unroll
                for (int u = 0; u < 4; ++u) { float send = b2 ? p[u] : p[u + 4]; float keep = b2 ? p[u + 4] : p[u]; q4[u] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
#pragma unroll
                for (int u = 0; u < 2; ++u) { float send = b1 ? q4[u] : q4[u + 2]; float keep = b1 ? q4[u + 2] : q4[u]; q2[u] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
                float send = b0 ? q2[0] : q2[1]; float keep = b0 ? q2[1] : q2[0];
                acc[r] = 0.1f + keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        } else if (FLAGS & 32) {
            float a0[RT], a1[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) { a0[r] = 0.1f; a1[r] = 0.f; }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float4 hv[RT];
#pragma unroll
                for (int r = 0; r < RT; ++r) hv[r] = reinterpret_cast<const float4 *>(h_s + r * HID)[k];
#pragma unroll
                for (int r = 0; r < RT; ++r) a0[r] = fmaf(w[4*k], hv[r].x, a0[r]);
#pragma unroll
                for (int r = 0; r < RT; ++r) a1[r] = fmaf(w[4*k+1], hv[r].y, a1[r]);
#pragma unroll
                for (int r = 0; r < RT; ++r) a0[r] = fmaf(w[4*k+2], hv[r].z, a0[r]);
#pragma unroll
                for (int r = 0; r < RT; ++r) a1[r] = fmaf(w[4*k+3], hv[r].w, a1[r]);
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) acc[r] = a0[r] + a1[r];
        } else
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float a0 = 0.1f, a1 = 0, a2 = 0, a3 = 0;
            if (FLAGS & 16) {
                float4 hv[16];
                const float4 *hp = reinterpret_cast<const float4 *>(h_s + r * HID);
#pragma unroll
                for (int k = 0; k < 16; ++k) hv[k] = hp[k];
#pragma unroll
                for (int k = 0; k < 16; ++k) { a0 = fmaf(w[4*k], hv[k].x, a0); a1 = fmaf(w[4*k+1], hv[k].y, a1); a2 = fmaf(w[4*k+2], hv[k].z, a2); a3 = fmaf(w[4*k+3], hv[k].w, a3); }
            } else a0 += h_s[r * HID + (tid & 63)] * w[0];
            acc[r] = (a0 + a1) + (a2 + a3);
        }
        if (FLAGS & 4) mbar_wait(&bars[slot], (uint32_t)((t / DEPTH) & 1));
        if (g < 2) {
#pragma unroll
            for (int r = 0; r < RT; ++r) { float x = gi_s[slot][r * G3 + tid] + acc[r]; rz_s[r * 128 + tid] = (FLAGS & 2) ? __fdividef(1.f, 1.f + expf(-x)) : x * 0.5f; }
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r) ghn_s[r * HID + tid - 128] = acc[r];
        }
        __syncthreads();
        for (int item = tid; item < RT * HID; item += 192) {
            const int r = item >> 6, i = item & 63, row = r0 + r;
            float rr = rz_s[r * 128 + i], zz = rz_s[r * 128 + 64 + i], ghn = ghn_s[item];
            float x = gi_s[slot][r * G3 + 128 + i] + rr * ghn;
            float nn = (FLAGS & 2) ? tanhf(x) : x * 0.25f;
            float hp = h_s[item], hn = nn + zz * (hp - nn);
            h_s[item] = hn;
            if ((FLAGS & 1) && row < R) {
                int64_t m = (int64_t)t * R + row;
                hout[m * HID + i] = hn;
                float *gp = gates + m * 256;
                gp[i] = rr; gp[64 + i] = zz; gp[128 + i] = nn; gp[192 + i] = ghn;
            }
        }
        if (FLAGS & 8) __syncthreads(); else __syncwarp();
        if ((FLAGS & 4) && tid == 0 && t + DEPTH < TT) { mbar_expect_tx(&bars[slot], bytes); bulk_g2s(gi_s[slot], gi + ((int64_t)(t + DEPTH) * R + r0) * G3, bytes, &bars[slot]); }
    }
    if (!(FLAGS & 1) && tid < RT * HID) hout[(int64_t)blockIdx.x * 192 + tid] = h_s[tid];
}
template <int RT, int FLAGS>
void run(const char *name, const float *W, const float *gi, float *hout, float *gates, int TT, int R) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = (R + RT - 1) / RT;
    for (int i = 0; i < 3; ++i) probe<RT, FLAGS><<<grid, 192>>>(W, gi, hout, gates, TT, R);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) probe<RT, FLAGS><<<grid, 192>>>(W, gi, hout, gates, TT, R);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-44s RT=%d grid=%3d  %.1f us/launch  %.0f cycles/step @1.965GHz  (%s)\n", name, RT, grid, ms * 100, ms * 100 / TT * 1965, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int TT = 201, R = 320;
    float *W, *gi, *hout, *gates;
    cudaMalloc(&W, 192 * 64 * 4); cudaMalloc(&gi, (size_t)TT * R * G3 * 4); cudaMalloc(&hout, (size_t)TT * R * 64 * 4); cudaMalloc(&gates, (size_t)TT * R * 256 * 4);
    cudaMemset(W, 0, 192 * 64 * 4); cudaMemset(gi, 0, (size_t)TT * R * G3 * 4);
    run<3, 31>("full (stores+transc+tma+bar2+matvec)", W, gi, hout, gates, TT, R);
    run<3, 95>("full, 8x8 tile + shuffle fold", W, gi, hout, gates, TT, R);
    run<3, 80>("tile matvec + barrier only", W, gi, hout, gates, TT, R);
    run<2, 95>("full, 8x8 tile", W, gi, hout, gates, TT, R);
    run<4, 95>("full, 8x8 tile", W, gi, hout, gates, TT, R);
    run<6, 95>("full, 8x8 tile", W, gi, hout, gates, TT, R);
    run<3, 91>("8x8 tile, no TMA/mbarrier", W, gi, hout, gates, TT, R);
    run<3, 93>("8x8 tile, no transcendentals", W, gi, hout, gates, TT, R);
    run<3, 94>("8x8 tile, no stores", W, gi, hout, gates, TT, R);
    run<3, 48>("matvec(k-outer) + barrier only", W, gi, hout, gates, TT, R);
    run<4, 63>("full, k-outer", W, gi, hout, gates, TT, R);
    run<6, 63>("full, k-outer", W, gi, hout, gates, TT, R);
    run<2, 63>("full, k-outer", W, gi, hout, gates, TT, R);
    run<3, 30>("no global stores", W, gi, hout, gates, TT, R);
    run<3, 29>("no transcendentals", W, gi, hout, gates, TT, R);
    run<3, 27>("no TMA/mbarrier", W, gi, hout, gates, TT, R);
    run<3, 23>("second barrier -> syncwarp (invalid, timing only)", W, gi, hout, gates, TT, R);
    run<3, 15>("no matvec", W, gi, hout, gates, TT, R);
    run<3, 0>("nothing but 1 barrier + smem traffic", W, gi, hout, gates, TT, R);
    run<3, 16>("matvec + barrier only", W, gi, hout, gates, TT, R);
    run<2, 31>("full", W, gi, hout, gates, TT, R);
    run<1, 31>("full", W, gi, hout, gates, TT, R);
    run<4, 31>("full", W, gi, hout, gates, TT, R);
    return 0;
}
