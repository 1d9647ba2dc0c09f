// Probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (not product code).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ta_probe tools/ta_probe.cu
// One CTA writes A[128 x 64] into TMEM with tcgen05.st.32x32b (lane = row m, column = k, value coded from m and k), stages a
// K-major SWIZZLE_128B B[64 x 64] = identity in shared memory and issues 8 K = 8 MMAs (A column offset 8 ks, B
// descriptor offset as in tc_gemm.cuh).  Hypothesis "lane = row, one 32-bit column per k": D[m][n] = A(m, n).
#include "../ma_league_b200/csrc/tc_gemm.cuh"
#include <vector>
#include <cstdlib>
void mal_set_error(const char *, ...) {}

__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) k_probe(float *D, int pass) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *Bs = sm;                                       // 2 slabs of [64 rows x 32 floats]
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<float *>(Bs + (k >> 5) * (64 * 128) + sw128_off(n, k & 31)) = (n == k) ? 1.0f : 0.0f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    // A at columns 64..127: lane = row, column = k
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(pass == 0 ? (float)((tid & 15) * 64 + c0 + e) : (float)tid);
        tmem_st8(tl + 64 + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dB = umma_desc_sw128(smem_u32(Bs));
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t wo = (uint64_t)(((ks >> 2) * (64 * 128) + (ks & 3) * 32) >> 4);
            umma_tf32_ta(tmem_base, tmem_base + 64 + ks * 8, dB + wo, idesc, ks == 0 ? 0u : 1u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t d1[32], d2[32];
    tmem_ld32_nowait(tl, d1);
    tmem_ld32_nowait(tl + 32, d2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) { D[tid * 64 + c] = __uint_as_float(d1[c]); D[tid * 64 + 32 + c] = __uint_as_float(d2[c]); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

int main() {
    float *dD;
    cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemset(dD, 0, 128 * 64 * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int pass = 0; pass < 2; ++pass) {
        k_probe<<<1, 128, 32768>>>(dD, pass);
        cudaError_t e = cudaDeviceSynchronize();
        printf("pass %d sync: %s\n", pass, cudaGetErrorString(e));
        std::vector<float> D(128 * 64);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n)
                if (D[m * 64 + n] != (pass == 0 ? (float)((m & 15) * 64 + n) : (float)m)) ++bad;
        printf("pass %d mismatches against 'lane = row, column = k': %d of %d\n", pass, bad, 128 * 64);
        for (int m : {0, 1, 5, 33, 127}) {
            printf("m=%3d:", m);
            for (int n : {0, 1, 7, 8, 9, 31, 32, 63}) printf(" D[%d]=%g", n, D[m * 64 + n]);
            printf("\n");
        }
    }
    return 0;
}
