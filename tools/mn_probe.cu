// Probe: which shared-memory layout does tcgen05.mma kind::tf32 expect for MN-MAJOR operands?  (not product code)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mn_probe tools/mn_probe.cu
// One CTA computes D[128 x 64] = A[128 x 32] * B[64 x 32]^T with both operands staged MN-major under a hypothesis
// (descriptor layout type, swizzle function, SBO) and the host compares with the exact integer result.
#include "../ma_league_b200/csrc/tc_gemm.cuh"
#include <vector>
#include <cstdlib>
void mal_set_error(const char *, ...) {}

struct Hyp { int layout_type, sbo, lbo, pass; };

// A is MN-major under test, B is K-major SWIZZLE_128B (known-good) and one-hot in k: D[m][n] = the value the tensor
// core fetched as A(m, k = n), n < 8 (one K = 8 MMA).  A's shared memory holds CODES of its own float index, so D decodes
// the address map the hardware applies for this (layout type, LBO, SBO).
__global__ void __launch_bounds__(128, 1) k_probe(float *D, Hyp h) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int NA = 32768;                                   // floats of the coded A region (128 KB)
    float *As = reinterpret_cast<float *>(sm);
    uint8_t *Bs = sm + NA * 4;                              // [64 rows x 32 floats] K-major SW128 slab
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < NA; i += 128) As[i] = (float)(h.pass == 0 ? (i & 1023) : (i >> 10)) ;
    for (int i = tid; i < 64 * 32; i += 128) { const int n = i >> 5, k = i & 31; *reinterpret_cast<float *>(Bs + sw128_off(n, k)) = (n == k && n < 8) ? 1.0f : 0.0f; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        uint64_t d = 0;
        d |= (uint64_t)((smem_u32(As) >> 4) & 0x3FFF);
        d |= (uint64_t)((h.lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((h.sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)h.layout_type << 61;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        umma_tf32(tmem_base, d, umma_desc_sw128(smem_u32(Bs)), idesc, 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t d1[32];
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    tmem_ld32_nowait(tl, d1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) D[tid * 8 + c] = __uint_as_float(d1[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u));
}

int main(int argc, char **argv) {
    Hyp h{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), 0};
    std::vector<float> D0(128 * 8), D1(128 * 8);
    float *dD;
    cudaMalloc(&dD, D0.size() * 4);
    const int smem = 32768 * 4 + 64 * 128;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pass = 0; pass < 2; ++pass) {
        h.pass = pass;
        k_probe<<<1, 128, smem>>>(dD, h);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("type %d sbo %d lbo %d: %s\n", h.layout_type, h.sbo, h.lbo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy((pass ? D1 : D0).data(), dD, D0.size() * 4, cudaMemcpyDeviceToHost);
    }
    printf("type %d sbo %d lbo %d: byte offset the tensor core read for A(m, k)\n", h.layout_type, h.sbo, h.lbo);
    const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 12, 16, 28, 31, 32, 33, 36, 64, 96, 127};
    for (int m : ms) {
        printf("  m %3d:", m);
        for (int k = 0; k < 8; ++k) printf(" %6d", 4 * ((int)D0[m * 8 + k] + 1024 * (int)D1[m * 8 + k]));
        printf("\n");
    }
    return 0;
}
