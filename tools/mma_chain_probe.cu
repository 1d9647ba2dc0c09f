// Probe: cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N and of the number of independent TMEM
// accumulators the instructions rotate over (1 = every MMA accumulates into the previous one's result).  (not product code)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_chain_probe tools/mma_chain_probe.cu
#include "../ma_league_b200/csrc/tc_gemm.cuh"
#include <cstdlib>
void mal_set_error(const char *, ...) {}

template <int N, int C, int NM>
__global__ void __launch_bounds__(128, 1) k_chain(long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float *>(sm)[i] = 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dA = umma_desc_sw128(smem_u32(sm)), dB = umma_desc_sw128(smem_u32(sm + 16384));
        constexpr int cols = (512 / C) & ~31;
        static_assert(cols >= N || C == 1 || true, "");
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NM; ++i)
            umma_tf32(tmem_base + (uint32_t)((i % C) * cols), dA + (uint64_t)(((i & 3) * 32) >> 4), dB + (uint64_t)(((i & 3) * 32) >> 4), idesc, i >= C ? 1u : 0u);
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

template <int N, int C, int NM> void run(long long *d) {
    long long h[2];
    cudaFuncSetAttribute(k_chain<N, C, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_chain<N, C, NM><<<1, 128, 64 * 1024>>>(d);
    k_chain<N, C, NM><<<1, 128, 64 * 1024>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("N %d C %d: error %s\n", N, C, cudaGetErrorString(cudaGetLastError())); exit(1); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N %3d  accumulators %d  mmas %3d: issue %6lld cycles, issue+complete %6lld cycles = %6.1f per MMA\n", N, C, NM, h[0], h[1], (double)h[1] / NM);
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    run<64, 1, 12>(d); run<64, 1, 48>(d); run<64, 2, 48>(d); run<64, 4, 48>(d); run<64, 8, 48>(d);
    run<128, 1, 48>(d); run<128, 2, 48>(d); run<128, 4, 48>(d);
    run<192, 1, 48>(d); run<192, 2, 48>(d);
    run<256, 1, 48>(d); run<256, 2, 48>(d);
    return 0;
}
