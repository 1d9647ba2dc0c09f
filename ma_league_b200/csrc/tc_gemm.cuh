// Tensor-core panel GEMM for the batched (rows x 64) GRU / hypernet projections:
//     Y[m, n] = epi( sum_k A[m,k] * W[n,k] )         (same problem descriptors as k_linear_group)
//
// 5th-gen tensor cores (tcgen05.mma kind::tf32, SASS UTC*MMA) with the accumulator in TMEM.  Plain TF32 (10-bit
// mantissa) misses the 1e-5 parity bar, so every operand is split in two TF32 pieces, x = hi + lo, and three MMAs
// (hi*hi + lo*hi + hi*lo, fp32 accumulate) recover fp32-level accuracy ("3xTF32").
//
// One CTA = 128 threads owns 128-row M tiles (persistent loop).  Per (n-tile, 64-wide k-chunk):
//   all threads : gather the A rows through the fused loaders, split hi/lo, store into the canonical K-major
//                 SWIZZLE_128B shared-memory layout (8-row x 128-byte atoms, 16-byte chunks XOR-swizzled by row);
//                 same for the W chunk;  fence.proxy.async + barrier
//   thread 0    : 8 k-steps x 3 tcgen05.mma (M=128, N=n-tile, K=8) into TMEM, tcgen05.commit -> mbarrier;
//                 hi*hi goes to accumulator D1, the two small correction products to a separate accumulator D2
//                 (the tensor core adds into the accumulator with truncation; keeping the 2^-11-scaled terms apart
//                 and summing D1 + D2 in fp32 in the epilogue keeps the result at fp32-level accuracy)
//   all threads : wait, tcgen05.ld 32x32b (warp w <-> TMEM lanes 32w..32w+31 = tile rows), epilogue, vector stores
#pragma once
#include "mal_common.cuh"
#include "learner.cuh"

#define TC_M 128
#define TC_KC 64                  // k-chunk (floats) = two 128-byte swizzle slabs
#define TC_NMAX 192               // widest n-tile (TMEM columns, smem budget)
#define TC_SLAB_A (TC_M * 128)    // bytes of one [128 rows x 32 floats] slab
#define TC_THREADS 256
#define TC_WARPS (TC_THREADS / 32)
#define TC_D2_COL 256             // TMEM column of the correction-term accumulator

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// byte offset of element (row r, k in [0,32)) inside one K-major SWIZZLE_128B slab of `rows` rows
__device__ __forceinline__ uint32_t sw128_off(int r, int kk) {
    const uint32_t chunk = (uint32_t)(kk >> 2) ^ (uint32_t)(r & 7);
    return (uint32_t)r * 128u + chunk * 16u + (uint32_t)(kk & 3) * 4u;
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    // start address >> 4 | LBO (unused for swizzled K-major) = 1 | SBO = 1024 B >> 4 | version 1 | SWIZZLE_128B (2)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_linear_tc(const __grid_constant__ LinGroup g) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    // carve: A_hi | A_lo (2 slabs each) | W_hi | W_lo (2 slabs of TC_NMAX rows each); every slab 1024-byte aligned
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    uint8_t *A_hi = base, *A_lo = base + 2 * TC_SLAB_A;
    uint8_t *W_hi = base + 4 * TC_SLAB_A, *W_lo = W_hi + 2 * TC_NMAX * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LinProb &p = g.p[blockIdx.y];
    const int n_mtiles = (p.M + TC_M - 1) / TC_M;
    if ((int)blockIdx.x >= n_mtiles) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    uint32_t bar_phase = 0;

    const int nkc = (p.K + TC_KC - 1) / TC_KC;
    // n-tiles: equal widths, multiples of 16, at most TC_NMAX
    const int n_ntiles = (p.Nout + TC_NMAX - 1) / TC_NMAX;
    const int nt_w = (((p.Nout + n_ntiles - 1) / n_ntiles) + 15) & ~15;
    int w_cached = -1;    // (n-tile * nkc + kc) currently staged in W_hi / W_lo

    for (int mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
        const int64_t m0 = (int64_t)mt * TC_M;
        for (int nt = 0; nt < n_ntiles; ++nt) {
            const int n0 = nt * nt_w;
            const int nw = (p.Nout - n0) < nt_w ? (((p.Nout - n0) + 15) & ~15) : nt_w;   // MMA N (multiple of 16)
            for (int kc = 0; kc < nkc; ++kc) {
                const int k0 = kc * TC_KC;
                // ---- stage the A chunk [128 x 64]: warp w handles rows w, w+8, ...; lanes along k (coalesced);
                //      loads of 4 rows are issued back to back before any conversion (memory-level parallelism)
                if (nkc > 1 || nt == 0) {
#pragma unroll 1
                    for (int rb = warp; rb < TC_M; rb += 4 * TC_WARPS) {
                        float v0[4], v1[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = rb + u * TC_WARPS;
                            const int64_t m = m0 + r;
                            v0[u] = 0.0f; v1[u] = 0.0f;
                            if (m < p.M) {
                                RowSrc rs = resolve_row(p.a_kind, g.bv, p.A, p.lda, p.shift, m);
                                if (k0 + lane < p.K) v0[u] = row_elem(p.a_kind, g.bv, rs, k0 + lane);
                                if (k0 + 32 + lane < p.K) v1[u] = row_elem(p.a_kind, g.bv, rs, k0 + 32 + lane);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = rb + u * TC_WARPS;
                            const uint32_t off = sw128_off(r, lane);
                            const uint32_t h0 = tf32_rna(v0[u]), h1 = tf32_rna(v1[u]);
                            *reinterpret_cast<uint32_t *>(A_hi + off) = h0;
                            *reinterpret_cast<uint32_t *>(A_hi + TC_SLAB_A + off) = h1;
                            *reinterpret_cast<uint32_t *>(A_lo + off) = tf32_rna(v0[u] - __uint_as_float(h0));
                            *reinterpret_cast<uint32_t *>(A_lo + TC_SLAB_A + off) = tf32_rna(v1[u] - __uint_as_float(h1));
                        }
                    }
                }
                // ---- stage the W chunk [nw x 64] unless it is already resident
                const int w_id = nt * nkc + kc;
                if (w_id != w_cached) {
                    const uint32_t slab_w = (uint32_t)TC_NMAX * 128u;
                    if (!p.w_trans) {
#pragma unroll 1
                        for (int jb = warp; jb < nw; jb += 4 * TC_WARPS) {
                            float v0[4], v1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = jb + u * TC_WARPS, n = n0 + j;
                                v0[u] = 0.0f; v1[u] = 0.0f;
                                if (j < nw && n < p.Nout) {
                                    const float *wr = p.W + (int64_t)n * p.ldw;
                                    if (k0 + lane < p.K) v0[u] = __ldg(wr + k0 + lane);
                                    if (k0 + 32 + lane < p.K) v1[u] = __ldg(wr + k0 + 32 + lane);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = jb + u * TC_WARPS;
                                if (j < nw) {
                                    const uint32_t off = sw128_off(j, lane);
                                    const uint32_t h0 = tf32_rna(v0[u]), h1 = tf32_rna(v1[u]);
                                    *reinterpret_cast<uint32_t *>(W_hi + off) = h0;
                                    *reinterpret_cast<uint32_t *>(W_hi + slab_w + off) = h1;
                                    *reinterpret_cast<uint32_t *>(W_lo + off) = tf32_rna(v0[u] - __uint_as_float(h0));
                                    *reinterpret_cast<uint32_t *>(W_lo + slab_w + off) = tf32_rna(v1[u] - __uint_as_float(h1));
                                }
                            }
                        }
                    } else {   // W(n,k) = W[k*ldw + n]: lanes along n (coalesced), loop over k
                        for (int idx = tid; idx < nw * TC_KC; idx += TC_THREADS) {
                            const int kk = idx / nw, j = idx - kk * nw;
                            const int n = n0 + j, k = k0 + kk;
                            float v = 0.0f;
                            if (n < p.Nout && k < p.K) v = __ldg(p.W + (int64_t)k * p.ldw + n);
                            const uint32_t off = (uint32_t)(kk >> 5) * slab_w + sw128_off(j, kk & 31);
                            const uint32_t h = tf32_rna(v);
                            *reinterpret_cast<uint32_t *>(W_hi + off) = h;
                            *reinterpret_cast<uint32_t *>(W_lo + off) = tf32_rna(v - __uint_as_float(h));
                        }
                    }
                    w_cached = w_id;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // ---- MMAs: D[128 x nw] (+)= A_chunk . W_chunk^T, three TF32 products per k-step
                if (tid == 0) {
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nw >> 3) << 17) |
                                           ((uint32_t)(TC_M >> 4) << 24);
                    const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), w_hi = smem_u32(W_hi), w_lo = smem_u32(W_lo);
                    const int ksteps = ((p.K - k0 < TC_KC ? p.K - k0 : TC_KC) + 7) / 8;
#pragma unroll 1
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t ao = (uint32_t)(ks >> 2) * TC_SLAB_A + (uint32_t)(ks & 3) * 32u;
                        const uint32_t wo = (uint32_t)(ks >> 2) * (TC_NMAX * 128u) + (uint32_t)(ks & 3) * 32u;
                        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
                        umma_tf32(tmem_base, umma_desc_sw128(a_hi + ao), umma_desc_sw128(w_hi + wo), idesc, first);
                        umma_tf32(tmem_base + TC_D2_COL, umma_desc_sw128(a_lo + ao), umma_desc_sw128(w_hi + wo), idesc, first);
                        umma_tf32(tmem_base + TC_D2_COL, umma_desc_sw128(a_hi + ao), umma_desc_sw128(w_lo + wo), idesc, 1u);
                    }
                    // arrives on the mbarrier once every MMA issued so far has completed (implies fence::before_thread_sync)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&mma_bar))
                                 : "memory");
                }
                mbar_wait(&mma_bar, bar_phase);   // smem chunks are free again, accumulator chunk is complete
                bar_phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            // ---- epilogue: thread <-> row (TMEM lane = 32*(warp&3) + lane); warps 0-3 take the even 32-column
            //      groups, warps 4-7 the odd ones
            const int q = warp & 3;
            const int64_t m = m0 + q * 32 + lane;
            int agent = 0;
            if (p.epi == EPI_FC1 && m < p.M) agent = (int)((m % g.bv.R) % g.bv.N);
            for (int c0 = (warp >> 2) * 32; c0 < nw; c0 += 64) {
                float v[32], v2[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TC_D2_COL + c0), v2);
                if (m < p.M) {
                    float *yrow = p.Y + m * p.ldy + n0 + c0;
                    const bool vec = ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0) && (n0 + c0 + 32 <= p.Nout);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = n0 + c0 + i;
                        if (n < p.Nout) {
                            float x = v[i] + v2[i];
                            if (p.bias) x += __ldg(p.bias + n);
                            if (p.epi == EPI_FC1) { x += __ldg(p.W + (int64_t)n * p.ldw + p.K + agent); x = fmaxf(x, 0.0f); }
                            else if (p.epi == EPI_RELU) x = fmaxf(x, 0.0f);
                            else if (p.epi == EPI_MASKPOS) x = (p.aux[m * p.ld_aux + n] > 0.0f) ? x : 0.0f;
                            v[i] = x;
                        }
                    }
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4 *>(yrow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (n0 + c0 + i < p.Nout) yrow[i] = v[i];
                    }
                }
            }
            // TMEM is overwritten by the next tile's first MMA: order the loads before it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
