// Tensor-core GRU recurrence for LARGE row counts (q_learner.py:49-51, 60-62 at 20v20 / B = 1024: 40 960 chains).
//
// The FFMA recurrences of gru_rec.cuh keep W_hh in registers and give every chain its own 64-thread CTA: ~280 SM-cycles
// per chain-step, fine while there are at most a few chains per SM.  Here a CTA owns a TILE of up to 128 chains and the
// per-step matvec of all of them is ONE 3xTF32 tcgen05 GEMM,   gh[128 x 192] = h_{t-1}[128 x 64] . W_hh^T ,
// with W_hh (hi + lo, 96 KB) resident in shared memory for the whole kernel and the accumulator in TMEM:
//   thread 0    : 16 correction MMAs (lo.hi, hi.lo) then 8 hi.hi MMAs (M = 128, N = 192, K = 8) into ONE accumulator --
//                 the 2^-11-scaled terms accumulate while the accumulator is still small, so the tensor core's
//                 truncating adds cost the same as with the separate accumulators of tc_gemm.cuh --, tcgen05.commit;
//   16 warps    : warp w reads TMEM lane quarter w & 3 (lane = chain) and hidden units 16 (w >> 2) .. +16: gate math
//                 (MUFU sigmoid / tanh, the same formulas as k_gru_fwd9), h_t split hi / lo straight into the A operand
//                 of the next step's MMA (K-major SWIZZLE_128B), h_t and the saved gates transposed through a 4 KB
//                 per-warp shared-memory tile so that global stores are full 64 / 128-byte row segments.
// gi = W_ih x + b_ih arrives in a TILED layout written by k_agent_in_tc (AgentInArgs.gi_tiled):
//     gi_tiled[m / 32][chunk c = 0..47][m % 32][4 floats]        (m = t R + row; needs R % 32 == 0)
// so that "lane = chain" loads are fully coalesced 512-byte requests with no shared-memory staging; a thread's twelve
// float4 of step t are requested BEFORE it waits for the step's MMAs (HBM latency hides under the tensor core).
// h_{t-1} of a thread's own (chain, 16 units) stays in registers for the whole time loop.
// Per tile-step: 24 MMAs (~2.4 K cycles) + epilogue (MUFU-bound, ~3 K cycles); 20v20 / B = 1024 runs 320 tiles in three
// waves.  HBM floor of the launch (gi in, h + gates out): 12.6 GB = 1.9 ms.
#pragma once
#include "tc_gemm.cuh"

#define GT_THREADS 512
#define GT_SLAB_A (TC_M * 128)                 // 16 KB: [128 rows x 32 floats]
#define GT_SLAB_W (G3 * 128)                   // 24 KB: [192 rows x 32 floats]
#define GT_OFF_W (4 * GT_SLAB_A)               // A hi | lo: 64 KB
#define GT_OFF_SCR (GT_OFF_W + 4 * GT_SLAB_W)  // W_hh hi | lo: 96 KB
#define GT_SCR_WARP 4096                       // per-warp transposition tile: 32 rows x 8 float4
#define GT_SMEM_BYTES (GT_OFF_SCR + (GT_THREADS / 32) * GT_SCR_WARP)   // 224 KB
#define GT_TMEM_COLS 256                       // one 192-column accumulator
#define GI_TILE_CHUNKS (G3 / 4)                // 48 float4 chunks per gi row

struct GruFwdTcArgs {
    GruFwdArgs g;
    int groups_per_tile;       // 32-chain groups per CTA tile (1..4)
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__global__ void __launch_bounds__(GT_THREADS, 1) k_gru_fwd_tc(const __grid_constant__ GruFwdTcArgs ta) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bhh_s[G3];
    const GruFwdArgs &a = ta.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, net = blockIdx.y;
    const int q = warp & 3, cg = warp >> 2;              // TMEM lane quarter, 16-unit column group
    const int G = a.R >> 5;                              // 32-chain groups per net
    const int g0 = (int)blockIdx.x * ta.groups_per_tile;
    if (g0 >= G) return;
    const int ng = (G - g0 < ta.groups_per_tile) ? G - g0 : ta.groups_per_tile;
    const bool active = q < ng;                          // this warp's 32 chains exist (warp-uniform)
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const float *P = a.params[net];
    const float *gi = a.gi[net];
    float *hout = a.hout[net];
    float4 *gates4 = net == 0 ? reinterpret_cast<float4 *>(a.gates) : nullptr;
    uint8_t *A_hi = tc_smem, *A_lo = tc_smem + 2 * GT_SLAB_A;
    uint8_t *W_hi = tc_smem + GT_OFF_W, *W_lo = W_hi + 2 * GT_SLAB_W;
    float4 *stg = reinterpret_cast<float4 *>(tc_smem + GT_OFF_SCR + warp * GT_SCR_WARP);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)GT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < G3) bhh_s[tid] = __ldg(P + L.b_hh + tid);
    {   // W_hh [192 x 64] -> K-major SWIZZLE_128B hi / lo images (step constants: staged before the dependency wait)
        const int c4 = tid & 15, rbase = tid >> 4;
        float4 wv[G3 / 32];
#pragma unroll
        for (int i = 0; i < G3 / 32; ++i)
            wv[i] = __ldg(reinterpret_cast<const float4 *>(P + L.w_hh + (int64_t)(rbase + 32 * i) * HID + 4 * c4));
#pragma unroll
        for (int i = 0; i < G3 / 32; ++i) {
            const int j = rbase + 32 * i;
            split_store_fast(W_hi, W_lo, (uint32_t)(c4 >> 3) * GT_SLAB_W + (uint32_t)j * 128u + (uint32_t)(((c4 & 7) ^ (j & 7)) << 4), wv[i]);
        }
    }
    // A operand: zeros (h_{-1} = 0; rows of absent groups stay zero for the whole kernel)
#pragma unroll
    for (int i = 0; i < 4 * GT_SLAB_A / (16 * GT_THREADS); ++i)
        reinterpret_cast<uint4 *>(tc_smem)[tid + GT_THREADS * i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                     // the zero fill precedes the h_{t0-1} rows below
    pdl_wait();                                          // gi (and hout for t0 > 0) come from the stream predecessor

    const int t0 = a.t0, t1 = a.t1;
    const int r = q * 32 + lane;                         // tile row = TMEM lane
    const int64_t row0 = (int64_t)g0 * 32;               // first chain of the tile
    float hprev[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) hprev[e] = 0.0f;
    if (t0 > 0 && active) {
        const float4 *hp = reinterpret_cast<const float4 *>(hout + (((int64_t)(t0 - 1) * a.R + row0 + r) * HID + cg * 16));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 v = hp[c];
            hprev[4 * c] = v.x; hprev[4 * c + 1] = v.y; hprev[4 * c + 2] = v.z; hprev[4 * c + 3] = v.w;
            const int cc = cg * 4 + c;
            split_store_fast(A_hi, A_lo, (uint32_t)(cc >> 3) * GT_SLAB_A + (uint32_t)r * 128u + (uint32_t)(((cc & 7) ^ (r & 7)) << 4), v);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    uint32_t bar_phase = 0;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(G3 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16);
    // gi_tiled[(t G + g0 + q)][gate * 16 + cg * 4 + c][lane][4]
    const float4 *gi4 = reinterpret_cast<const float4 *>(gi) + ((int64_t)(g0 + q) * GI_TILE_CHUNKS + cg * 4) * 32 + lane;
    const int64_t gi_tstride = (int64_t)G * GI_TILE_CHUNKS * 32;          // float4 per timestep

    for (int t = t0; t < t1; ++t) {
        if (t + PDL_LEAD_STEPS == t1) pdl_trigger();
        if (tid == 0) {
            const uint64_t dA_hi = umma_desc_sw128(smem_u32(A_hi)), dA_lo = umma_desc_sw128(smem_u32(A_lo));
            const uint64_t dW_hi = umma_desc_sw128(smem_u32(W_hi)), dW_lo = umma_desc_sw128(smem_u32(W_lo));
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {                      // the small correction products first
                const uint64_t ao = (uint64_t)(((ks >> 2) * GT_SLAB_A + (ks & 3) * 32) >> 4);
                const uint64_t wo = (uint64_t)(((ks >> 2) * GT_SLAB_W + (ks & 3) * 32) >> 4);
                umma_tf32(tmem_base, dA_lo + ao, dW_hi + wo, idesc, ks == 0 ? 0u : 1u);
                umma_tf32(tmem_base, dA_hi + ao, dW_lo + wo, idesc, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                const uint64_t ao = (uint64_t)(((ks >> 2) * GT_SLAB_A + (ks & 3) * 32) >> 4);
                const uint64_t wo = (uint64_t)(((ks >> 2) * GT_SLAB_W + (ks & 3) * 32) >> 4);
                umma_tf32(tmem_base, dA_hi + ao, dW_hi + wo, idesc, 1u);
            }
            umma_commit(&mma_bar);
        }
        // this step's gi: requested right behind the MMA issue: the HBM latency hides under the tensor core
        float4 gv[3][4];
        if (active) {
            const float4 *gp = gi4 + (int64_t)t * gi_tstride;
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int c = 0; c < 4; ++c) gv[g][c] = __ldg(gp + (g * 16 + c) * 32);
        }
        mbar_wait(&mma_bar, bar_phase);
        bar_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (active) {
            const int64_t mw = (int64_t)t * a.R + row0 + q * 32;           // first global row of this warp's group
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t ar[8], az[8], an[8];
                tmem_ld8_nowait(tl + (uint32_t)(half * 8), ar);
                tmem_ld8_nowait(tl + (uint32_t)(HID + half * 8), az);
                tmem_ld8_nowait(tl + (uint32_t)(2 * HID + half * 8), an);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) {
                    // four units in lock-step: four independent MUFU chains per thread
                    const int c = half * 2 + qt;                          // float4 chunk of the thread's 16 units
                    const float4 bn4 = *reinterpret_cast<const float4 *>(&bhh_s[2 * HID + cg * 16 + 4 * c]);
                    const float g_r[4] = {gv[0][c].x, gv[0][c].y, gv[0][c].z, gv[0][c].w};   // gi_r + b_ih_r + b_hh_r
                    const float g_z[4] = {gv[1][c].x, gv[1][c].y, gv[1][c].z, gv[1][c].w};
                    const float g_n[4] = {gv[2][c].x, gv[2][c].y, gv[2][c].z, gv[2][c].w};
                    const float bn[4] = {bn4.x, bn4.y, bn4.z, bn4.w};
                    float ea[4], eb[4], rr[4], zz[4], nn[4], ghn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // sigmoid(x) = 1 / (1 + 2^(-x log2 e)); exponents clamped at 63 so that the product below is finite
                        ea[j] = ex2_approx(fminf(-1.4426950408889634f * (__uint_as_float(ar[4 * qt + j]) + g_r[j]), 63.0f));
                        eb[j] = ex2_approx(fminf(-1.4426950408889634f * (__uint_as_float(az[4 * qt + j]) + g_z[j]), 63.0f));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                         // one reciprocal for both gates
                        const float pa = 1.0f + ea[j], pb = 1.0f + eb[j];
                        const float ip = rcp_approx(pa * pb);
                        rr[j] = ip * pb; zz[j] = ip * pa;
                        ghn[j] = __uint_as_float(an[4 * qt + j]) + bn[j];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) nn[j] = tanh_mufu(g_n[j] + rr[j] * ghn[j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int u = 4 * c + j;
                        hprev[u] = nn[j] + zz[j] * (hprev[u] - nn[j]);
                    }
                    if (gates4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) stg[lane * 8 + ((4 * qt + j) ^ (lane & 7))] = make_float4(rr[j], zz[j], nn[j], ghn[j]);
                    }
                }
                if (gates4) {                                             // [m][unit] float4: 128-byte row segments
                    __syncwarp();
                    const int ch = lane & 7, rsub = lane >> 3;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int rr_ = k * 4 + rsub;
                        gates4[(mw + rr_) * HID + cg * 16 + half * 8 + ch] = stg[rr_ * 8 + (ch ^ (rr_ & 7))];
                    }
                    __syncwarp();
                }
            }
            // h_t: A operand of the next step (hi / lo split) + global rows through the transposition tile
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 hv = make_float4(hprev[4 * c], hprev[4 * c + 1], hprev[4 * c + 2], hprev[4 * c + 3]);
                const int cc = cg * 4 + c;
                split_store_fast(A_hi, A_lo, (uint32_t)(cc >> 3) * GT_SLAB_A + (uint32_t)r * 128u + (uint32_t)(((cc & 7) ^ (r & 7)) << 4), hv);
                stg[lane * 8 + (c ^ (lane & 7))] = hv;
            }
            __syncwarp();
            {
                const int ch = lane & 3, rsub = lane >> 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr_ = k * 8 + rsub;
                    *reinterpret_cast<float4 *>(hout + (mw + rr_) * HID + cg * 16 + 4 * ch) = stg[rr_ * 8 + (ch ^ (rr_ & 7))];
                }
            }
        }
        // h_t is in the A operand (generic-proxy stores -> async proxy), the accumulator has been read: next step's MMAs
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)GT_TMEM_COLS));
}
