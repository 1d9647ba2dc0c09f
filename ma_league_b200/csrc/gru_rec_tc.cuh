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
    int gates_tiled;           // 1: saved gates as [m / 32][unit][m % 32] float4 (read by k_gru_bwd_tc); 0: [m][unit] float4 (k_gru_bwd9)
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
            float4 *gt4 = gates4 ? gates4 + ((mw >> 5) * HID + cg * 16) * 32 + lane : nullptr;
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
                        if (ta.gates_tiled) {                             // [m / 32][unit][m % 32] float4: one 512-byte store per unit
#pragma unroll
                            for (int j = 0; j < 4; ++j) gt4[(4 * c + j) * 32] = make_float4(rr[j], zz[j], nn[j], ghn[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) stg[lane * 8 + ((4 * qt + j) ^ (lane & 7))] = make_float4(rr[j], zz[j], nn[j], ghn[j]);
                        }
                    }
                }
                if (gates4 && !ta.gates_tiled) {                          // [m][unit] float4 through the transposition tile: 128-byte row segments
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

// =====================================================================================================================
// Tensor-core BPTT.  Per timestep (t = TT-1 .. 0) and tile of up to 128 chains:
//     dh[128 x 64] = [d r' | d z' | d gh_n](t+1)[128 x 192] . W_hh[192 x 64]  + carry + head seed
// as ONE 3xTF32 tcgen05 GEMM whose A operand lives in TENSOR MEMORY (tools/ta_probe.cu: lane = row, one 32-bit column
// per k): the epilogue threads (lane = chain) write the hi / lo halves of their d(gates) straight into TMEM columns with
// tcgen05.st -- no shared-memory operand tile, no swizzle arithmetic -- and only B = W_hh^T (hi + lo, K-major
// SWIZZLE_128B, 96 KB) sits in shared memory.   TMEM: D at 0 | A hi at 64 | A lo at 256 (448 of 512 columns).
//   thread 0 : 48 correction MMAs (lo.hi, hi.lo) then 24 hi.hi MMAs (M = 128, N = 64, K = 8), tcgen05.commit;
//   16 warps : warp w owns TMEM lane quarter w & 3 and hidden units 16 (w >> 2) .. +16.  The saved gates of its 32
//              chains x 16 units (tiled layout of k_gru_fwd_tc: its own 8 KB per step) arrive through a per-warp
//              cp.async slot requested at the end of the previous step's epilogue; h_{t-1} and the head seed are
//              requested into registers behind the MMA issue.  d_g keeps its row-major [m][256] layout (the dx GEMM
//              and k_reduce_gru read it): the four gate blocks of a half (8 units) go through the half's consumed gate
//              slot as a transposition tile and leave as 32-byte row pieces.
// =====================================================================================================================
#define GB_SLAB_B (HID * 128)                      // 8 KB: [64 rows x 32 floats]
#define GB_OFF_SLOT (12 * GB_SLAB_B)               // W_hh^T hi (6 slabs) | lo (6 slabs): 96 KB
#define GB_SLOT 8192                               // per warp: 16 units x 32 chains x float4
#define GB_SMEM_BYTES (GB_OFF_SLOT + (GT_THREADS / 32) * GB_SLOT)   // 224 KB
#define GB_COL_AHI 64
#define GB_COL_ALO 256

struct GruBwdTcArgs {
    GruBwdArgs g;
    int groups_per_tile;
};

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
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// the four values of a quarter as a TF32 hi / exact-remainder lo pair of TMEM column groups
__device__ __forceinline__ void tmem_st_split4(uint32_t t_hi, uint32_t t_lo, const float (&v)[4]) {
    uint32_t h[4];
    float l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = tf32_rna_fast(v[j]); l[j] = v[j] - __uint_as_float(h[j]); }
    tmem_st4(t_hi, h[0], h[1], h[2], h[3]);
    tmem_st4(t_lo, __float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
}

__global__ void __launch_bounds__(GT_THREADS, 1) k_gru_bwd_tc(const __grid_constant__ GruBwdTcArgs ta) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const GruBwdArgs &a = ta.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = warp >> 2;
    const int G = a.R >> 5;
    const int g0 = (int)blockIdx.x * ta.groups_per_tile;
    if (g0 >= G) return;
    const int ng = (G - g0 < ta.groups_per_tile) ? G - g0 : ta.groups_per_tile;
    const bool active = q < ng;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const int T = a.TT - 1;
    uint8_t *B_hi = tc_smem, *B_lo = tc_smem + 6 * GB_SLAB_B;
    float4 *slot4 = reinterpret_cast<float4 *>(tc_smem + GB_OFF_SLOT + warp * GB_SLOT);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B[n][k] = W_hh[k][n] (n = hidden column, k = gate row): coalesced float4 reads along n, scattered 4-byte stores
#pragma unroll
    for (int i = 0; i < G3 * (HID / 4) / GT_THREADS; ++i) {
        const int idx = tid + GT_THREADS * i;
        const int k = idx >> 4, n4 = idx & 15;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)k * HID + 4 * n4));
        const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t off = (uint32_t)(k >> 5) * GB_SLAB_B + sw128_off(4 * n4 + e, k & 31);
            const uint32_t h = tf32_rna_fast(ve[e]);
            *reinterpret_cast<uint32_t *>(B_hi + off) = h;
            *reinterpret_cast<float *>(B_lo + off) = ve[e] - __uint_as_float(h);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    // A = 0: nothing flows into the last timestep from the future
#pragma unroll
    for (int c = 0; c < 6; ++c) {                          // this warp's 16 columns of the six 64-column blocks (3 gates x hi / lo)
#pragma unroll
        for (int e = 0; e < 4; ++e) tmem_st4(tl + (uint32_t)(GB_COL_AHI + c * HID + cg * 16 + 4 * e), 0u, 0u, 0u, 0u);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    pdl_wait();                                            // gates / h / head seeds come from the stream predecessors

    const int r = q * 32 + lane;
    const int64_t row0 = (int64_t)g0 * 32;
    const float4 *gates4 = reinterpret_cast<const float4 *>(a.gates);
    auto fetch_gates = [&](int t) {                        // the warp's 8 KB of step t -> its slot (every thread: its own 16 float4)
        const float4 *src = gates4 + (((int64_t)t * G + g0 + q) * HID + cg * 16) * 32 + lane;
#pragma unroll
        for (int u = 0; u < 16; ++u) cp_async16(&slot4[u * 32 + lane], src + u * 32);
    };
    if (active) fetch_gates(a.TT - 1);
    cp_async_commit();
    float carry[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) carry[u] = 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t bar_phase = 0;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    for (int i = 0; i < a.TT; ++i) {
        const int t = a.TT - 1 - i;
        if (i + PDL_LEAD_STEPS == a.TT) pdl_trigger();
        if (tid == 0) {
            const uint64_t dB_hi = umma_desc_sw128(smem_u32(B_hi)), dB_lo = umma_desc_sw128(smem_u32(B_lo));
#pragma unroll
            for (int ks = 0; ks < G3 / 8; ++ks) {                          // the small correction products first
                const uint64_t wo = (uint64_t)(((ks >> 2) * GB_SLAB_B + (ks & 3) * 32) >> 4);
                umma_tf32_ta(tmem_base, tmem_base + (uint32_t)(GB_COL_ALO + 8 * ks), dB_hi + wo, idesc, ks == 0 ? 0u : 1u);
                umma_tf32_ta(tmem_base, tmem_base + (uint32_t)(GB_COL_AHI + 8 * ks), dB_lo + wo, idesc, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < G3 / 8; ++ks) {
                const uint64_t wo = (uint64_t)(((ks >> 2) * GB_SLAB_B + (ks & 3) * 32) >> 4);
                umma_tf32_ta(tmem_base, tmem_base + (uint32_t)(GB_COL_AHI + 8 * ks), dB_hi + wo, idesc, 1u);
            }
            umma_commit(&mma_bar);
        }
        // h_{t-1} and the head seed of this thread's (chain, 16 units): requested behind the MMA issue
        float4 hp4[4], dh4[4];
        if (active) {
            const int64_t m = (int64_t)t * a.R + row0 + r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                hp4[c] = t > 0 ? __ldg(reinterpret_cast<const float4 *>(a.hout + (m - a.R) * HID + cg * 16) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                dh4[c] = t < T ? __ldg(reinterpret_cast<const float4 *>(a.dh_head + m * HID + cg * 16) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        mbar_wait(&mma_bar, bar_phase);
        bar_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (active) {
            const int64_t mw = (int64_t)t * a.R + row0 + q * 32;
            cp_async_wait<0>();                                           // this thread's own gate copies have landed
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t acc[8];
                tmem_ld8_nowait(tl + (uint32_t)(cg * 16 + half * 8), acc);
                float4 g4[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) g4[e] = slot4[(half * 8 + e) * 32 + lane];
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                __syncwarp();                                             // the half's gate slot is consumed: it becomes the d_g transposition tile
                float4 *tile = slot4 + half * 256;                        // [gate block 4][chain 32][2 pieces of 16 B], pieces swizzled
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) {
                    const int c = half * 2 + qt;
                    const float hp[4] = {hp4[c].x, hp4[c].y, hp4[c].z, hp4[c].w};
                    const float dhh[4] = {dh4[c].x, dh4[c].y, dh4[c].z, dh4[c].w};
                    float drp[4], dzp[4], dnp[4], dghn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 g = g4[4 * qt + j];                  // (r, z, n, gh_n)
                        const float dh = __uint_as_float(acc[4 * qt + j]) + (carry[4 * c + j] + dhh[j]);
                        const float dn = dh * (1.0f - g.y);
                        const float dz = dh * (hp[j] - g.z);
                        dnp[j] = dn * (1.0f - g.z * g.z);
                        dzp[j] = dz * g.y * (1.0f - g.y);
                        drp[j] = dnp[j] * g.w * g.x * (1.0f - g.x);
                        dghn[j] = dnp[j] * g.x;
                        carry[4 * c + j] = dh * g.y;
                    }
                    // next step's A operand: [d r' | d z' | d gh_n], hi / lo, this thread's TMEM lane
                    const uint32_t ca = (uint32_t)(cg * 16 + 4 * c);
                    tmem_st_split4(tl + GB_COL_AHI + ca, tl + GB_COL_ALO + ca, drp);
                    tmem_st_split4(tl + GB_COL_AHI + HID + ca, tl + GB_COL_ALO + HID + ca, dzp);
                    tmem_st_split4(tl + GB_COL_AHI + 2 * HID + ca, tl + GB_COL_ALO + 2 * HID + ca, dghn);
                    const int sw = qt ^ ((lane >> 2) & 1);
                    tile[(0 * 32 + lane) * 2 + sw] = make_float4(drp[0], drp[1], drp[2], drp[3]);
                    tile[(1 * 32 + lane) * 2 + sw] = make_float4(dzp[0], dzp[1], dzp[2], dzp[3]);
                    tile[(2 * 32 + lane) * 2 + sw] = make_float4(dnp[0], dnp[1], dnp[2], dnp[3]);
                    tile[(3 * 32 + lane) * 2 + sw] = make_float4(dghn[0], dghn[1], dghn[2], dghn[3]);
                }
                __syncwarp();
                {   // d_g[m][gate block * 64 + unit]: two lanes per 32-byte row piece
                    const int p = lane & 1, rs = lane >> 1;
#pragma unroll
                    for (int gb = 0; gb < 4; ++gb)
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {
                            const int rr_ = pass * 16 + rs;
                            *reinterpret_cast<float4 *>(a.d_g + (mw + rr_) * (4 * HID) + gb * HID + cg * 16 + half * 8 + 4 * p) =
                                tile[(gb * 32 + rr_) * 2 + (p ^ ((rr_ >> 2) & 1))];
                        }
                }
            }
            __syncwarp();                                                 // every lane is done with the tiles: the slot may be refilled
            if (t > 0) fetch_gates(t - 1);
            cp_async_commit();
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
