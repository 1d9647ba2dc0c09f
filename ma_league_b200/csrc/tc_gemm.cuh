// Tensor-core panel GEMM for the batched (rows x 64) GRU / hypernet projections:
//     Y[m, n] = epi( sum_k A[m,k] * W[n,k] )         (same problem descriptors as k_linear_group)
//
// 5th-gen tensor cores (tcgen05.mma kind::tf32, SASS UTCHMMA) with the accumulators in TMEM.  Plain TF32 (10-bit
// mantissa) misses the 1e-5 parity bar, so every operand is split in two TF32 pieces, x = hi + lo, and three MMAs
// (hi*hi + lo*hi + hi*lo, fp32 accumulate) recover fp32-level accuracy ("3xTF32").  hi*hi accumulates in D1, the two
// 2^-11-scaled correction products in a separate accumulator D2 (the tensor core adds into the accumulator with
// truncation; keeping the small terms apart and summing D1 + D2 in fp32 in the epilogue keeps fp32-level accuracy:
// measured 1.4e-7 relative at K = 64, the same as the FFMA kernel).
//
// One CTA = 256 threads, persistent over 128-row M tiles; <= 80-wide N tiles (wider outputs are split into
// independent problems on the host) so that two CTAs fit per SM and overlap each other's phases.
// Per (m-tile, n-tile, 64-wide k-chunk):
//   all threads : load the A rows (float4, coalesced) through the fused loaders, split hi/lo, STS.128 into the
//                 canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunks XOR-ed with row&7);
//                 the W chunk likewise (kept resident across m-tiles when there is only one);  fence.proxy.async
//   thread 0    : 8 k-steps x 3 tcgen05.mma (M=128, N=n-tile, K=8), tcgen05.commit -> mbarrier
//   all threads : wait, tcgen05.ld 32x32b (TMEM lane = tile row), bias / ReLU / ReLU-mask / fc1 epilogue, STG.128
#pragma once
#include "mal_common.cuh"
#include "learner.cuh"

#define TC_M 128
#define TC_KC 64                  // k-chunk (floats) = two 128-byte swizzle slabs
#define TC_NMAX 80                // widest n-tile
#define TC_SLAB_A (TC_M * 128)    // bytes of one [128 rows x 32 floats] slab
#define TC_SLAB_W (TC_NMAX * 128)
#define TC_THREADS 256
#define TC_WARPS (TC_THREADS / 32)
#define TC_TMEM_COLS 256
#define TC_D2_COL 128             // TMEM column of the correction-term accumulator
#define TC_SMEM_BYTES (4 * TC_SLAB_A + 4 * TC_SLAB_W)

enum { TCA_VEC_DENSE = 0, TCA_VEC_STATE = 1, TCA_GENERIC = 2, TCA_AGENT = 3 };   // TCA_AGENT: fc1 input rows [obs | last-action one-hot], float4 loads over the obs part
enum { TCE_BIAS_ACT = 0, TCE_MASKPOS = 1, TCE_FC1 = 2 };

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// byte offset of element (row r, k in [0,32)) inside one K-major SWIZZLE_128B slab
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

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[32]) {   // fills r[0..15]
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// split a float4 into TF32 hi / lo pieces and store both with one STS.128 each
__device__ __forceinline__ void split_store(uint8_t *hi_base, uint8_t *lo_base, uint32_t off, float4 v) {
    uint4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - __uint_as_float(h.x)); l.y = tf32_rna(v.y - __uint_as_float(h.y));
    l.z = tf32_rna(v.z - __uint_as_float(h.z)); l.w = tf32_rna(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4 *>(hi_base + off) = h;
    *reinterpret_cast<uint4 *>(lo_base + off) = l;
}

// x = q*d + r for 0 <= x < 2^24 with a float reciprocal (exact after one correction step)
__device__ __forceinline__ void fast_divmod(int x, int d, float inv_d, int &q, int &r) {
    q = __float2int_rz(__int2float_rn(x) * inv_d);
    r = x - q * d;
    if (r < 0) { --q; r += d; }
    if (r >= d) { ++q; r -= d; }
}
// Transposed, coalesced epilogue of one 32-row x gw-column group (gw = 16 or 32) of a tile.
// TMEM hands every lane one ROW (d1 + d2 = the 3xTF32 sum); storing that way scatters 16-byte pieces over 32 cache
// lines per instruction.  The warp therefore transposes the group through a swizzled 4 KB shared-memory tile (16-byte
// chunk c of row r lives in slot c ^ (r & 7)) and stores with lanes along the columns: 8 lanes x 16 B = one full
// 128-byte line per row.  Bias / activation / ReLU-mask / fc1 agent-id term are applied after the transposition.
struct TcEpi {
    float *Y; const float *aux, *W;
    int64_t ldy, ld_aux, ldw;
    int M, Nout, K, R, N;
    bool relu, vec_ok;
};
template <int EK>
__device__ __forceinline__ void tc_epilogue_group(const TcEpi &e, float4 *stg, const uint32_t (&d1)[32], const uint32_t (&d2)[32],
                                                  int gw, int c0, int64_t mw, int lane, const float *bias_s, bool skip) {
    // ReLU-mask operand (x of the rows this lane will store): all eight float4 are requested up front, UNCONDITIONALLY
    // (row / column clamped into range instead of guarded) and in one basic block, so that their latencies overlap each
    // other and the shared-memory round trip.  The r2 profile of the dx GEMM had 36 % of all stall samples on eight serial
    // load -> compare pairs per group: each load sat in its own guarded block and its result was turned into predicate
    // bits right away, one DRAM latency after the other (mal_debug_linear M=1M K=64 N=64: 657 us masked vs 147 us plain).
    float4 avs[8];
    const bool mask_vec = EK == TCE_MASKPOS && e.vec_ok && e.Nout >= 4 && e.M > 0;
    if (mask_vec) {
        const int cpr_ = gw >> 2, ch_ = lane & (cpr_ - 1), rsub_ = lane / cpr_, rpp_ = 32 / cpr_;
        const int col_ = (c0 + 4 * ch_ + 4 <= e.Nout) ? c0 + 4 * ch_ : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int64_t m_ = mw + (k * rpp_ + rsub_ < 32 ? k * rpp_ + rsub_ : 31);
            if (m_ > (int64_t)e.M - 1) m_ = (int64_t)e.M - 1;
            avs[k] = __ldg(reinterpret_cast<const float4 *>(e.aux + m_ * e.ld_aux + col_));
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c < (gw >> 2))
            stg[lane * 8 + (c ^ (lane & 7))] =
                make_float4(__uint_as_float(d1[4 * c]) + __uint_as_float(d2[4 * c]),
                            __uint_as_float(d1[4 * c + 1]) + __uint_as_float(d2[4 * c + 1]),
                            __uint_as_float(d1[4 * c + 2]) + __uint_as_float(d2[4 * c + 2]),
                            __uint_as_float(d1[4 * c + 3]) + __uint_as_float(d2[4 * c + 3]));
    }
    __syncwarp();
    const int cpr = gw >> 2;                      // chunks per row: 8 or 4
    const int ch = lane & (cpr - 1), rsub = lane / cpr;
    const int rpp = 32 / cpr;                     // rows per pass: 4 or 8
    const int col = c0 + 4 * ch;
    float4 xs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = k * rpp + rsub;
        if (r < 32) xs[k] = stg[r * 8 + (ch ^ (r & 7))];
    }
    if (col < e.Nout && !skip) {
        const float4 b4 = *reinterpret_cast<const float4 *>(&bias_s[col]);
        const bool vec = e.vec_ok && col + 4 <= e.Nout;
        float *yp = e.Y + (mw + rsub) * e.ldy + col;
        const float *ap = (EK == TCE_MASKPOS) ? e.aux + (mw + rsub) * e.ld_aux + col : nullptr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = k * rpp + rsub;
            const int64_t m = mw + r;
            if (r < 32 && m < e.M) {
                float4 x = xs[k];
                x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                if (EK == TCE_FC1) {
                    const int agent = (int)((m % e.R) % e.N);
                    const float *wa = e.W + (int64_t)col * e.ldw + e.K + agent;
                    x.x = fmaxf(x.x + __ldg(wa), 0.0f);
                    if (col + 1 < e.Nout) x.y = fmaxf(x.y + __ldg(wa + e.ldw), 0.0f);
                    if (col + 2 < e.Nout) x.z = fmaxf(x.z + __ldg(wa + 2 * e.ldw), 0.0f);
                    if (col + 3 < e.Nout) x.w = fmaxf(x.w + __ldg(wa + 3 * e.ldw), 0.0f);
                } else if (EK == TCE_BIAS_ACT) {
                    if (e.relu) { x.x = fmaxf(x.x, 0.0f); x.y = fmaxf(x.y, 0.0f); x.z = fmaxf(x.z, 0.0f); x.w = fmaxf(x.w, 0.0f); }
                }
                float *y = yp + (int64_t)(k * rpp) * e.ldy;
                if (vec) {
                    if (EK == TCE_MASKPOS) {
                        const float4 av = mask_vec ? avs[k] : __ldg(reinterpret_cast<const float4 *>(ap + (int64_t)(k * rpp) * e.ld_aux));
                        x.x = av.x > 0.0f ? x.x : 0.0f; x.y = av.y > 0.0f ? x.y : 0.0f;
                        x.z = av.z > 0.0f ? x.z : 0.0f; x.w = av.w > 0.0f ? x.w : 0.0f;
                    }
                    *reinterpret_cast<float4 *>(y) = x;
                } else {
                    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (col + q < e.Nout) {
                            float val = xv[q];
                            if (EK == TCE_MASKPOS) val = (ap[(int64_t)(k * rpp) * e.ld_aux + q] > 0.0f) ? val : 0.0f;
                            y[q] = val;
                        }
                    }
                }
            }
        }
    }
    __syncwarp();                                 // the staging tile is rewritten by the next group
}

// Round-to-nearest (ties away) TF32 of a finite float in two integer instructions: add half an ulp of the 10-bit
// mantissa to the bit pattern, clear the 13 low bits.  Bit-identical to cvt.rna.tf32.f32 for finite inputs (which ptxas
// expands into ~5 instructions for its NaN / Inf handling: measured 120 of 380 instructions per warp and 32-row block of
// k_reduce_tc); Inf stays Inf, and a NaN still reaches the accumulator through the remainder v - hi.
__device__ __forceinline__ uint32_t tf32_rna_fast(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// hi = RNA-rounded TF32, lo = exact fp32 remainder (the tensor core ignores its 13 low mantissa bits)
__device__ __forceinline__ void split_store_fast(uint8_t *hi_base, uint8_t *lo_base, uint32_t off, float4 v) {
    uint4 h;
    float4 l;
    h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
    l.x = v.x - __uint_as_float(h.x); l.y = v.y - __uint_as_float(h.y);
    l.z = v.z - __uint_as_float(h.z); l.w = v.w - __uint_as_float(h.w);
    *reinterpret_cast<uint4 *>(hi_base + off) = h;
    *reinterpret_cast<float4 *>(lo_base + off) = l;
}

template <int AK, int EK>
__global__ void __launch_bounds__(TC_THREADS, 2) k_linear_tc(const __grid_constant__ LinGroup g) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[TC_NMAX + 16];
    uint8_t *A_hi = tc_smem, *A_lo = tc_smem + 2 * TC_SLAB_A;
    uint8_t *W_hi = tc_smem + 4 * TC_SLAB_A, *W_lo = W_hi + 2 * TC_SLAB_W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LinProb &p = g.p[blockIdx.y];
    const int n_mtiles = (p.M + TC_M - 1) / TC_M;
    if ((int)blockIdx.x >= n_mtiles) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TC_TMEM_COLS));
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

    const int K = p.K, M = p.M, Nout = p.Nout;
    const int nkc = (K + TC_KC - 1) / TC_KC;
    const int n_ntiles = (Nout + TC_NMAX - 1) / TC_NMAX;
    const int nt_w = (((Nout + n_ntiles - 1) / n_ntiles) + 15) & ~15;
    int w_cached = -1;    // (n-tile * nkc + kc) currently staged in W_hi / W_lo
    int bias_cached = -1;
    const bool relu = (p.epi == EPI_RELU);
    // vectorised staging map: thread -> float4 column c4 (16 per 64-wide chunk), rows r = tid/16 + 16*i
    const int c4 = tid & 15, rbase = tid >> 4;
    const uint32_t a_slab = (uint32_t)(c4 >> 3) * TC_SLAB_A, w_slab = (uint32_t)(c4 >> 3) * TC_SLAB_W;

    // vector loaders: the A rows of the CTA's NEXT m-tile are prefetched into registers while the tensor core and the
    // epilogue work on the current one (single n-tile, single k-chunk problems: fc1-sized K, w_ih, mixer layer 2)
    float4 pre[8];
    bool have_pre = false;
    const bool can_prefetch = (AK != TCA_GENERIC) && nkc == 1 && n_ntiles == 1;
    const float invR = 1.0f / (float)(g.bv.R > 0 ? g.bv.R : 1), invN = 1.0f / (float)(g.bv.N > 0 ? g.bv.N : 1);
    auto load_a = [&](int64_t m0_, int k0_, float4 (&v)[8]) {
        const int kcol = k0_ + 4 * c4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t m = m0_ + rbase + 16 * i;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && kcol < K && !(g.dbg & 2)) {
                if (AK == TCA_AGENT) {
                    int t, rr, b, n;
                    fast_divmod((int)m, g.bv.R, invR, t, rr);
                    fast_divmod(rr, g.bv.N, invN, b, n);
                    if (kcol < g.bv.OBS) {       // OBS % 4 == 0 (host-checked): the float4 stays inside the obs row
                        v[i] = __ldg(reinterpret_cast<const float4 *>(field_ptr<float>(g.bv.obs, b, t) + (int64_t)n * g.bv.OBS + kcol));
                    } else if (t > 0) {
                        const float *oh = field_ptr<float>(g.bv.onehot, b, t - 1) + (int64_t)n * g.bv.A;
                        const int a0 = kcol - g.bv.OBS;
                        if (a0 < g.bv.A) v[i].x = __ldg(oh + a0);
                        if (a0 + 1 < g.bv.A) v[i].y = __ldg(oh + a0 + 1);
                        if (a0 + 2 < g.bv.A) v[i].z = __ldg(oh + a0 + 2);
                        if (a0 + 3 < g.bv.A) v[i].w = __ldg(oh + a0 + 3);
                    }
                } else {
                    const float *rp;
                    if (AK == TCA_VEC_DENSE) rp = p.A + m * p.lda;
                    else {
                        const int b = (int)((unsigned)m / (unsigned)g.bv.T), t = (int)m - b * g.bv.T;
                        rp = field_ptr<float>(g.bv.state, b, t + p.shift);
                    }
                    v[i] = __ldg(reinterpret_cast<const float4 *>(rp + kcol));
                }
            }
        }
    };

    pdl_wait();                                                  // the A rows may come from the stream predecessor
    for (int mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
        const int64_t m0 = (int64_t)mt * TC_M;
        if (mt + (int)gridDim.x >= n_mtiles) pdl_trigger();      // last tile of this CTA
        for (int nt = 0; nt < n_ntiles; ++nt) {
            const int n0 = nt * nt_w;
            const int nw = (Nout - n0) < nt_w ? (((Nout - n0) + 15) & ~15) : nt_w;   // MMA N (multiple of 16)
            if (bias_cached != nt) {
                if (tid < TC_NMAX + 16) bias_s[tid] = (p.bias && n0 + tid < Nout) ? __ldg(p.bias + n0 + tid) : 0.0f;
                bias_cached = nt;    // made visible by the barrier below
            }
            for (int kc = 0; kc < nkc; ++kc) {
                const int k0 = kc * TC_KC;
                // ---------------- stage the A chunk [128 x 64]
                {   // (the epilogue's transposition reuses A_hi as scratch, so the chunk is staged for every n-tile)
                    if (AK == TCA_GENERIC) {
                        // lanes along k (coalesced scalar loads), 4 rows in flight per warp iteration
#pragma unroll 1
                        for (int rb = warp; rb < TC_M; rb += 4 * TC_WARPS) {
                            float v0[4], v1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int64_t m = m0 + rb + u * TC_WARPS;
                                v0[u] = 0.0f; v1[u] = 0.0f;
                                if (m < M) {
                                    RowSrc rs = resolve_row(p.a_kind, g.bv, p.A, p.lda, p.shift, m);
                                    if (k0 + lane < K) v0[u] = row_elem(p.a_kind, g.bv, rs, k0 + lane);
                                    if (k0 + 32 + lane < K) v1[u] = row_elem(p.a_kind, g.bv, rs, k0 + 32 + lane);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint32_t off = sw128_off(rb + u * TC_WARPS, lane);
                                const uint32_t h0 = tf32_rna(v0[u]), h1 = tf32_rna(v1[u]);
                                *reinterpret_cast<uint32_t *>(A_hi + off) = h0;
                                *reinterpret_cast<uint32_t *>(A_hi + TC_SLAB_A + off) = h1;
                                *reinterpret_cast<uint32_t *>(A_lo + off) = tf32_rna(v0[u] - __uint_as_float(h0));
                                *reinterpret_cast<uint32_t *>(A_lo + TC_SLAB_A + off) = tf32_rna(v1[u] - __uint_as_float(h1));
                            }
                        }
                    } else {
                        float4 v[8];
                        if (have_pre) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = pre[i];
                        } else {
                            load_a(m0, k0, v);
                        }
                        if (!(g.dbg & 4))
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rbase + 16 * i;
                            split_store_fast(A_hi, A_lo, a_slab + (uint32_t)r * 128u + (uint32_t)(((c4 & 7) ^ (r & 7)) << 4), v[i]);
                        }
                    }
                }
                // ---------------- stage the W chunk [nw x 64] unless it is already resident
                const int w_id = nt * nkc + kc;
                if (w_id != w_cached) {
                    if (!p.w_trans && (p.ldw & 3) == 0 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0)) {
                        const int kcol = k0 + 4 * c4;
                        float4 wv[TC_NMAX / 16];          // all loads in flight before the first conversion
#pragma unroll
                        for (int i = 0; i < TC_NMAX / 16; ++i) {
                            const int j = rbase + 16 * i, n = n0 + j;
                            wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (j < nw && n < Nout && kcol < K) wv[i] = __ldg(reinterpret_cast<const float4 *>(p.W + (int64_t)n * p.ldw + kcol));
                        }
#pragma unroll
                        for (int i = 0; i < TC_NMAX / 16; ++i) {
                            const int j = rbase + 16 * i;
                            if (j < nw) split_store_fast(W_hi, W_lo, w_slab + (uint32_t)j * 128u + (uint32_t)(((c4 & 7) ^ (j & 7)) << 4), wv[i]);
                        }
                    } else if (!p.w_trans) {
                        for (int idx = tid; idx < nw * TC_KC; idx += TC_THREADS) {
                            const int j = idx >> 6, kk = idx & 63;
                            const int n = n0 + j, k = k0 + kk;
                            float v = 0.0f;
                            if (n < Nout && k < K) v = __ldg(p.W + (int64_t)n * p.ldw + k);
                            const uint32_t off = (uint32_t)(kk >> 5) * TC_SLAB_W + sw128_off(j, kk & 31);
                            const uint32_t h = tf32_rna(v);
                            *reinterpret_cast<uint32_t *>(W_hi + off) = h;
                            *reinterpret_cast<float *>(W_lo + off) = v - __uint_as_float(h);
                        }
                    } else {   // W(n,k) = W[k*ldw + n]: lanes along n (coalesced)
                        for (int idx = tid; idx < nw * TC_KC; idx += TC_THREADS) {
                            const int kk = idx / nw, j = idx - kk * nw;
                            const int n = n0 + j, k = k0 + kk;
                            float v = 0.0f;
                            if (n < Nout && k < K) v = __ldg(p.W + (int64_t)k * p.ldw + n);
                            const uint32_t off = (uint32_t)(kk >> 5) * TC_SLAB_W + sw128_off(j, kk & 31);
                            const uint32_t h = tf32_rna(v);
                            *reinterpret_cast<uint32_t *>(W_hi + off) = h;
                            *reinterpret_cast<float *>(W_lo + off) = v - __uint_as_float(h);
                        }
                    }
                    w_cached = w_id;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // ---------------- MMAs: D[128 x nw] (+)= A_chunk . W_chunk^T, three TF32 products per k-step
                if (tid == 0) {
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nw >> 3) << 17) |
                                           ((uint32_t)(TC_M >> 4) << 24);
                    const uint64_t dA_hi = umma_desc_sw128(smem_u32(A_hi)), dA_lo = umma_desc_sw128(smem_u32(A_lo));
                    const uint64_t dW_hi = umma_desc_sw128(smem_u32(W_hi)), dW_lo = umma_desc_sw128(smem_u32(W_lo));
                    const int ksteps = ((K - k0 < TC_KC ? K - k0 : TC_KC) + 7) / 8;
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        if (ks < ksteps && !(g.dbg & 8)) {       // only the 14-bit start-address field changes between k-steps
                            const uint64_t ao = (uint64_t)(((ks >> 2) * TC_SLAB_A + (ks & 3) * 32) >> 4);
                            const uint64_t wo = (uint64_t)(((ks >> 2) * TC_SLAB_W + (ks & 3) * 32) >> 4);
                            const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
                            umma_tf32(tmem_base, dA_hi + ao, dW_hi + wo, idesc, first);
                            umma_tf32(tmem_base + TC_D2_COL, dA_lo + ao, dW_hi + wo, idesc, first);
                            umma_tf32(tmem_base + TC_D2_COL, dA_hi + ao, dW_lo + wo, idesc, 1u);
                        }
                    }
                    // arrives once every MMA issued so far has completed (implies tcgen05.fence::before_thread_sync)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&mma_bar))
                                 : "memory");
                }
                have_pre = false;
                if (can_prefetch && mt + (int)gridDim.x < n_mtiles) {      // next tile's rows fly under the MMA + epilogue
                    load_a((int64_t)(mt + (int)gridDim.x) * TC_M, 0, pre);
                    have_pre = true;
                }
                mbar_wait(&mma_bar, bar_phase);   // smem chunks are free again, this chunk's products are in TMEM
                bar_phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            // ---------------- epilogue: warp <-> TMEM lane quarter (warp & 3); warps 0-3 take the even 32-column groups,
            //                  warps 4-7 the odd ones; each group is transposed through the (now idle) A_hi staging
            //                  buffer so that the stores are full 128-byte lines (tc_epilogue_group)
            {
                const int q = warp & 3;
                const int64_t mw = m0 + q * 32;
                const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
                float4 *stg = reinterpret_cast<float4 *>(A_hi) + warp * 256;
                TcEpi e;
                e.Y = p.Y + n0; e.aux = p.aux ? p.aux + n0 : nullptr; e.W = p.W + (int64_t)n0 * p.ldw;
                e.ldy = p.ldy; e.ld_aux = p.ld_aux; e.ldw = p.ldw;
                e.M = M; e.Nout = Nout - n0; e.K = K; e.R = g.bv.R; e.N = g.bv.N; e.relu = relu;
                e.vec_ok = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.Y) & 15) == 0) &&
                           (EK != TCE_MASKPOS || (((p.ld_aux & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.aux) & 15) == 0)));
#pragma unroll 1
                for (int c0 = (warp >> 2) * 32; c0 < nw; c0 += 64) {
                    const int gw = (nw - c0) < 32 ? 16 : 32;
                    uint32_t d1[32], d2[32];
                    if (gw == 32) {
                        tmem_ld32_nowait(tlane + (uint32_t)c0, d1);
                        tmem_ld32_nowait(tlane + (uint32_t)(TC_D2_COL + c0), d2);
                    } else {
                        tmem_ld16_nowait(tlane + (uint32_t)c0, d1);
                        tmem_ld16_nowait(tlane + (uint32_t)(TC_D2_COL + c0), d2);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tc_epilogue_group<EK>(e, stg, d1, d2, gw, c0, mw, lane, bias_s, (g.dbg & 1) != 0);
                }
            }
            // TMEM is overwritten by the next tile's first MMA: order the loads before it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS));
}

// =============================================================================================
// k_linear_tc2: warp-specialised, pipelined version of the panel GEMM (same math and epilogues, Nout <= TC_NMAX).
// One persistent CTA per SM (288 threads) walks a flat sequence of (m-tile, k-chunk) steps with TWO shared-memory
// stages (A chunk + W chunk each) and TWO TMEM accumulator sets:
//   warps 0-3  loaders : wait empty[stage]; split/store the A chunk that was register-prefetched one step earlier
//                        (+ the W chunk unless the stage already holds it); issue the global loads of the next step;
//                        fence.proxy.async; arrive full[stage]
//   warp  8    MMA     : wait full[stage] (and acc_empty[set] on a tile's first chunk); 3 tcgen05.mma per k-step;
//                        tcgen05.commit -> empty[stage] (and -> acc_full[set] after the tile's last chunk)
//   warps 4-7  epilogue: wait acc_full[set]; tcgen05.ld; bias / ReLU / ReLU-mask / fc1 epilogue; STG.128;
//                        arrive acc_empty[set]
// so global-load latency, the operand split, tensor-core time and the TMEM -> registers -> global epilogue of
// neighbouring tiles all overlap.
// =============================================================================================
#define TC2_STAGE_A (4 * TC_SLAB_A)                 // A hi (2 slabs) + A lo (2 slabs) = 64 KB
#define TC2_STAGE_W (4 * TC_SLAB_W)                 // 40 KB
#define TC2_STAGE (TC2_STAGE_A + TC2_STAGE_W)
#define TC2_SMEM_BYTES (2 * TC2_STAGE)              // 208 KB: one CTA per SM
#define TC2_TMEM_COLS 512
#define TC2_ACC_COLS (2 * TC_NMAX)                  // D1 | D2 of one accumulator set
#define TC2_THREADS 288
#define TC2_LOADERS 128


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// hi = RNA-rounded TF32, lo = exact fp32 remainder (the tensor core ignores its 13 low mantissa bits)
__device__ __forceinline__ void split_store2(uint8_t *hi_base, uint8_t *lo_base, uint32_t off, float4 v) {
    uint4 h;
    float4 l;
    h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
    l.x = v.x - __uint_as_float(h.x); l.y = v.y - __uint_as_float(h.y);
    l.z = v.z - __uint_as_float(h.z); l.w = v.w - __uint_as_float(h.w);
    *reinterpret_cast<uint4 *>(hi_base + off) = h;
    *reinterpret_cast<float4 *>(lo_base + off) = l;
}
// four consecutive columns of one logical input row through the generic fused loaders (kept out of line)
__device__ __noinline__ float4 tc_load_row4(int a_kind, const BatchView &bv, const float *A, int64_t lda, int shift,
                                            int64_t m, int kcol, int K) {
    const RowSrc rs = resolve_row(a_kind, bv, A, lda, shift, m);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = row_elem(a_kind, bv, rs, kcol);
    if (kcol + 1 < K) v.y = row_elem(a_kind, bv, rs, kcol + 1);
    if (kcol + 2 < K) v.z = row_elem(a_kind, bv, rs, kcol + 2);
    if (kcol + 3 < K) v.w = row_elem(a_kind, bv, rs, kcol + 3);
    return v;
}

template <int AK, int EK>
__global__ void __launch_bounds__(TC2_THREADS, 1) k_linear_tc2(const __grid_constant__ LinGroup g) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], accf_bar[2], acce_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[TC_NMAX + 16];
    __shared__ __align__(16) float4 epi_s[4][32 * 8];   // per epilogue warp: 32 rows x 32 columns, 16-byte chunks XOR-swizzled
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LinProb &p = g.p[blockIdx.y];
    const int n_mtiles = (p.M + TC_M - 1) / TC_M;
    if ((int)blockIdx.x >= n_mtiles) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TC2_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], TC2_LOADERS);
            mbar_init(&empty_bar[i], 1);
            mbar_init(&accf_bar[i], 1);
            mbar_init(&acce_bar[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int K = p.K, M = p.M, Nout = p.Nout;
    const int nw = (Nout + 15) & ~15;                // MMA N (multiple of 16, <= TC_NMAX)
    if (tid < TC_NMAX + 16) bias_s[tid] = (p.bias && tid < Nout) ? __ldg(p.bias + tid) : 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();

    const int nkc = (K + TC_KC - 1) / TC_KC;
    const int n_my = (n_mtiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    const int J = n_my * nkc;
    const float invR = 1.0f / (float)(g.bv.R > 0 ? g.bv.R : 1), invN = 1.0f / (float)(g.bv.N > 0 ? g.bv.N : 1);
    const float invT = 1.0f / (float)(g.bv.T > 0 ? g.bv.T : 1);

    if (warp < 4) {
        // ================================================================== loaders
        const int c4 = tid & 15, rb = tid >> 4;      // float4 column c4 of the 64-wide chunk, rows rb + 8 i
        const uint32_t a_slab = (uint32_t)(c4 >> 3) * TC_SLAB_A, w_slab = (uint32_t)(c4 >> 3) * TC_SLAB_W;
        const uint32_t sw = (uint32_t)(((c4 & 7) ^ rb) << 4);    // (r & 7) == rb for every row of this thread
        int w_st0 = -1, w_st1 = -1;                  // k-chunk of W currently held by stage 0 / 1
        float4 pre[16];                              // the A chunk of the next step, in flight
        auto load_chunk = [&](int j) {
            const int q = j / nkc, kc = j - q * nkc;
            const int64_t m0 = (int64_t)((int)blockIdx.x + q * (int)gridDim.x) * TC_M;
            const int kcol = kc * TC_KC + 4 * c4;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int64_t m = m0 + rb + 8 * i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M && kcol < K && !(g.dbg & 2)) {
                    if (AK == TCA_VEC_DENSE) {
                        v = __ldg(reinterpret_cast<const float4 *>(p.A + m * p.lda + kcol));
                    } else if (AK == TCA_VEC_STATE) {
                        int b, t;
                        fast_divmod((int)m, g.bv.T, invT, b, t);
                        v = __ldg(reinterpret_cast<const float4 *>(field_ptr<float>(g.bv.state, b, t + p.shift) + kcol));
                    } else if (AK == TCA_AGENT) {
                        int t, rr, b, n;
                        fast_divmod((int)m, g.bv.R, invR, t, rr);
                        fast_divmod(rr, g.bv.N, invN, b, n);
                        if (kcol < g.bv.OBS) {       // OBS % 4 == 0 (host-checked): the float4 stays inside the obs row
                            v = __ldg(reinterpret_cast<const float4 *>(field_ptr<float>(g.bv.obs, b, t) + (int64_t)n * g.bv.OBS + kcol));
                        } else if (t > 0) {
                            const float *oh = field_ptr<float>(g.bv.onehot, b, t - 1) + (int64_t)n * g.bv.A;
                            const int a0 = kcol - g.bv.OBS;
                            if (a0 < g.bv.A) v.x = __ldg(oh + a0);
                            if (a0 + 1 < g.bv.A) v.y = __ldg(oh + a0 + 1);
                            if (a0 + 2 < g.bv.A) v.z = __ldg(oh + a0 + 2);
                            if (a0 + 3 < g.bv.A) v.w = __ldg(oh + a0 + 3);
                        }
                    } else {
                        v = tc_load_row4(p.a_kind, g.bv, p.A, p.lda, p.shift, m, kcol, K);
                    }
                }
                pre[i] = v;
            }
        };
        auto stage_w = [&](uint8_t *W_hi, uint8_t *W_lo, int k0) {
            if (!p.w_trans && (p.ldw & 3) == 0 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0)) {
                const int kcol = k0 + 4 * c4;
                float4 wv[TC_NMAX / 8];
#pragma unroll
                for (int i = 0; i < TC_NMAX / 8; ++i) {
                    const int j = rb + 8 * i;
                    wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < Nout && kcol < K) wv[i] = __ldg(reinterpret_cast<const float4 *>(p.W + (int64_t)j * p.ldw + kcol));
                }
#pragma unroll
                for (int i = 0; i < TC_NMAX / 8; ++i) {
                    const int j = rb + 8 * i;
                    if (j < nw) split_store2(W_hi, W_lo, w_slab + (uint32_t)j * 128u + sw, wv[i]);
                }
            } else if (!p.w_trans) {
                for (int idx = tid; idx < nw * TC_KC; idx += TC2_LOADERS) {
                    const int j = idx >> 6, kk = idx & 63;
                    const int k = k0 + kk;
                    float v = 0.0f;
                    if (j < Nout && k < K) v = __ldg(p.W + (int64_t)j * p.ldw + k);
                    const uint32_t off = (uint32_t)(kk >> 5) * TC_SLAB_W + sw128_off(j, kk & 31);
                    const uint32_t h = tf32_rna(v);
                    *reinterpret_cast<uint32_t *>(W_hi + off) = h;
                    *reinterpret_cast<float *>(W_lo + off) = v - __uint_as_float(h);
                }
            } else {   // W(n,k) = W[k*ldw + n]: lanes along n (coalesced)
                for (int idx = tid; idx < nw * TC_KC; idx += TC2_LOADERS) {
                    const int kk = idx / nw, j = idx - kk * nw;
                    const int k = k0 + kk;
                    float v = 0.0f;
                    if (j < Nout && k < K) v = __ldg(p.W + (int64_t)k * p.ldw + j);
                    const uint32_t off = (uint32_t)(kk >> 5) * TC_SLAB_W + sw128_off(j, kk & 31);
                    const uint32_t h = tf32_rna(v);
                    *reinterpret_cast<uint32_t *>(W_hi + off) = h;
                    *reinterpret_cast<float *>(W_lo + off) = v - __uint_as_float(h);
                }
            }
        };
        load_chunk(0);
        for (int j = 0; j < J; ++j) {
            const int s = j & 1, u = j >> 1;
            const int q = j / nkc, kc = j - q * nkc;
            uint8_t *A_hi = tc_smem + (size_t)s * TC2_STAGE, *A_lo = A_hi + 2 * TC_SLAB_A;
            uint8_t *W_hi = A_hi + TC2_STAGE_A, *W_lo = W_hi + 2 * TC_SLAB_W;
            if (u >= 1) mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));   // the MMAs of the stage's previous use are done
            if (!(g.dbg & 4))
#pragma unroll
            for (int i = 0; i < 16; ++i)
                split_store2(A_hi, A_lo, a_slab + (uint32_t)(rb + 8 * i) * 128u + sw, pre[i]);
            if (j + 1 < J) load_chunk(j + 1);
            if ((s ? w_st1 : w_st0) != kc) {
                stage_w(W_hi, W_lo, kc * TC_KC);
                if (s) w_st1 = kc; else w_st0 = kc;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
            mbar_arrive(&full_bar[s]);
        }
    } else if (warp == 8) {
        // ================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nw >> 3) << 17) |
                                   ((uint32_t)(TC_M >> 4) << 24);
            for (int j = 0; j < J; ++j) {
                const int s = j & 1, u = j >> 1;
                const int q = j / nkc, kc = j - q * nkc;
                const int a = q & 1, v = q >> 1;
                mbar_wait(&full_bar[s], (uint32_t)(u & 1));
                if (kc == 0 && v >= 1) mbar_wait(&acce_bar[a], (uint32_t)((v - 1) & 1));   // epilogue drained this set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // descriptors: only the 14-bit start-address field changes between k-steps (+2 per 32 bytes)
                const uint32_t a_hi = smem_u32(tc_smem + (size_t)s * TC2_STAGE);
                const uint64_t dA_hi = umma_desc_sw128(a_hi), dA_lo = dA_hi + ((2 * TC_SLAB_A) >> 4);
                const uint64_t dW_hi = dA_hi + (TC2_STAGE_A >> 4), dW_lo = dW_hi + ((2 * TC_SLAB_W) >> 4);
                const uint32_t d1 = tmem_base + (uint32_t)(a * TC2_ACC_COLS), d2 = d1 + TC_NMAX;
                const int k0 = kc * TC_KC;
                const int ksteps = ((K - k0 < TC_KC ? K - k0 : TC_KC) + 7) / 8;
#pragma unroll
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
                    if (ks < ksteps && !(g.dbg & 8)) {
                        const uint64_t ao = (uint64_t)(((ks >> 2) * TC_SLAB_A + (ks & 3) * 32) >> 4);
                        const uint64_t wo = (uint64_t)(((ks >> 2) * TC_SLAB_W + (ks & 3) * 32) >> 4);
                        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
                        umma_tf32(d1, dA_hi + ao, dW_hi + wo, idesc, first);
                        umma_tf32(d2, dA_lo + ao, dW_hi + wo, idesc, first);
                        umma_tf32(d2, dA_hi + ao, dW_lo + wo, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);                      // stage s may be refilled once these MMAs are done
                if (kc == nkc - 1) umma_commit(&accf_bar[a]);    // ... and the tile's accumulators are complete
            }
        }
    } else {
        // ================================================================== epilogue (warps 4-7 <-> TMEM lane quarters)
        // TMEM hands every lane one ROW of the tile; storing that way scatters 16-byte pieces over 32 cache lines per
        // instruction.  Each warp therefore transposes 32-column groups through a swizzled 4 KB shared-memory tile and
        // stores with lanes along the columns (8 lanes x 16 B = one full 128-byte line per row, 4 rows per instruction);
        // bias / activation / ReLU-mask / fc1 agent-id term are applied after the transposition, on coalesced operands.
        const int wq = warp - 4;
        const bool relu = (p.epi == EPI_RELU);
        float4 *stg = epi_s[wq];
        TcEpi epi;
        epi.Y = p.Y; epi.aux = p.aux; epi.W = p.W; epi.ldy = p.ldy; epi.ld_aux = p.ld_aux; epi.ldw = p.ldw;
        epi.M = M; epi.Nout = Nout; epi.K = K; epi.R = g.bv.R; epi.N = g.bv.N; epi.relu = relu;
        epi.vec_ok = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) &&
                     (EK != TCE_MASKPOS || (((p.ld_aux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0)));
        for (int q = 0; q < n_my; ++q) {
            const int a = q & 1, v = q >> 1;
            mbar_wait(&accf_bar[a], (uint32_t)(v & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t mw = (int64_t)((int)blockIdx.x + q * (int)gridDim.x) * TC_M + wq * 32;   // first row of the warp
            const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * TC2_ACC_COLS);
#pragma unroll 1
            for (int c0 = 0; c0 < nw; c0 += 32) {
                const int gw = (nw - c0) < 32 ? 16 : 32;      // group width (nw is a multiple of 16)
                uint32_t d1[32], d2[32];
                if (gw == 32) {
                    tmem_ld32_nowait(tlane + (uint32_t)c0, d1);
                    tmem_ld32_nowait(tlane + (uint32_t)(TC_NMAX + c0), d2);
                } else {
                    tmem_ld16_nowait(tlane + (uint32_t)c0, d1);
                    tmem_ld16_nowait(tlane + (uint32_t)(TC_NMAX + c0), d2);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_epilogue_group<EK>(epi, stg, d1, d2, gw, c0, mw, lane, bias_s, (g.dbg & 1) != 0);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acce_bar[a]);               // this accumulator set may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC2_TMEM_COLS));
}
