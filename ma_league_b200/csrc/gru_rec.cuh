// GRU recurrences of the learner step (q_learner.py:49-51, 60-62 and their autograd transposes).
// One batch row (chain) per CTA, 64 threads, several CTAs per SM; thread = hidden unit for the gate math, the
// saved-state ring and the stores.  Everything that does not depend on h_{t-1} was hoisted into the batched GEMMs
// (gi = W_ih x + b_ih); what remains per timestep is gh = W_hh h_{t-1} with W_hh resident in REGISTERS, the gate math
// (MUFU ex2/rcp sigmoid and tanh), one h / d(gates) exchange through shared memory and ONE barrier between the two warps.
// gi[t] (forward) and saved gates / h_{t-1} / head seed (backward) stream through per-thread cp.async rings PF steps
// ahead.  Saved gates layout (private to these two kernels): [m][unit][r, z, n, gh_n].
//
// Register-tiled matvec.  With one thread per output every lane needs the whole h_{t-1} (64 floats) resp. d(gates) vector
// (192 floats) in its own registers: 16 resp. 48 broadcast LDS.128 per thread-step, each writing 512 B of registers per
// warp -- operand delivery, not the FMA pipe, set the pace of the backward kernel (measured r2: 742 -> 648 cycles per
// step at 160 chains, 2 866 -> 2 414 at 1 280 with the tiling below; two lanes per unit with half of K each move the same
// bytes and gained nothing, and neither did one persistent 128-thread CTA per SM owning all of its chains).
// A thread owns a U x (K/S) tile with U = S (still 192 weights in registers):
//     forward   U = S = 2: units (2g, 2g+1) x half of the 64 columns          ->  8 LDS.128 + 96 FFMA2 per thread-step
//     backward  U = S = 4: columns (4g .. 4g+3) x a quarter of the 192 rows   -> 12 LDS.128 + 96 FFMA2
// and the S partial sums per output meet in a transposing xor-shuffle reduction after which lane p of a group holds the
// complete sum of ITS unit / column.  The S sub-ranges of the shared operand are shifted by 16 bytes against each other
// so that the S addresses of one LDS.128 fall into different banks.  fp32 throughout (FFMA2 = two fp32 FMAs per
// instruction; it issues at half rate, i.e. the same FMA-pipe time as FFMA in half the issue slots).
//
// What an SM hosts sets the pace at league batch sizes (tools/gru_bench.cu, cycles per timestep of k_gru_fwd9 against the
// chains per SM): 527 / 562 / 851 / 878 for 1 / 2 / 3 / 4 -- up to two chains every warp has a sub-partition to itself, a
// third chain costs the FMA-pipe time of one warp-step.  k_gru_fwd9 therefore has a BALANCED mode (GruFwdArgs.bal_D): the
// chains x TT steps are laid end to end and dealt out to 2 x SMs workers of equal length, a chain changes workers at most
// once and hands its hidden state over through global memory and a flag (DESIGN.md section 4 has the schedule, the
// measurements and when the launcher picks it).
#pragma once
#include <type_traits>
#include "learner.cuh"

#define GT_PAD 4                           // floats between consecutive K sub-ranges of a shared operand row

template <int DBG = 0>
__global__ void __launch_bounds__(64, 4) k_gru_fwd7(GruFwdArgs a) {
    constexpr int PF = 8;
    constexpr int KC = HID / 2;            // columns per lane
    constexpr int HROW = HID + GT_PAD;
    __shared__ __align__(16) float h_s[2][HROW];
    __shared__ float st_s[PF][3][HID];
    __shared__ float pdl_anchor_s[HID];
    const int i = threadIdx.x, net = blockIdx.y, row = blockIdx.x;
    const int ug = i >> 1, part = i & 1;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const float *P = net ? a.params[1] : a.params[0];
    const float *gi = net ? a.gi[1] : a.gi[0];
    float *hout = net ? a.hout[1] : a.hout[0];
    float4 *gates = net == 0 ? reinterpret_cast<float4 *>(a.gates) : nullptr;
    const int t0 = a.t0, t1 = a.t1;

    float anchor = 0.0f;
    unsigned long long w[2][3][KC / 2];    // packed pairs of W_hh[g*64 + 2*ug + u][part*32 + 2j, +1]
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float4 *wr = reinterpret_cast<const float4 *>(P + L.w_hh + (int64_t)(g * HID + 2 * ug + u) * HID + part * KC);
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                const float4 v = __ldg(wr + q);
                w[u][g][2 * q] = pack2(v.x, v.y); w[u][g][2 * q + 1] = pack2(v.z, v.w);
                anchor += v.x;
            }
        }
    const float b_r = __ldg(P + L.b_hh + i), b_z = __ldg(P + L.b_hh + HID + i), b_n = __ldg(P + L.b_hh + 2 * HID + i);
    // step constants are loaded BEFORE the dependency wait; the shared store of a value that depends on every load pins them above it (ptxas sinks free-standing loads below the wait)
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[i]) = anchor + b_r + b_z + b_n;
    pdl_wait();
    float hprev = t0 > 0 ? hout[((int64_t)(t0 - 1) * a.R + row) * HID + i] : 0.0f;   // init_hidden: zeros
    const int hpos = i + (i >= KC ? GT_PAD : 0);
    h_s[0][hpos] = hprev; h_s[1][hpos] = 0.0f;
    const int64_t tstride = (int64_t)a.R * G3;
    const float *p_g = gi + (int64_t)row * G3 + i;
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        if (t0 + p < t1) {
#pragma unroll
            for (int g = 0; g < 3; ++g) cp_async4(&st_s[p][g][i], p_g + (int64_t)(t0 + p) * tstride + g * HID);
        }
        cp_async_commit();
    }
    __syncthreads();

    for (int tt = t0; tt < t1; tt += 2) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int t = tt + p;
            if (t >= t1) break;
            const int buf = p;                   // == (t - t0) & 1
            const int slot = (t - t0) % PF;
            if (t + PDL_LEAD_STEPS == t1) pdl_trigger();
            cp_async_wait<PF - 1>();             // this thread's group of step t has landed
            const float g_r = st_s[slot][0][i], g_z = st_s[slot][1][i], g_n = st_s[slot][2][i];
            if (t + PF < t1) {
#pragma unroll
                for (int g = 0; g < 3; ++g) cp_async4(&st_s[slot][g][i], p_g + (int64_t)(t + PF) * tstride + g * HID);
            }
            cp_async_commit();
            const float4 *hp = reinterpret_cast<const float4 *>(h_s[buf] + part * (KC + GT_PAD));
            unsigned long long s[2][3];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 3; ++g) s[u][g] = 0ull;
#pragma unroll
            for (int q = 0; q < ((DBG & 8) ? 1 : KC / 4); ++q) {
                const float4 hv = hp[q];
                const unsigned long long hxy = pack2(hv.x, hv.y), hzw = pack2(hv.z, hv.w);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        s[u][g] = fma2(w[u][g][2 * q], hxy, s[u][g]);
                        s[u][g] = fma2(w[u][g][2 * q + 1], hzw, s[u][g]);
                    }
            }
            // transposing reduction over the lane pair: lane `part` ends with the complete sums of unit 2*ug + part
            float x[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float a0, a1, c0, c1;
                unpack2(s[0][g], a0, a1); unpack2(s[1][g], c0, c1);
                const float p0 = a0 + a1, p1 = c0 + c1;          // this lane's partial for units 2ug, 2ug+1
                const float mine = part ? p1 : p0, other = part ? p0 : p1;
                x[g] = mine + __shfl_xor_sync(0xffffffffu, other, 1);
            }
            const float xr = x[0] + (g_r + b_r), xz = x[1] + (g_z + b_z), ghn = x[2] + b_n;
            const float rr = sigmoid_mufu(xr), zz = sigmoid_mufu(xz);
            const float nn = tanh_mufu(g_n + rr * ghn);
            const float hn = nn + zz * (hprev - nn);
            hprev = hn;
            h_s[buf ^ 1][hpos] = hn;
            if (!(DBG & 1)) {
                const int64_t m = (int64_t)t * a.R + row;
                hout[m * HID + i] = hn;
                if (gates) gates[m * HID + i] = make_float4(rr, zz, nn, ghn);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(64, 4) k_gru_bwd7(GruBwdArgs a) {
    constexpr int PF = 8;
    constexpr int JC = G3 / 4;             // rows j of W_hh per lane (48)
    constexpr int DSEG = JC + GT_PAD;      // padded quarter of the d(gates) operand
    __shared__ __align__(16) float dg_s[2][4 * DSEG];   // d gi_r | d gi_z | d gh_n of the previous step, in four shifted quarters
    __shared__ __align__(16) float4 g4_s[PF][HID];
    __shared__ float hp_s[PF][HID], dh_s[PF][HID];
    __shared__ float pdl_anchor_s[HID];
    const int k = threadIdx.x, row = blockIdx.x;
    const int kg = k >> 2, part = k & 3;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const int T = a.TT - 1;

    float anchor = 0.0f;
    for (int idx = k; idx < 2 * 4 * DSEG; idx += HID) (&dg_s[0][0])[idx] = 0.0f;
    unsigned long long wT[4][JC / 2];      // packed pairs (W_hh[48*part + 2j][4kg + c], W_hh[48*part + 2j + 1][4kg + c])
#pragma unroll
    for (int j = 0; j < JC / 2; ++j) {
        const int jj = part * JC + 2 * j;
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)jj * HID + 4 * kg));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)(jj + 1) * HID + 4 * kg));
        wT[0][j] = pack2(w0.x, w1.x); wT[1][j] = pack2(w0.y, w1.y); wT[2][j] = pack2(w0.z, w1.z); wT[3][j] = pack2(w0.w, w1.w);
        anchor += w0.x + w1.y;
    }
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[k]) = anchor;   // pins the loads above the wait (see k_gru_fwd7)

    pdl_wait();                                  // W_hh is a step constant; gates / h / dh_head come from the predecessors
    const float4 *gates4 = reinterpret_cast<const float4 *>(a.gates);
    auto fetch = [&](int i, int slot) {          // step i <-> t = TT-1-i
        const int t = a.TT - 1 - i;
        const int64_t m = (int64_t)t * a.R + row;
        cp_async16(&g4_s[slot][k], gates4 + m * HID + k);
        cp_async4(&hp_s[slot][k], a.hout + (t > 0 ? (m - a.R) * HID + k : 0), t > 0 ? 4 : 0);    // h_{-1} = 0
        cp_async4(&dh_s[slot][k], a.dh_head + (t < T ? m * HID + k : 0), t < T ? 4 : 0);          // no q at t = T
    };
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        if (p < a.TT) fetch(p, p);
        cp_async_commit();
    }
    float carry = 0.0f;
    // position of logical element e of [drp | dzp | dghn] inside the shifted-quarter layout
    auto dpos = [&](int e) { return e + (e / JC) * GT_PAD; };
    const int pos_r = dpos(k), pos_z = dpos(HID + k), pos_n = dpos(2 * HID + k);
    __syncthreads();

    for (int i0 = 0; i0 < a.TT; i0 += 2) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int i = i0 + p;
            if (i >= a.TT) break;
            const int buf = p;
            const int slot = i % PF;
            if (i + PDL_LEAD_STEPS == a.TT) pdl_trigger();
            cp_async_wait<PF - 1>();
            const float4 g4 = g4_s[slot][k];
            const float hp = hp_s[slot][k], dhh = dh_s[slot][k];
            if (i + PF < a.TT) fetch(i + PF, slot);
            cp_async_commit();
            const float4 *dp = reinterpret_cast<const float4 *>(dg_s[buf] + part * DSEG);
            unsigned long long s[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int u = 0; u < JC / 4; ++u) {
                const float4 d = dp[u];
                const unsigned long long dxy = pack2(d.x, d.y), dzw = pack2(d.z, d.w);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    s[c] = fma2(wT[c][2 * u], dxy, s[c]);
                    s[c] = fma2(wT[c][2 * u + 1], dzw, s[c]);
                }
            }
            // transposing reduction over the lane quad: lane `part` ends with the complete sum of column 4*kg + part
            float pc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { float lo, hi; unpack2(s[c], lo, hi); pc[c] = lo + hi; }
            const bool b0 = part & 1, b1 = part & 2;
            // stage 1 (xor 1): lanes with bit0 = 0 keep columns {0, 2}, the others {1, 3}
            const float k0 = b0 ? pc[1] : pc[0], k1 = b0 ? pc[3] : pc[2];
            const float g0 = b0 ? pc[0] : pc[1], g1 = b0 ? pc[2] : pc[3];
            const float q0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 1);
            const float q1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 1);
            // stage 2 (xor 2): lanes with bit1 = 0 keep the first of the pair, the others the second
            const float keep = b1 ? q1 : q0, give = b1 ? q0 : q1;
            const float dh = (keep + __shfl_xor_sync(0xffffffffu, give, 2)) + (carry + dhh);
            const float rr = g4.x, zz = g4.y, nn = g4.z, ghn = g4.w;
            const float dn = dh * (1.0f - zz);
            const float dz = dh * (hp - nn);
            const float dnp = dn * (1.0f - nn * nn);
            const float dzp = dz * zz * (1.0f - zz);
            const float drp = dnp * ghn * rr * (1.0f - rr);
            const float dghn = dnp * rr;
            carry = dh * zz;
            float *sm = dg_s[buf ^ 1];
            sm[pos_r] = drp; sm[pos_z] = dzp; sm[pos_n] = dghn;
            float *dg = a.d_g + ((int64_t)(a.TT - 1 - i) * a.R + row) * 4 * HID + k;
            dg[0] = drp; dg[HID] = dzp; dg[2 * HID] = dnp; dg[3 * HID] = dghn;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 128-thread variants: ONE chain per CTA spread over all four sub-partitions of its SM (one warp each), so that the FMA
// phase of a timestep is 48 FFMA2 per warp instead of 96 and an SM that hosts three chains (320 chains over 148 SMs)
// has three warps on EVERY sub-partition instead of two chains sharing two of them.
//     forward   U = 2 units x K/4 = 16 columns per lane   ->  4 LDS.128 + 48 FFMA2, two xor-shuffle stages
//     backward  U = 4 columns x 192/8 = 24 rows per lane  ->  6 LDS.128 + 48 FFMA2, three xor-shuffle stages
// The last shuffle stage is a butterfly: two lanes end with the same complete sum and do the gate math redundantly (no
// extra latency); the lower one of the pair ("owner") stores.  The per-step operands (gi / saved gates, h_{t-1}, head
// seed) arrive through a CTA-wide cp.async ring of 16-byte pieces: step s + PF - 1 is requested at step s into the slot
// that step s - 1 used (everybody is past that step's barrier), and the issuing threads wait for step s + 1's group
// before the barrier that ends step s, which publishes it.
// ---------------------------------------------------------------------------------------------------------------------
template <int DBG = 0>
__global__ void __launch_bounds__(128, 3) k_gru_fwd8(GruFwdArgs a) {
    constexpr int PF = 8;
    constexpr int KC = HID / 4;            // columns per lane
    constexpr int SEG = KC + GT_PAD;
    __shared__ __align__(16) float h_s[2][4 * SEG];
    __shared__ __align__(16) float st_s[PF][G3];
    __shared__ float pdl_anchor_s[128];
    const int i = threadIdx.x, net = blockIdx.y, row = blockIdx.x;
    const int grp = i >> 2, part = i & 3;
    const int unit = 2 * grp + (part & 1);           // the unit this lane finishes (lanes part and part ^ 2 both do)
    const bool owner = part < 2;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const float *P = net ? a.params[1] : a.params[0];
    const float *gi = net ? a.gi[1] : a.gi[0];
    float *hout = net ? a.hout[1] : a.hout[0];
    float4 *gates = net == 0 ? reinterpret_cast<float4 *>(a.gates) : nullptr;
    const int t0 = a.t0, t1 = a.t1;

    float anchor = 0.0f;
    unsigned long long w[2][3][KC / 2];    // packed pairs of W_hh[g*64 + 2*grp + u][part*16 + 2j, +1]
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float4 *wr = reinterpret_cast<const float4 *>(P + L.w_hh + (int64_t)(g * HID + 2 * grp + u) * HID + part * KC);
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                const float4 v = __ldg(wr + q);
                w[u][g][2 * q] = pack2(v.x, v.y); w[u][g][2 * q + 1] = pack2(v.z, v.w);
                anchor += v.x;
            }
        }
    const float b_r = __ldg(P + L.b_hh + unit), b_z = __ldg(P + L.b_hh + HID + unit), b_n = __ldg(P + L.b_hh + 2 * HID + unit);
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[i]) = anchor + b_r + b_z + b_n;   // pins the loads above the wait (see k_gru_fwd7)
    pdl_wait();
    float hprev = t0 > 0 ? hout[((int64_t)(t0 - 1) * a.R + row) * HID + unit] : 0.0f;   // init_hidden: zeros
    const int hpos = unit + (unit / KC) * GT_PAD;
    if (owner) { h_s[0][hpos] = hprev; h_s[1][hpos] = 0.0f; }
    const int64_t tstride = (int64_t)a.R * G3;
    const float *p_g = gi + (int64_t)row * G3 + 4 * i;          // threads 0..47: 16-byte piece i of the 768-byte gi row
    const bool loader = i < G3 / 4;
#pragma unroll
    for (int p = 0; p < PF - 1; ++p) {
        if (loader && t0 + p < t1) cp_async16(&st_s[p][4 * i], p_g + (int64_t)(t0 + p) * tstride);
        cp_async_commit();
    }
    cp_async_wait<PF - 2>();
    __syncthreads();

    for (int tt = t0; tt < t1; tt += 2) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int t = tt + p;
            if (t >= t1) break;
            const int buf = p;                   // == (t - t0) & 1
            const int s = t - t0;
            const int slot = s % PF;
            if (t + PDL_LEAD_STEPS == t1) pdl_trigger();
            const float g_r = st_s[slot][unit], g_z = st_s[slot][HID + unit], g_n = st_s[slot][2 * HID + unit];
            if (loader && t + PF - 1 < t1) cp_async16(&st_s[(s + PF - 1) % PF][4 * i], p_g + (int64_t)(t + PF - 1) * tstride);
            cp_async_commit();
            const float4 *hp = reinterpret_cast<const float4 *>(h_s[buf] + part * SEG);
            unsigned long long acc[2][3];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 3; ++g) acc[u][g] = 0ull;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                const float4 hv = hp[q];
                const unsigned long long hxy = pack2(hv.x, hv.y), hzw = pack2(hv.z, hv.w);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[u][g] = fma2(w[u][g][2 * q], hxy, acc[u][g]);
                        acc[u][g] = fma2(w[u][g][2 * q + 1], hzw, acc[u][g]);
                    }
            }
            // transposing reduction over the lane quad: xor 1 hands each lane its unit, xor 2 completes the sum
            float x[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float a0, a1, c0, c1;
                unpack2(acc[0][g], a0, a1); unpack2(acc[1][g], c0, c1);
                const float p0 = a0 + a1, p1 = c0 + c1;
                const float mine = (part & 1) ? p1 : p0, other = (part & 1) ? p0 : p1;
                const float y = mine + __shfl_xor_sync(0xffffffffu, other, 1);
                x[g] = y + __shfl_xor_sync(0xffffffffu, y, 2);
            }
            const float xr = x[0] + (g_r + b_r), xz = x[1] + (g_z + b_z), ghn = x[2] + b_n;
            const float rr = sigmoid_mufu(xr), zz = sigmoid_mufu(xz);
            const float nn = tanh_mufu(g_n + rr * ghn);
            const float hn = nn + zz * (hprev - nn);
            hprev = hn;
            if (owner) {
                h_s[buf ^ 1][hpos] = hn;
                if (!(DBG & 1)) {
                    const int64_t m = (int64_t)t * a.R + row;
                    hout[m * HID + unit] = hn;
                    if (gates) gates[m * HID + unit] = make_float4(rr, zz, nn, ghn);
                }
            }
            cp_async_wait<PF - 2>();             // step t + 1's operands have landed (issuing threads); the barrier publishes them
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(128, 3) k_gru_bwd8(GruBwdArgs a) {
    constexpr int PF = 8;
    constexpr int JC = G3 / 8;             // rows j of W_hh per lane (24)
    constexpr int DSEG = JC + GT_PAD;      // padded eighth of the d(gates) operand
    __shared__ __align__(16) float dg_s[2][8 * DSEG];   // d gi_r | d gi_z | d gh_n of the previous step, in eight shifted pieces
    __shared__ __align__(16) float4 g4_s[PF][HID];
    __shared__ __align__(16) float hp_s[PF][HID];
    __shared__ __align__(16) float dh_s[PF][HID];
    __shared__ float pdl_anchor_s[128];
    const int k = threadIdx.x, row = blockIdx.x;
    const int grp = k >> 3, part = k & 7;
    const int col = 4 * grp + (part & 3);            // the column this lane finishes (lanes part and part ^ 4 both do)
    const bool owner = part < 4;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const int T = a.TT - 1;

    float anchor = 0.0f;
    for (int idx = k; idx < 2 * 8 * DSEG; idx += 128) (&dg_s[0][0])[idx] = 0.0f;
    unsigned long long wT[4][JC / 2];      // packed pairs (W_hh[24*part + 2j][4grp + c], W_hh[24*part + 2j + 1][4grp + c])
#pragma unroll
    for (int j = 0; j < JC / 2; ++j) {
        const int jj = part * JC + 2 * j;
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)jj * HID + 4 * grp));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)(jj + 1) * HID + 4 * grp));
        wT[0][j] = pack2(w0.x, w1.x); wT[1][j] = pack2(w0.y, w1.y); wT[2][j] = pack2(w0.z, w1.z); wT[3][j] = pack2(w0.w, w1.w);
        anchor += w0.x + w1.y;
    }
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[k]) = anchor;   // pins the loads above the wait (see k_gru_fwd7)

    pdl_wait();                                  // W_hh is a step constant; gates / h / dh_head come from the predecessors
    const float4 *gates4 = reinterpret_cast<const float4 *>(a.gates);
    auto fetch = [&](int i, int slot) {          // step i <-> t = TT-1-i; threads 0..63 gates, 64..79 h_{t-1}, 80..95 head seed
        const int t = a.TT - 1 - i;
        const int64_t m = (int64_t)t * a.R + row;
        if (k < 64) cp_async16(&g4_s[slot][k], gates4 + m * HID + k);
        else if (k < 80) cp_async16(&hp_s[slot][4 * (k - 64)], a.hout + (t > 0 ? (m - a.R) * HID + 4 * (k - 64) : 0), t > 0 ? 16 : 0);   // h_{-1} = 0
        else if (k < 96) cp_async16(&dh_s[slot][4 * (k - 80)], a.dh_head + (t < T ? m * HID + 4 * (k - 80) : 0), t < T ? 16 : 0);        // no q at t = T
    };
#pragma unroll
    for (int p = 0; p < PF - 1; ++p) {
        if (p < a.TT) fetch(p, p);
        cp_async_commit();
    }
    float carry = 0.0f;
    // position of logical element e of [drp | dzp | dghn] inside the shifted-piece layout
    auto dpos = [&](int e) { return e + (e / JC) * GT_PAD; };
    const int pos_r = dpos(col), pos_z = dpos(HID + col), pos_n = dpos(2 * HID + col);
    cp_async_wait<PF - 2>();
    __syncthreads();

    for (int i0 = 0; i0 < a.TT; i0 += 2) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int i = i0 + p;
            if (i >= a.TT) break;
            const int buf = p;
            const int slot = i % PF;
            if (i + PDL_LEAD_STEPS == a.TT) pdl_trigger();
            const float4 g4 = g4_s[slot][col];
            const float hp = hp_s[slot][col], dhh = dh_s[slot][col];
            if (i + PF - 1 < a.TT) fetch(i + PF - 1, (i + PF - 1) % PF);
            cp_async_commit();
            const float4 *dp = reinterpret_cast<const float4 *>(dg_s[buf] + part * DSEG);
            unsigned long long s[4][2];
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c][0] = s[c][1] = 0ull;
#pragma unroll
            for (int u = 0; u < JC / 4; ++u) {
                const float4 d = dp[u];
                const unsigned long long dxy = pack2(d.x, d.y), dzw = pack2(d.z, d.w);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    s[c][0] = fma2(wT[c][2 * u], dxy, s[c][0]);
                    s[c][1] = fma2(wT[c][2 * u + 1], dzw, s[c][1]);
                }
            }
            // transposing reduction over the lane octet: xor 1 and xor 2 hand each lane its column, xor 4 completes the sum
            float pc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { float l0, h0, l1, h1; unpack2(s[c][0], l0, h0); unpack2(s[c][1], l1, h1); pc[c] = (l0 + h0) + (l1 + h1); }
            const bool b0 = part & 1, b1 = part & 2;
            const float k0 = b0 ? pc[1] : pc[0], k1 = b0 ? pc[3] : pc[2];
            const float g0 = b0 ? pc[0] : pc[1], g1 = b0 ? pc[2] : pc[3];
            const float q0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 1);
            const float q1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 1);
            const float keep = b1 ? q1 : q0, give = b1 ? q0 : q1;
            const float v2 = keep + __shfl_xor_sync(0xffffffffu, give, 2);
            const float dh = (v2 + __shfl_xor_sync(0xffffffffu, v2, 4)) + (carry + dhh);
            const float rr = g4.x, zz = g4.y, nn = g4.z, ghn = g4.w;
            const float dn = dh * (1.0f - zz);
            const float dz = dh * (hp - nn);
            const float dnp = dn * (1.0f - nn * nn);
            const float dzp = dz * zz * (1.0f - zz);
            const float drp = dnp * ghn * rr * (1.0f - rr);
            const float dghn = dnp * rr;
            carry = dh * zz;
            if (owner) {
                float *sm = dg_s[buf ^ 1];
                sm[pos_r] = drp; sm[pos_z] = dzp; sm[pos_n] = dghn;
                float *dg = a.d_g + ((int64_t)(a.TT - 1 - i) * a.R + row) * 4 * HID + col;
                dg[0] = drp; dg[HID] = dzp; dg[2 * HID] = dnp; dg[3 * HID] = dghn;
            }
            cp_async_wait<PF - 2>();
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 9 = the 64-thread layout of variant 7 with the per-step instruction count cut down.  The source-level profile
// of variant 7 (profiles/r02_gru7_source_summary.txt) shows a single warp per sub-partition issuing one instruction every
// ~4.4 cycles, FMA or not: 96 FFMA2 and ~100 other instructions per step, ~40 of them ring-slot / address / boundary
// arithmetic.  Here the time loop is unrolled over the PF ring slots (slot and h buffer are compile-time constants, so
// every shared-memory access has an immediate offset), the global pointers advance by a constant per step, and the
// steady-state loop carries no boundary predicates (the last <= 2 PF - 1 steps run in a guarded tail).
// ---------------------------------------------------------------------------------------------------------------------
template <int DBG = 0, int PACKED = 1>
__global__ void __launch_bounds__(64, 4) k_gru_fwd9(GruFwdArgs a) {
    constexpr int PF = 8;
    constexpr int KC = HID / 2;            // columns per lane
    constexpr int HROW = HID + GT_PAD;
    __shared__ __align__(16) float h_s[2][HROW];
    __shared__ float st_s[PF][3][HID];
    __shared__ float pdl_anchor_s[HID];
    const int i = threadIdx.x;
    const int ug = i >> 1, part = i & 1;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    // Work of this CTA as a range [lo, hi) of the chain-major step sequence (chain c = net * R + row owns [c TT, (c + 1) TT)).
    // Plain launches (grid R x nets): the CTA's chain, steps [t0, t1).  Balanced launches (grid workers x 1): bal_D steps that
    // may straddle chains -- 2 R chains on 2 SMs-worth of two-warp CTAs run a step in ~560 cycles with every warp alone on its
    // sub-partition, but ~850 once a third chain lands on an SM (measured: tools/gru_bench.cu), so 320 chains are run as 296
    // equal workers instead of leaving 24 SMs with three chains.
    int lo, hi;
    if (a.bal_D > 0) {
        lo = a.bal_D * (int)blockIdx.x;
        hi = min(lo + a.bal_D, a.bal_chains * a.TT);
        if (lo >= hi) return;
    } else {
        lo = ((int)blockIdx.y * a.R + (int)blockIdx.x) * a.TT + a.t0;
        hi = lo + (a.t1 - a.t0);
    }
    int cur_net = -1;
    const float *gi = nullptr;
    float *hout = nullptr;
    bool save_gates = false;
    unsigned long long w[2][3][KC / 2];    // packed pairs of W_hh[g*64 + 2*ug + u][part*32 + 2j, +1]
    float b_r = 0.f, b_z = 0.f, b_n = 0.f;
    const int64_t hstride = (int64_t)a.R * HID, tstride = (int64_t)a.R * G3;
    const int hpos = i + (i >= KC ? GT_PAD : 0);
    const float4 *hp0 = reinterpret_cast<const float4 *>(h_s[0] + part * (KC + GT_PAD));
    const float4 *hp1 = reinterpret_cast<const float4 *>(h_s[1] + part * (KC + GT_PAD));
    float *hdst = nullptr;                 // h of the current step
    float4 *gdst = nullptr;
    const float *gsrc = nullptr;           // gi row of the next step to request
    float hprev = 0.0f;
    int n = 0;

    // one timestep; SLOT / BUF are compile-time, `more` = the step PF ahead exists (always true in the steady-state loop)
    auto step = [&](auto slot_c, auto buf_c, bool more) {
        constexpr int SLOT = decltype(slot_c)::value, BUF = decltype(buf_c)::value;
        cp_async_wait<PF - 1>();             // this thread's group of this step has landed
        const float g_r = st_s[SLOT][0][i], g_z = st_s[SLOT][1][i], g_n = st_s[SLOT][2][i];
        if (more) {
#pragma unroll
            for (int g = 0; g < 3; ++g) cp_async4(&st_s[SLOT][g][i], gsrc + g * HID);
        }
        cp_async_commit();
        gsrc += tstride;
        const float4 *hp = BUF ? hp1 : hp0;
        float4 hv[KC / 4];
#pragma unroll
        for (int q = 0; q < KC / 4; ++q) hv[q] = hp[q];
        float x[3];
        if (PACKED) {
            unsigned long long s[2][3];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 3; ++g) s[u][g] = 0ull;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                const unsigned long long hxy = pack2(hv[q].x, hv[q].y), hzw = pack2(hv[q].z, hv[q].w);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int g = 0; g < 3; ++g) s[u][g] = fma2(w[u][g][2 * q], hxy, s[u][g]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int g = 0; g < 3; ++g) s[u][g] = fma2(w[u][g][2 * q + 1], hzw, s[u][g]);
            }
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float a0, a1, c0, c1;
                unpack2(s[0][g], a0, a1); unpack2(s[1][g], c0, c1);
                const float p0 = a0 + a1, p1 = c0 + c1;
                const float mine = part ? p1 : p0, other = part ? p0 : p1;
                x[g] = mine + __shfl_xor_sync(0xffffffffu, other, 1);
            }
        } else {
            // plain FFMA: 192 single FMAs over six accumulators (an FFMA2 with three 64-bit register operands issued every
            // ~4 cycles from a lone warp in this kernel; FFMA issues every ~1.1, tools/ffma_probe.cu)
            float sa[2][3];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 3; ++g) sa[u][g] = 0.0f;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                const float hh[4] = {hv[q].x, hv[q].y, hv[q].z, hv[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            float wl, wh;
                            unpack2(w[u][g][2 * q + (e >> 1)], wl, wh);
                            sa[u][g] = fmaf((e & 1) ? wh : wl, hh[e], sa[u][g]);
                        }
            }
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float mine = part ? sa[1][g] : sa[0][g], other = part ? sa[0][g] : sa[1][g];
                x[g] = mine + __shfl_xor_sync(0xffffffffu, other, 1);
            }
        }
        const float xr = x[0] + (g_r + b_r), xz = x[1] + (g_z + b_z), ghn = x[2] + b_n;
        const float rr = sigmoid_mufu(xr), zz = sigmoid_mufu(xz);
        const float nn = tanh_mufu(g_n + rr * ghn);
        const float hn = nn + zz * (hprev - nn);
        hprev = hn;
        h_s[BUF ^ 1][hpos] = hn;
        if (!(DBG & 1)) {
            *hdst = hn;
            if (save_gates) *gdst = make_float4(rr, zz, nn, ghn);
        }
        hdst += hstride; gdst += hstride;
        __syncthreads();
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
#pragma unroll 1
    for (int pend = hi; pend > lo;) {                 // pieces of [lo, hi), last first
        const int chain = (pend - 1) / a.TT;
        const int pbeg = max(lo, chain * a.TT);
        const int tb = pbeg - chain * a.TT, te = pend - chain * a.TT;
        const int net = chain >= a.R ? 1 : 0, row = chain - net * a.R;
        const bool last_piece = pbeg == lo;
        pend = pbeg;
        if (net != cur_net) {                             // W_hh / b_hh of this net into registers
            cur_net = net;
            const float *P = net ? a.params[1] : a.params[0];
            gi = net ? a.gi[1] : a.gi[0];
            hout = net ? a.hout[1] : a.hout[0];
            save_gates = net == 0;
            float anchor = 0.0f;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float4 *wr = reinterpret_cast<const float4 *>(P + L.w_hh + (int64_t)(g * HID + 2 * ug + u) * HID + part * KC);
#pragma unroll
                    for (int q = 0; q < KC / 4; ++q) {
                        const float4 v = __ldg(wr + q);
                        w[u][g][2 * q] = pack2(v.x, v.y); w[u][g][2 * q + 1] = pack2(v.z, v.w);
                        anchor += v.x;
                    }
                }
            b_r = __ldg(P + L.b_hh + i); b_z = __ldg(P + L.b_hh + HID + i); b_n = __ldg(P + L.b_hh + 2 * HID + i);
            *reinterpret_cast<volatile float *>(&pdl_anchor_s[i]) = anchor + b_r + b_z + b_n;   // pins the loads above the wait (see k_gru_fwd7)
            pdl_wait();                                   // (returns at once after the first time)
        }
        n = te - tb;
        hdst = hout + ((int64_t)tb * a.R + row) * HID + i;
        gdst = reinterpret_cast<float4 *>(a.gates) + ((int64_t)tb * a.R + row) * HID + i;
        gsrc = gi + ((int64_t)tb * a.R + row) * G3 + i;
#pragma unroll
        for (int p = 0; p < PF; ++p) {                    // the ring does not depend on the chain's head: request it before the wait
            if (p < n) {
#pragma unroll
                for (int g = 0; g < 3; ++g) cp_async4(&st_s[p][g][i], gsrc + g * HID);
            }
            cp_async_commit();
            gsrc += tstride;
        }
        if (a.bal_D > 0 && tb > 0) {                      // the head of this chain belongs to the next worker: wait for it
            if (i == 0) {
                volatile int *f = a.chain_flags + chain;
                while (*f == 0) { }
                *f = 0;                                   // single consumer: leave the flag clear for the next launch
                __threadfence();
            }
            __syncthreads();
        }
        hprev = tb > 0 ? __ldcg(hdst - hstride) : 0.0f;   // init_hidden: zeros; (written by another SM in balanced mode: L2)
        h_s[0][hpos] = hprev; h_s[1][hpos] = 0.0f;
        __syncthreads();
        int sdone = 0;
        for (; sdone + 2 * PF <= n; sdone += PF) {       // steady state: every prefetch target exists
            step(std::integral_constant<int, 0>{}, I0{}, true); step(std::integral_constant<int, 1>{}, I1{}, true);
            step(std::integral_constant<int, 2>{}, I0{}, true); step(std::integral_constant<int, 3>{}, I1{}, true);
            step(std::integral_constant<int, 4>{}, I0{}, true); step(std::integral_constant<int, 5>{}, I1{}, true);
            step(std::integral_constant<int, 6>{}, I0{}, true); step(std::integral_constant<int, 7>{}, I1{}, true);
        }
        if (last_piece) pdl_trigger();                   // at most 2 PF - 1 steps left: the dependent kernel may start its prologue
        // guarded tail (sdone is a multiple of PF, so slots and buffers line up with the unrolled order)
#define GRU9_TAIL(S, B) if (sdone + S < n) step(std::integral_constant<int, S>{}, B{}, sdone + S + PF < n)
        for (; sdone < n; sdone += PF) {
            GRU9_TAIL(0, I0); GRU9_TAIL(1, I1); GRU9_TAIL(2, I0); GRU9_TAIL(3, I1);
            GRU9_TAIL(4, I0); GRU9_TAIL(5, I1); GRU9_TAIL(6, I0); GRU9_TAIL(7, I1);
        }
#undef GRU9_TAIL
        cp_async_wait<0>();                               // (empty groups of the tail steps)
        if (a.bal_D > 0 && te < a.TT) {                   // head piece done: its last h row is what the tail's worker starts from
            __syncthreads();                              // (every thread's stores of this piece precede thread 0's fence)
            if (i == 0) {
                __threadfence();
                *reinterpret_cast<volatile int *>(a.chain_flags + chain) = 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 11 (forward, small row counts): ONE chain per 256-thread CTA, thread = (unit, quarter of K).  At B = 32 an SM
// hosts two or three chains and the 64-thread layouts leave every sub-partition with one or two warps that each issue 96
// half-rate FFMA2 and then walk the serial gate tail alone (~870 cycles per step, the same whether one or two such warps
// share a sub-partition).  Here a thread issues 24 FFMA2 (3 gates x 16 columns), the four partial sums of a unit meet in
// a two-stage xor butterfly (all four lanes end with the same bits and do the gate math redundantly), and an SM runs 16
// to 24 warps whose FMA phases and tails interleave.  Ring / barrier protocol of variant 8 (CTA-wide cp.async ring of
// 16-byte pieces, step s + PF - 1 requested at step s into the slot step s - 1 used), time loop unrolled over the ring
// slots like variant 9.
// ---------------------------------------------------------------------------------------------------------------------
template <int DBG = 0>
__global__ void __launch_bounds__(256, 3) k_gru_fwd11(GruFwdArgs a) {
    constexpr int PF = 8;
    constexpr int KC = HID / 4;            // columns per lane
    constexpr int SEG = KC + GT_PAD;       // 80-byte pitch: the four quarters of one LDS.128 fall into disjoint banks
    __shared__ __align__(16) float h_s[2][4 * SEG];
    __shared__ __align__(16) float st_s[PF][G3];
    __shared__ float pdl_anchor_s[256];
    const int i = threadIdx.x, net = blockIdx.y, row = blockIdx.x;
    const int unit = i >> 2, part = i & 3;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const float *P = net ? a.params[1] : a.params[0];
    const float *gi = net ? a.gi[1] : a.gi[0];
    float *hout = net ? a.hout[1] : a.hout[0];
    const bool h_owner = part == 0, g_owner = part == 1 && net == 0;
    const int t0 = a.t0, n = a.t1 - a.t0;

    float anchor = 0.0f;
    unsigned long long w[3][KC / 2];       // packed pairs of W_hh[g*64 + unit][part*16 + 2j, +1]
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const float4 *wr = reinterpret_cast<const float4 *>(P + L.w_hh + (int64_t)(g * HID + unit) * HID + part * KC);
#pragma unroll
        for (int q = 0; q < KC / 4; ++q) {
            const float4 v = __ldg(wr + q);
            w[g][2 * q] = pack2(v.x, v.y); w[g][2 * q + 1] = pack2(v.z, v.w);
            anchor += v.x;
        }
    }
    const float b_r = __ldg(P + L.b_hh + unit), b_z = __ldg(P + L.b_hh + HID + unit), b_n = __ldg(P + L.b_hh + 2 * HID + unit);
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[i]) = anchor + b_r + b_z + b_n;   // pins the loads above the wait (see k_gru_fwd7)
    pdl_wait();
    const int64_t hstride = (int64_t)a.R * HID, tstride = (int64_t)a.R * G3;
    float *hdst = hout + ((int64_t)t0 * a.R + row) * HID + unit;              // h of the current step
    float4 *gdst = reinterpret_cast<float4 *>(a.gates) + ((int64_t)t0 * a.R + row) * HID + unit;
    float hprev = t0 > 0 ? *(hdst - hstride) : 0.0f;                           // init_hidden: zeros
    const int hpos = unit + (unit / KC) * GT_PAD;
    if (h_owner) { h_s[0][hpos] = hprev; h_s[1][hpos] = 0.0f; }
    const bool loader = i < G3 / 4;                                            // 48 16-byte pieces of a 768-byte gi row
    const float *gsrc = gi + ((int64_t)t0 * a.R + row) * G3 + 4 * (loader ? i : 0);
#pragma unroll
    for (int p = 0; p < PF - 1; ++p) {
        if (loader && p < n) cp_async16(&st_s[p][4 * i], gsrc);
        cp_async_commit();
        gsrc += tstride;
    }
    cp_async_wait<PF - 2>();
    __syncthreads();
    const float4 *hp0 = reinterpret_cast<const float4 *>(h_s[0] + part * SEG);
    const float4 *hp1 = reinterpret_cast<const float4 *>(h_s[1] + part * SEG);

    // one timestep; SLOT / BUF are compile-time, `more` = step s + PF - 1 exists (always true in the steady-state loop)
    auto step = [&](auto slot_c, auto buf_c, bool more) {
        constexpr int SLOT = decltype(slot_c)::value, BUF = decltype(buf_c)::value;
        const float g_r = st_s[SLOT][unit], g_z = st_s[SLOT][HID + unit], g_n = st_s[SLOT][2 * HID + unit];
        if (loader && more) cp_async16(&st_s[(SLOT + PF - 1) % PF][4 * i], gsrc);
        cp_async_commit();
        gsrc += tstride;
        const float4 *hp = BUF ? hp1 : hp0;
        float4 hv[KC / 4];
#pragma unroll
        for (int q = 0; q < KC / 4; ++q) hv[q] = hp[q];
        unsigned long long s[3] = {0ull, 0ull, 0ull};
#pragma unroll
        for (int q = 0; q < KC / 4; ++q) {
            const unsigned long long hxy = pack2(hv[q].x, hv[q].y), hzw = pack2(hv[q].z, hv[q].w);
#pragma unroll
            for (int g = 0; g < 3; ++g) s[g] = fma2(w[g][2 * q], hxy, s[g]);
#pragma unroll
            for (int g = 0; g < 3; ++g) s[g] = fma2(w[g][2 * q + 1], hzw, s[g]);
        }
        float x[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float a0, a1;
            unpack2(s[g], a0, a1);
            float y = a0 + a1;
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            x[g] = y;
        }
        const float xr = x[0] + (g_r + b_r), xz = x[1] + (g_z + b_z), ghn = x[2] + b_n;
        const float rr = sigmoid_mufu(xr), zz = sigmoid_mufu(xz);
        const float nn = tanh_mufu(g_n + rr * ghn);
        const float hn = nn + zz * (hprev - nn);
        hprev = hn;
        if (h_owner) {
            h_s[BUF ^ 1][hpos] = hn;
            if (!(DBG & 1)) *hdst = hn;
        }
        if (g_owner && !(DBG & 1)) *gdst = make_float4(rr, zz, nn, ghn);
        hdst += hstride; gdst += hstride;
        cp_async_wait<PF - 2>();             // step s + 1's operands have landed (issuing threads); the barrier publishes them
        __syncthreads();
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
    int sdone = 0;
    for (; sdone + 2 * PF <= n; sdone += PF) {       // steady state: every prefetch target exists
        step(std::integral_constant<int, 0>{}, I0{}, true); step(std::integral_constant<int, 1>{}, I1{}, true);
        step(std::integral_constant<int, 2>{}, I0{}, true); step(std::integral_constant<int, 3>{}, I1{}, true);
        step(std::integral_constant<int, 4>{}, I0{}, true); step(std::integral_constant<int, 5>{}, I1{}, true);
        step(std::integral_constant<int, 6>{}, I0{}, true); step(std::integral_constant<int, 7>{}, I1{}, true);
    }
    pdl_trigger();                                   // at most 2 PF - 1 steps left: the dependent kernel may start its prologue
#define GRU11_TAIL(S, B) if (sdone + S < n) step(std::integral_constant<int, S>{}, B{}, sdone + S + PF - 1 < n)
    for (; sdone < n; sdone += PF) {
        GRU11_TAIL(0, I0); GRU11_TAIL(1, I1); GRU11_TAIL(2, I0); GRU11_TAIL(3, I1);
        GRU11_TAIL(4, I0); GRU11_TAIL(5, I1); GRU11_TAIL(6, I0); GRU11_TAIL(7, I1);
    }
#undef GRU11_TAIL
}

__global__ void __launch_bounds__(64, 4) k_gru_bwd9(GruBwdArgs a) {
    constexpr int PF = 8;
    constexpr int JC = G3 / 4;             // rows j of W_hh per lane (48)
    constexpr int DSEG = JC + GT_PAD;      // padded quarter of the d(gates) operand
    __shared__ __align__(16) float dg_s[2][4 * DSEG];   // d gi_r | d gi_z | d gh_n of the previous step, in four shifted quarters
    __shared__ __align__(16) float4 g4_s[PF][HID];
    __shared__ float hp_s[PF][HID], dh_s[PF][HID];
    __shared__ float pdl_anchor_s[HID];
    const int k = threadIdx.x, row = blockIdx.x;
    const int kg = k >> 2, part = k & 3;
    const AgentLayout L = agent_layout(a.d_in, a.n_actions);
    const int n = a.TT;

    float anchor = 0.0f;
    for (int idx = k; idx < 2 * 4 * DSEG; idx += HID) (&dg_s[0][0])[idx] = 0.0f;
    unsigned long long wT[4][JC / 2];      // packed pairs (W_hh[48*part + 2j][4kg + c], W_hh[48*part + 2j + 1][4kg + c])
#pragma unroll
    for (int j = 0; j < JC / 2; ++j) {
        const int jj = part * JC + 2 * j;
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)jj * HID + 4 * kg));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(a.params + L.w_hh + (int64_t)(jj + 1) * HID + 4 * kg));
        wT[0][j] = pack2(w0.x, w1.x); wT[1][j] = pack2(w0.y, w1.y); wT[2][j] = pack2(w0.z, w1.z); wT[3][j] = pack2(w0.w, w1.w);
        anchor += w0.x + w1.y;
    }
    *reinterpret_cast<volatile float *>(&pdl_anchor_s[k]) = anchor;   // pins the loads above the wait (see k_gru_fwd7)

    pdl_wait();                                  // W_hh is a step constant; gates / h / dh_head come from the predecessors
    // step s <-> t = TT-1-s; every pointer is that of the NEXT step to request and walks backwards by one timestep
    const int64_t hstride = (int64_t)a.R * HID;
    const int64_t m_last = (int64_t)(a.TT - 1) * a.R + row;
    const float4 *gsrc = reinterpret_cast<const float4 *>(a.gates) + m_last * HID + k;
    const float *hsrc = a.hout + (m_last - a.R) * HID + k;          // h_{t-1}; not dereferenced at t = 0
    const float *dsrc = a.dh_head + m_last * HID + k;               // not dereferenced at t = TT-1 (no q there)
    float *dgdst = a.d_g + m_last * 4 * HID + k;
    // generic request of step s into `slot` (prologue and tail: with the sequence-boundary guards)
    auto fetch_guarded = [&](int s, int slot) {
        const int t = n - 1 - s;
        cp_async16(&g4_s[slot][k], gsrc);
        cp_async4(&hp_s[slot][k], t > 0 ? hsrc : a.hout, t > 0 ? 4 : 0);          // h_{-1} = 0
        cp_async4(&dh_s[slot][k], t < n - 1 ? dsrc : a.dh_head, t < n - 1 ? 4 : 0);
    };
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        if (p < n) fetch_guarded(p, p);
        cp_async_commit();
        gsrc -= hstride; hsrc -= hstride; dsrc -= hstride;
    }
    float carry = 0.0f;
    auto dpos = [&](int e) { return e + (e / JC) * GT_PAD; };
    const int pos_r = dpos(k), pos_z = dpos(HID + k), pos_n = dpos(2 * HID + k);
    __syncthreads();
    const float4 *dp0 = reinterpret_cast<const float4 *>(dg_s[0] + part * DSEG);
    const float4 *dp1 = reinterpret_cast<const float4 *>(dg_s[1] + part * DSEG);
    const bool b0 = part & 1, b1 = part & 2;

    // one step; MODE 0: steady state (the step PF ahead exists and is away from both sequence ends), 1: guarded request, 2: none
    auto step = [&](auto slot_c, auto buf_c, int mode, int s) {
        constexpr int SLOT = decltype(slot_c)::value, BUF = decltype(buf_c)::value;
        cp_async_wait<PF - 1>();
        const float4 g4 = g4_s[SLOT][k];
        const float hp = hp_s[SLOT][k], dhh = dh_s[SLOT][k];
        if (mode == 0) {
            cp_async16(&g4_s[SLOT][k], gsrc);
            cp_async4(&hp_s[SLOT][k], hsrc);
            cp_async4(&dh_s[SLOT][k], dsrc);
        } else if (mode == 1) fetch_guarded(s + PF, SLOT);
        cp_async_commit();
        gsrc -= hstride; hsrc -= hstride; dsrc -= hstride;
        const float4 *dp = BUF ? dp1 : dp0;
        float4 d[JC / 4];
#pragma unroll
        for (int u = 0; u < JC / 4; ++u) d[u] = dp[u];
        unsigned long long acc[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = 0ull;
#pragma unroll
        for (int u = 0; u < JC / 4; ++u) {
            const unsigned long long dxy = pack2(d[u].x, d[u].y), dzw = pack2(d[u].z, d[u].w);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][0] = fma2(wT[c][2 * u], dxy, acc[c][0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][1] = fma2(wT[c][2 * u + 1], dzw, acc[c][1]);
        }
        float pc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { float l0, h0, l1, h1; unpack2(acc[c][0], l0, h0); unpack2(acc[c][1], l1, h1); pc[c] = (l0 + h0) + (l1 + h1); }
        const float k0 = b0 ? pc[1] : pc[0], k1 = b0 ? pc[3] : pc[2];
        const float g0 = b0 ? pc[0] : pc[1], g1 = b0 ? pc[2] : pc[3];
        const float q0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 1);
        const float q1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 1);
        const float keep = b1 ? q1 : q0, give = b1 ? q0 : q1;
        const float dh = (keep + __shfl_xor_sync(0xffffffffu, give, 2)) + (carry + dhh);
        const float rr = g4.x, zz = g4.y, nn = g4.z, ghn = g4.w;
        const float dn = dh * (1.0f - zz);
        const float dz = dh * (hp - nn);
        const float dnp = dn * (1.0f - nn * nn);
        const float dzp = dz * zz * (1.0f - zz);
        const float drp = dnp * ghn * rr * (1.0f - rr);
        const float dghn = dnp * rr;
        carry = dh * zz;
        float *sm = dg_s[BUF ^ 1];
        sm[pos_r] = drp; sm[pos_z] = dzp; sm[pos_n] = dghn;
        dgdst[0] = drp; dgdst[HID] = dzp; dgdst[2 * HID] = dnp; dgdst[3 * HID] = dghn;
        dgdst -= 4 * hstride;
        __syncthreads();
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
    int sdone = 0;
    for (; sdone + 2 * PF + 1 <= n; sdone += PF) {   // steady state: steps s + PF stay >= 1 timestep away from t = 0
        step(std::integral_constant<int, 0>{}, I0{}, 0, 0); step(std::integral_constant<int, 1>{}, I1{}, 0, 0);
        step(std::integral_constant<int, 2>{}, I0{}, 0, 0); step(std::integral_constant<int, 3>{}, I1{}, 0, 0);
        step(std::integral_constant<int, 4>{}, I0{}, 0, 0); step(std::integral_constant<int, 5>{}, I1{}, 0, 0);
        step(std::integral_constant<int, 6>{}, I0{}, 0, 0); step(std::integral_constant<int, 7>{}, I1{}, 0, 0);
    }
    pdl_trigger();
#define GRU9_TAIL(S, B) if (sdone + S < n) step(std::integral_constant<int, S>{}, B{}, sdone + S + PF < n ? 1 : 2, sdone + S)
    for (; sdone < n; sdone += PF) {
        GRU9_TAIL(0, I0); GRU9_TAIL(1, I1); GRU9_TAIL(2, I0); GRU9_TAIL(3, I1);
        GRU9_TAIL(4, I0); GRU9_TAIL(5, I1); GRU9_TAIL(6, I0); GRU9_TAIL(7, I1);
    }
#undef GRU9_TAIL
}
