// K9: fused single-timestep agent forward + epsilon-greedy action selection.
//
// Reference: BasicMAC.select_actions / forward / _build_inputs (marl/controllers/basic_controller.py:29-54,80-92),
// DRQNAgentNetwork.forward (marl/modules/agents/drqn_agent.py:29-35) and EpsilonGreedyActionSelector.select
// (marl/components/action_selectors.py:44-62) -- ~25 ATen launches per env step in the reference, ONE here.
//
// Layout: a CTA owns AS_ROWS (= 8) agent rows; every weight row is streamed once per CTA straight from
// L2 with lanes along k (coalesced), the 8 per-row partial dot products are folded with a 9-shuffle
// multi-value butterfly, gate math and the argmax/selection run warp-per-row.
#pragma once
#include "mal_common.cuh"

#define AS_ROWS 8
#define AS_THREADS 256
#define AS_WARPS (AS_THREADS / 32)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter layout identical to curand_init(seed, subsequence, offset) as used by
// ATen/native/cuda/DistributionTemplates.h (distribution_elementwise_grid_stride_kernel).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// The uniform in (0,1] that torch's grid-stride distribution kernel hands to element `li` of a tensor of
// `numel` elements when the generator state is (seed, offset): thread idx = li % (256*grid), the (li / (256*grid))-th
// 32-bit output of that thread's stream.
__device__ __forceinline__ float torch_philox_uniform(uint64_t seed, uint64_t offset, uint32_t grid, uint64_t li) {
    const uint64_t span = 256ull * grid;
    const uint64_t idx = li % span, q = li / span;
    const uint64_t ctr = offset / 4 + q / 4;
    uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)idx, (uint32_t)(idx >> 32));
    uint4 o = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t comp = (uint32_t)(q & 3);
    uint32_t x = comp == 0 ? o.x : comp == 1 ? o.y : comp == 2 ? o.z : o.w;
    return x * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f);   // _curand_uniform
}
__device__ __forceinline__ float torch_uniform01(float r) { return r == 1.0f ? 0.0f : r; }   // uniform_kernel
__device__ __forceinline__ float torch_exponential1(float r) {                                 // transformation::exponential
    const float eps_half = 1.1920928955078125e-07f / 2.0f;
    float lg = (r >= 1.0f - eps_half) ? -eps_half : logf(r);
    return -1.0f * lg;
}

// Rollout-step fusion (steppers/episode_stepper.py:110-142,177-186): besides selecting actions the launch also performs
// the EpisodeBatch.update calls around select_actions -- the environment's pre-transition data of step t, the previous
// step's reward / terminated, and the selected actions + their one-hot -- straight into the episode records.
struct RolloutIO {
    int enabled;
    int S;
    const float *env_state; int64_t env_state_sb;     // [bs,S]
    const uint8_t *alive;                              // [bs] or NULL: matches that have ended select on a dummy avail row
    const float *prev_reward; const uint8_t *prev_done;   // [bs] outcome of step t-1, NULL at t = 0
    float *state_t; int64_t state_sb;                 // episode batch fields at time index t: &field[0, t], batch strides in elements
    int32_t *avail_t; int64_t avail_sb;
    float *obs_t; int64_t obs_sb;
    long long *filled_t; int64_t filled_sb;
    long long *actions_t; int64_t actions_sb;
    float *onehot_t; int64_t onehot_sb;
    float *reward_tm1; int64_t reward_sb;             // &reward[0, t-1]
    uint8_t *term_tm1; int64_t term_sb;
};

struct SelectArgs {
    const int32_t *avail;
    int64_t avail_sb;
    int32_t N;
    float epsilon;
    int32_t rng_mode;
    const float *u, *e;
    uint64_t seed, offset_u, offset_e;
    uint32_t grid_u, grid_e;
    int64_t *actions, *greedy;
    int32_t *status;
};

// One warp selects for one row.  qv = this lane's Q-value (lane < A).  With `io` the avail row, the chosen action and
// its one-hot also go into the episode batch at time index t.
__device__ __forceinline__ void select_row(const SelectArgs &s, int row, int A, int lane, float qv, const RolloutIO *io = nullptr) {
    const bool valid = lane < A;
    const int b = row / s.N, n = row - b * s.N;
    int av = valid ? s.avail[(int64_t)b * s.avail_sb + (int64_t)n * A + lane] : 0;
    if (io && valid) io->avail_t[(int64_t)b * io->avail_sb + (int64_t)n * A + lane] = av;
    if (io && io->alive && !io->alive[b]) av = (lane == 0) ? 1 : 0;       // an ended match: any valid row (its data is cleared later)
    // greedy branch: masked_q[avail == 0] = -inf ; max(dim=2)[1]
    float gv = (valid && av != 0) ? qv : -INFINITY;
    int gi = lane;
    warp_argmax(gv, gi);
    // random branch: Categorical(avail.float()).sample() == argmax(p / Exp(1))
    const float avf = (float)av;
    const float tot = warp_sum(avf);
    float ev;
    if (s.rng_mode == 0) ev = valid ? s.e[(int64_t)row * A + lane] : 1.0f;
    else ev = valid ? torch_exponential1(torch_philox_uniform(s.seed, s.offset_e, s.grid_e, (uint64_t)row * A + lane)) : 1.0f;
    float rv = valid ? (avf / tot) / ev : -INFINITY;
    int ri = lane;
    warp_argmax(rv, ri);
    int act = 0;
    if (lane == 0) {
        float uu = (s.rng_mode == 0) ? s.u[row]
                                     : torch_uniform01(torch_philox_uniform(s.seed, s.offset_u, s.grid_u, (uint64_t)row));
        const long long pick_random = (uu < s.epsilon) ? 1 : 0;
        act = (int)(pick_random * ri + (1 - pick_random) * gi);
        s.actions[row] = act;
        s.greedy[row] = 1 - pick_random;
        if (!(tot > 0.0f) && s.status) atomicExch(s.status, 1);
    }
    if (io) {
        act = __shfl_sync(0xffffffffu, act, 0);
        if (lane == 0) io->actions_t[(int64_t)b * io->actions_sb + n] = act;
        if (valid) io->onehot_t[(int64_t)b * io->onehot_sb + (int64_t)n * A + lane] = (lane == act) ? 1.0f : 0.0f;   // OneHot, transforms.py:16-19
    }
}

// Per-match part of the fused rollout step (rows r0 .. r0+AS_ROWS-1 of this CTA; the row with n == 0 acts for its match):
// state / filled at t, reward / terminated at t-1.
__device__ __forceinline__ void rollout_match_fields(const RolloutIO &io, int r0, int rows, int N, int tid, int nthreads) {
    for (int r = 0; r < AS_ROWS; ++r) {
        const int row = r0 + r;
        if (row >= rows) break;
        const int b = row / N;
        if (row - b * N != 0) continue;
        for (int k = tid; k < io.S; k += nthreads) io.state_t[(int64_t)b * io.state_sb + k] = io.env_state[(int64_t)b * io.env_state_sb + k];
        if (tid == 0) {
            io.filled_t[(int64_t)b * io.filled_sb] = 1;
            if (io.prev_reward) {
                io.reward_tm1[(int64_t)b * io.reward_sb] = io.prev_reward[b];
                io.term_tm1[(int64_t)b * io.term_sb] = io.prev_done[b] ? 1 : 0;
            }
        }
    }
}

__global__ void __launch_bounds__(AS_THREADS) k_eps_greedy_select(const float *q, int64_t q_ld, int rows, int A,
                                                                  SelectArgs s) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    float qv = lane < A ? q[(int64_t)warp * q_ld + lane] : 0.0f;
    select_row(s, warp, A, lane, qv);
}

// fold 8 per-lane partials across the warp: afterwards lane L holds the full sum of value index (L >> 2).
__device__ __forceinline__ float warp_fold8(const float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w[4], u[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float send = b4 ? v[q] : v[q + 4];
        float keep = b4 ? v[q + 4] : v[q];
        w[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float send = b3 ? w[q] : w[q + 2];
        float keep = b3 ? w[q + 2] : w[q];
        u[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float send = b2 ? u[0] : u[1];
    float keep = b2 ? u[1] : u[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

struct AgentStepArgs {
    const float *params;
    int kind;                                   // MAL_AGENT_RNN | MAL_AGENT_DQN (feed-forward: no hidden state; stream kernel only)
    int rows, N, OBS, A;
    int dense;                                  // 1: obs is the full [rows, d_in] agent input (row stride obs_sb)
    const float *obs; int64_t obs_sb;
    const float *onehot; int64_t onehot_sb;     // NULL at t == 0
    const float *h_in;                          // NULL -> zeros
    float *h_out, *q;
    int do_select;
    SelectArgs sel;
    RolloutIO io;
};

// Latency plan: a CTA needs every weight exactly once, and nothing but the inputs depends on anything, so ALL
// global loads (inputs, 48 GRU weight rows per warp, fc1 / fc2 rows, biases) are issued back to back at kernel entry
// into registers; the phases below then only touch registers and shared memory.
#define AS_FC1_MAXK 7    // ceil(224 / 32): obs+action columns handled per lane
__global__ void __launch_bounds__(AS_THREADS, 1) k_agent_step(AgentStepArgs a) {
    extern __shared__ float as_smem[];
    const AgentLayout L = agent_layout(a.dense ? a.OBS : a.OBS + a.A + a.N, a.A);
    const int Kin = a.dense ? a.OBS : a.OBS + a.A;
    const int ldin = Kin + 1;
    float *in_s = as_smem;                       // [8][ldin]
    float *x_s = in_s + AS_ROWS * ldin;          // [8][64]
    float *h_s = x_s + AS_ROWS * HID;            // [8][64]
    float *g_s = h_s + AS_ROWS * HID;            // [8][384]  W_ih x | W_hh h (no biases)
    float *hn_s = g_s + AS_ROWS * 2 * G3;        // [8][64]
    float *q_s = hn_s + AS_ROWS * HID;           // [8][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * AS_ROWS;
    const float *P = a.params;

    // ---- issue every global load up front ---------------------------------------------------------------------
    float2 wg[2 * G3 / AS_WARPS];                // GRU rows warp, warp+8, ... of [W_ih ; W_hh]   (48 x float2)
#pragma unroll
    for (int i = 0; i < 2 * G3 / AS_WARPS; ++i) {
        const int jj = warp + AS_WARPS * i;
        const bool hh = jj >= G3;
        const int j = hh ? jj - G3 : jj;
        wg[i] = __ldg(reinterpret_cast<const float2 *>(P + (hh ? L.w_hh : L.w_ih) + (int64_t)j * HID) + lane);
    }
    float w1[HID / AS_WARPS][AS_FC1_MAXK];       // fc1 rows warp, warp+8, ...; columns lane, lane+32, ...
    float b1x[HID / AS_WARPS];                   // fc1 bias + agent-id column, for the row this lane finalises
#pragma unroll
    for (int i = 0; i < HID / AS_WARPS; ++i) {
        const int j = warp + AS_WARPS * i;
        const float *w = P + L.fc1_w + (int64_t)j * L.d_in;
#pragma unroll
        for (int c = 0; c < AS_FC1_MAXK; ++c) w1[i][c] = (lane + 32 * c < Kin) ? __ldg(w + lane + 32 * c) : 0.0f;
        const int row = r0 + (lane >> 2);
        const int n = row < a.rows ? row % a.N : 0;
        b1x[i] = __ldg(P + L.fc1_b + j) + (a.dense ? 0.0f : __ldg(w + Kin + n));
    }
    float2 w2[MAL_MAX_ACTIONS / AS_WARPS];
    float b2[MAL_MAX_ACTIONS / AS_WARPS];
#pragma unroll
    for (int i = 0; i < MAL_MAX_ACTIONS / AS_WARPS; ++i) {
        const int j = warp + AS_WARPS * i;
        w2[i] = make_float2(0.f, 0.f);
        b2[i] = 0.0f;
        if (j < a.A) {
            w2[i] = __ldg(reinterpret_cast<const float2 *>(P + L.fc2_w + (int64_t)j * HID) + lane);
            b2[i] = __ldg(P + L.fc2_b + j);
        }
    }
    // gate biases of the hidden unit this thread combines (items tid and tid+256 share i = tid & 63)
    const int gi_ = tid & 63;
    const float bir = __ldg(P + L.b_ih + gi_) + __ldg(P + L.b_hh + gi_);
    const float biz = __ldg(P + L.b_ih + HID + gi_) + __ldg(P + L.b_hh + HID + gi_);
    const float bin = __ldg(P + L.b_ih + 2 * HID + gi_), bhn = __ldg(P + L.b_hh + 2 * HID + gi_);

    // ---- stage inputs (obs | last action one-hot) and previous hidden state
    for (int idx = tid; idx < AS_ROWS * Kin; idx += AS_THREADS) {
        int r = idx / Kin, k = idx - r * Kin, row = r0 + r;
        float v = 0.0f;
        if (row < a.rows) {
            const int b = row / a.N, n = row - b * a.N;
            if (a.dense) v = a.obs[(int64_t)row * a.obs_sb + k];
            else if (k < a.OBS) {
                v = a.obs[(int64_t)b * a.obs_sb + (int64_t)n * a.OBS + k];
                if (a.io.enabled) a.io.obs_t[(int64_t)b * a.io.obs_sb + (int64_t)n * a.OBS + k] = v;   // pre-transition update of step t
            }
            else if (a.onehot) v = a.onehot[(int64_t)b * a.onehot_sb + (int64_t)n * a.A + (k - a.OBS)];
        }
        in_s[r * ldin + k] = v;
    }
    if (a.io.enabled) rollout_match_fields(a.io, r0, a.rows, a.N, tid, AS_THREADS);
    for (int idx = tid; idx < AS_ROWS * HID; idx += AS_THREADS) {
        int r = idx >> 6, row = r0 + r;
        h_s[idx] = (a.h_in && row < a.rows) ? a.h_in[(int64_t)row * HID + (idx & 63)] : 0.0f;
    }
    __syncthreads();

    // ---- fc1 + relu (agent-id one-hot column folded in as a bias gather)
#pragma unroll
    for (int i = 0; i < HID / AS_WARPS; ++i) {
        const int j = warp + AS_WARPS * i;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < AS_FC1_MAXK; ++c) {
            const int k = lane + 32 * c;
            if (k < Kin) {
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fmaf(w1[i][c], in_s[r * ldin + k], acc[r]);
            }
        }
        float s = warp_fold8(acc, lane);
        if ((lane & 3) == 0) x_s[(lane >> 2) * HID + j] = fmaxf(s + b1x[i], 0.0f);
    }
    __syncthreads();

    // ---- W_ih x  and  W_hh h   (384 weight rows of 64, already in registers)
    {
        float xv[8][2], hv[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            xv[r][0] = x_s[r * HID + 2 * lane]; xv[r][1] = x_s[r * HID + 2 * lane + 1];
            hv[r][0] = h_s[r * HID + 2 * lane]; hv[r][1] = h_s[r * HID + 2 * lane + 1];
        }
#pragma unroll
        for (int i = 0; i < 2 * G3 / AS_WARPS; ++i) {
            const int jj = warp + AS_WARPS * i;
            const bool hh = jj >= G3;          // compile-time per i after unrolling only if warp-independent; cheap anyway
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                acc[r] = hh ? fmaf(wg[i].y, hv[r][1], wg[i].x * hv[r][0]) : fmaf(wg[i].y, xv[r][1], wg[i].x * xv[r][0]);
            float s = warp_fold8(acc, lane);
            if ((lane & 3) == 0) g_s[(lane >> 2) * 2 * G3 + jj] = s;
        }
    }
    __syncthreads();

    // ---- gate math: r,z,n ; h' = n + z (h - n)
    for (int idx = tid; idx < AS_ROWS * HID; idx += AS_THREADS) {
        int r = idx >> 6, i = idx & 63, row = r0 + r;
        const float *g = g_s + r * 2 * G3;
        float rr = sigmoidf_acc(g[i] + g[G3 + i] + bir);
        float zz = sigmoidf_acc(g[HID + i] + g[G3 + HID + i] + biz);
        float nn = tanhf(g[2 * HID + i] + bin + rr * (g[G3 + 2 * HID + i] + bhn));
        float hp = h_s[idx];
        float hn = nn + zz * (hp - nn);
        hn_s[idx] = hn;
        if (row < a.rows) a.h_out[(int64_t)row * HID + i] = hn;
    }
    __syncthreads();

    // ---- fc2
    {
        float hv[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) { hv[r][0] = hn_s[r * HID + 2 * lane]; hv[r][1] = hn_s[r * HID + 2 * lane + 1]; }
#pragma unroll
        for (int i = 0; i < MAL_MAX_ACTIONS / AS_WARPS; ++i) {
            const int j = warp + AS_WARPS * i;
            if (j < a.A) {      // warp-uniform
                float acc[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fmaf(w2[i].y, hv[r][1], w2[i].x * hv[r][0]);
                float s = warp_fold8(acc, lane);
                if ((lane & 3) == 0) {
                    int r = lane >> 2, row = r0 + r;
                    s += b2[i];
                    q_s[r * 32 + j] = s;
                    if (row < a.rows) a.q[(int64_t)row * a.A + j] = s;
                }
            }
        }
    }
    if (!a.do_select) return;
    __syncthreads();
    // ---- epsilon-greedy selection: one warp per row
    {
        int row = r0 + warp;
        if (row < a.rows) select_row(a.sel, row, a.A, lane, lane < a.A ? q_s[warp * 32 + lane] : 0.0f, a.io.enabled ? &a.io : nullptr);
    }
}

// Throughput variant for large row counts (many CTAs per SM): weight rows are streamed from L2 as they are used
// instead of being parked in 255 registers per thread.
__global__ void __launch_bounds__(AS_THREADS) k_agent_step_stream(AgentStepArgs a) {
    extern __shared__ float as_smem[];
    const AgentLayout L = agent_layout(a.dense ? a.OBS : a.OBS + a.A + a.N, a.A, a.kind);
    const bool rnn = a.kind == MAL_AGENT_RNN;
    const int Kin = a.dense ? a.OBS : a.OBS + a.A;
    const int ldin = Kin + 1;
    float *in_s = as_smem;                       // [8][ldin]
    float *x_s = in_s + AS_ROWS * ldin;          // [8][64]
    float *h_s = x_s + AS_ROWS * HID;            // [8][64]
    float *g_s = h_s + AS_ROWS * HID;            // [8][384]  gi | gh
    float *hn_s = g_s + AS_ROWS * 2 * G3;        // [8][64]
    float *q_s = hn_s + AS_ROWS * HID;           // [8][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * AS_ROWS;
    const float *P = a.params;

    // ---- stage inputs (obs | last action one-hot) and previous hidden state
    for (int idx = tid; idx < AS_ROWS * Kin; idx += AS_THREADS) {
        int r = idx / Kin, k = idx - r * Kin, row = r0 + r;
        float v = 0.0f;
        if (row < a.rows) {
            const int b = row / a.N, n = row - b * a.N;
            if (a.dense) v = a.obs[(int64_t)row * a.obs_sb + k];
            else if (k < a.OBS) {
                v = a.obs[(int64_t)b * a.obs_sb + (int64_t)n * a.OBS + k];
                if (a.io.enabled) a.io.obs_t[(int64_t)b * a.io.obs_sb + (int64_t)n * a.OBS + k] = v;   // pre-transition update of step t
            }
            else if (a.onehot) v = a.onehot[(int64_t)b * a.onehot_sb + (int64_t)n * a.A + (k - a.OBS)];
        }
        in_s[r * ldin + k] = v;
    }
    if (a.io.enabled) rollout_match_fields(a.io, r0, a.rows, a.N, tid, AS_THREADS);
    for (int idx = tid; idx < AS_ROWS * HID; idx += AS_THREADS) {
        int r = idx >> 6, row = r0 + r;
        h_s[idx] = (rnn && a.h_in && row < a.rows) ? a.h_in[(int64_t)row * HID + (idx & 63)] : 0.0f;
    }
    __syncthreads();

    // ---- fc1 + relu (agent-id one-hot column folded in as a bias gather)
    for (int j = warp; j < HID; j += AS_WARPS) {
        const float *w = P + L.fc1_w + (int64_t)j * L.d_in;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = lane; k < Kin; k += 32) {
            float wv = __ldg(w + k);
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(wv, in_s[r * ldin + k], acc[r]);
        }
        float s = warp_fold8(acc, lane);
        if ((lane & 3) == 0) {
            int r = lane >> 2, row = r0 + r;
            int n = row < a.rows ? row % a.N : 0;
            s += __ldg(P + L.fc1_b + j) + (a.dense ? 0.0f : __ldg(w + Kin + n));
            x_s[r * HID + j] = fmaxf(s, 0.0f);
        }
    }
    __syncthreads();

    // ---- gi = W_ih x + b_ih ; gh = W_hh h + b_hh   (384 weight rows of 64)
    if (rnn) {
        float xv[8][2], hv[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            xv[r][0] = x_s[r * HID + 2 * lane]; xv[r][1] = x_s[r * HID + 2 * lane + 1];
            hv[r][0] = h_s[r * HID + 2 * lane]; hv[r][1] = h_s[r * HID + 2 * lane + 1];
        }
#pragma unroll 4
        for (int jj = warp; jj < 2 * G3; jj += AS_WARPS) {
            const bool hh = jj >= G3;
            const int j = hh ? jj - G3 : jj;
            const float2 wv = __ldg(reinterpret_cast<const float2 *>(P + (hh ? L.w_hh : L.w_ih) + (int64_t)j * HID) + lane);
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                acc[r] = hh ? fmaf(wv.y, hv[r][1], wv.x * hv[r][0]) : fmaf(wv.y, xv[r][1], wv.x * xv[r][0]);
            float s = warp_fold8(acc, lane);
            if ((lane & 3) == 0) g_s[(lane >> 2) * 2 * G3 + jj] = s + __ldg(P + (hh ? L.b_hh : L.b_ih) + j);
        }
    }
    __syncthreads();

    // ---- gate math: r,z,n ; h' = n + z (h - n)        (DQN: q = fc2(relu(fc1(.))), dqn_agent.py:34-37)
    if (!rnn) {
        for (int idx = tid; idx < AS_ROWS * HID; idx += AS_THREADS) hn_s[idx] = x_s[idx];
    } else
    for (int idx = tid; idx < AS_ROWS * HID; idx += AS_THREADS) {
        int r = idx >> 6, i = idx & 63, row = r0 + r;
        const float *g = g_s + r * 2 * G3;
        float rr = sigmoidf_acc(g[i] + g[G3 + i]);
        float zz = sigmoidf_acc(g[HID + i] + g[G3 + HID + i]);
        float nn = tanhf(g[2 * HID + i] + rr * g[G3 + 2 * HID + i]);
        float hp = h_s[idx];
        float hn = nn + zz * (hp - nn);
        hn_s[idx] = hn;
        if (row < a.rows) a.h_out[(int64_t)row * HID + i] = hn;
    }
    __syncthreads();

    // ---- fc2
    {
        float hv[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) { hv[r][0] = hn_s[r * HID + 2 * lane]; hv[r][1] = hn_s[r * HID + 2 * lane + 1]; }
        for (int j = warp; j < a.A; j += AS_WARPS) {
            const float2 wv = __ldg(reinterpret_cast<const float2 *>(P + L.fc2_w + (int64_t)j * HID) + lane);
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(wv.y, hv[r][1], wv.x * hv[r][0]);
            float s = warp_fold8(acc, lane);
            if ((lane & 3) == 0) {
                int r = lane >> 2, row = r0 + r;
                s += __ldg(P + L.fc2_b + j);
                q_s[r * 32 + j] = s;
                if (row < a.rows) a.q[(int64_t)row * a.A + j] = s;
            }
        }
    }
    if (!a.do_select) return;
    __syncthreads();
    // ---- epsilon-greedy selection: one warp per row
    {
        int row = r0 + warp;
        if (row < a.rows) select_row(a.sel, row, a.A, lane, lane < a.A ? q_s[warp * 32 + lane] : 0.0f, a.io.enabled ? &a.io : nullptr);
    }
}
