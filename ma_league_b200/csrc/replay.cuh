// K10: replay-buffer data movement over packed episode records.
//
// Reference: ReplayBuffer.sample -> EpisodeBatch.__getitem__ (index-array gather, one advanced-index
// copy per scheme key; marl/components/replay_buffers/replay_buffer.py:46-53, episode_batch.py:226-238)
// and insert_episode_batch -> EpisodeBatch.update (slice assignment per key; replay_buffer.py:22-41).
//
// B200 design: an episode is ONE 128-byte-aligned record holding every scheme key, so sample/insert is a
// pure record copy.  Each CTA streams 16 KB tiles HBM -> shared memory -> HBM with the bulk-copy engine
// (cp.async.bulk, SASS UBLKCP) through a 4-deep mbarrier ring; a single elected thread drives the engine,
// no register staging, no per-element address math.  Algorithmic bytes = 2 * n * record_bytes.
#pragma once
#include "mal_common.cuh"

#define RC_TILE 16384
#define RC_STAGES 4

struct RecordCopyArgs {
    uint8_t *dst;
    const uint8_t *src;
    int64_t dst_stride, src_stride;
    const int64_t *dst_ids, *src_ids;
    int32_t n;
    int64_t bytes;          // per record, multiple of 16
    int32_t tiles_per_rec;
    int64_t n_tiles;
};

__device__ __forceinline__ void rc_tile_addr(const RecordCopyArgs &a, int64_t tile, const uint8_t *&s, uint8_t *&d,
                                             uint32_t &len) {
    int64_t rec = tile / a.tiles_per_rec;
    int64_t off = (tile - rec * a.tiles_per_rec) * (int64_t)RC_TILE;
    int64_t srec = a.src_ids ? a.src_ids[rec] : rec;
    int64_t drec = a.dst_ids ? a.dst_ids[rec] : rec;
    int64_t left = a.bytes - off;
    len = (uint32_t)(left < RC_TILE ? left : RC_TILE);
    s = a.src + srec * a.src_stride + off;
    d = a.dst + drec * a.dst_stride + off;
}

__global__ void __launch_bounds__(32, 1) k_record_copy_tma(RecordCopyArgs a) {
    extern __shared__ __align__(128) uint8_t rc_smem[];
    __shared__ __align__(8) uint64_t bars[RC_STAGES];
    if (threadIdx.x != 0) return;   // one elected thread drives the copy engine
    for (int s = 0; s < RC_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");

    const int64_t first = blockIdx.x, step = gridDim.x;
    const int64_t mine = (a.n_tiles > first) ? (a.n_tiles - first + step - 1) / step : 0;
    // prologue: fill the ring
    for (int64_t i = 0; i < mine && i < RC_STAGES; ++i) {
        const uint8_t *s; uint8_t *d; uint32_t len;
        rc_tile_addr(a, first + i * step, s, d, len);
        mbar_expect_tx(&bars[i], len);
        bulk_g2s(rc_smem + i * RC_TILE, s, len, &bars[i]);
    }
    for (int64_t i = 0; i < mine; ++i) {
        const int st = (int)(i % RC_STAGES);
        const uint8_t *s; uint8_t *d; uint32_t len;
        rc_tile_addr(a, first + i * step, s, d, len);
        mbar_wait(&bars[st], (uint32_t)((i / RC_STAGES) & 1));
        bulk_s2g(d, rc_smem + st * RC_TILE, len);
        bulk_commit();
        const int64_t nxt = i + RC_STAGES;
        if (nxt < mine) {
            bulk_wait_read0();   // the store has finished reading this stage: safe to refill it
            const uint8_t *s2; uint8_t *d2; uint32_t len2;
            rc_tile_addr(a, first + nxt * step, s2, d2, len2);
            mbar_expect_tx(&bars[st], len2);
            bulk_g2s(rc_smem + st * RC_TILE, s2, len2, &bars[st]);
        }
    }
    bulk_wait_all();
}

// EpisodeBatch.max_t_filled (episode_batch.py:240-242): max over episodes of the number of filled steps.
__global__ void k_max_t_filled(const int64_t *filled, int64_t sb, int64_t st, int B, int TT, int *out) {
    __shared__ int best;
    if (threadIdx.x == 0) best = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    int local = 0;
    for (int b = warp; b < B; b += nwarp) {
        long long s = 0;
        for (int t = lane; t < TT; t += 32) s += filled[b * sb + t * st];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        local = max(local, (int)s);
    }
    if (lane == 0) atomicMax(&best, local);
    __syncthreads();
    if (threadIdx.x == 0) out[0] = best;
}

// =============================================================================================
// Compact wire records for the host <-> device path of a host-resident replay buffer (`buffer_cpu_only`: the sampled
// batch crosses PCIe on every learner step, runs/train/ma_experiment.py:238-239 `batch.to(args.device)`).
// The packed episode record carries two fields that are pure functions of others -- `actions_onehot` (f32 [N,A] per
// step, OneHot of `actions`, transforms.py:16-19) and `filled` (i64) -- and three that are far wider than their content
// (`actions` i64, `avail_actions` i32 [N,A] of 0/1 flags, `terminated`).  The wire record ships
//     state f32 | obs f32 | reward f32 | actions u8 [N] | avail bitmask u32 [N] | flags u8 (bit 0 filled, bit 1 terminated)
// per stored step (26 % fewer bytes at 5v5, 30 % at 20v20) and k_wire_unpack re-expands it on the device into the full
// record (bit-exact round trip for 0/1 avail flags, actions < 256, filled / terminated in {0, 1}; the packer reports
// anything else through `status`).  One warp per (episode, step); HBM-bound, algorithmic bytes = wire + full record.
// =============================================================================================
struct WireArgs {
    int B, TT, N, A, OBS, S;
    mal_field_t obs, onehot, actions, avail, state, reward, terminated, filled;   // the full (packed-record) batch
    uint8_t *wire;
    int64_t record_bytes, off_state, off_obs, off_reward, off_actions, off_avail, off_flags;
    int *status;                                       // pack: set to 1 when a value does not fit the wire encoding
};

template <bool PACK>
__global__ void __launch_bounds__(256) k_wire(WireArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = (int64_t)a.B * a.TT;
    for (int64_t w = (int64_t)blockIdx.x * 8 + warp; w < total; w += (int64_t)gridDim.x * 8) {
        const int b = (int)(w / a.TT), t = (int)(w - (int64_t)b * a.TT);
        uint8_t *rec = a.wire + (int64_t)b * a.record_bytes;
        float *w_state = reinterpret_cast<float *>(rec + a.off_state) + (int64_t)t * a.S;
        float *w_obs = reinterpret_cast<float *>(rec + a.off_obs) + (int64_t)t * a.N * a.OBS;
        float *w_reward = reinterpret_cast<float *>(rec + a.off_reward) + t;
        uint8_t *w_act = rec + a.off_actions + (int64_t)t * a.N;
        uint32_t *w_avail = reinterpret_cast<uint32_t *>(rec + a.off_avail) + (int64_t)t * a.N;
        uint8_t *w_flags = rec + a.off_flags + t;
        float *f_state = const_cast<float *>(field_ptr<float>(a.state, b, t));
        float *f_obs = const_cast<float *>(field_ptr<float>(a.obs, b, t));
        float *f_onehot = const_cast<float *>(field_ptr<float>(a.onehot, b, t));
        long long *f_act = const_cast<long long *>(field_ptr<long long>(a.actions, b, t));
        int *f_avail = const_cast<int *>(field_ptr<int>(a.avail, b, t));
        float *f_reward = const_cast<float *>(field_ptr<float>(a.reward, b, t));
        unsigned char *f_term = const_cast<unsigned char *>(field_ptr<unsigned char>(a.terminated, b, t));
        long long *f_filled = const_cast<long long *>(field_ptr<long long>(a.filled, b, t));
        if (PACK) {
            for (int k = lane; k < a.S; k += 32) w_state[k] = f_state[k];
            for (int k = lane; k < a.N * a.OBS; k += 32) w_obs[k] = f_obs[k];
            const long long fl = *f_filled;                 // every lane: same address, one broadcast load
            const unsigned char tm = *f_term;
            bool bad = (fl != 0 && fl != 1) || tm > 1;
            for (int n = lane; n < a.N; n += 32) {
                const long long act = f_act[n];
                bad = bad || act < 0 || act > 255;
                w_act[n] = (uint8_t)act;
                uint32_t bits = 0;
                for (int j = 0; j < a.A; ++j) {
                    const int av = f_avail[n * a.A + j];
                    bad = bad || (av != 0 && av != 1);
                    bits |= (av != 0 ? 1u : 0u) << j;
                    // the wire drops actions_onehot: the unpacker writes OneHot(action) on filled steps and zeros elsewhere,
                    // which is what EpisodeBatch.update leaves (episode_batch.py:163-166, 183-195) -- anything else is reported
                    const float oh = f_onehot[n * a.A + j];
                    bad = bad || oh != ((fl != 0 && j == act) ? 1.0f : 0.0f);
                }
                w_avail[n] = bits;
            }
            if (lane == 0) {
                *w_reward = *f_reward;
                *w_flags = (uint8_t)((fl != 0 ? 1 : 0) | (tm != 0 ? 2 : 0));
            }
            if (bad && a.status) atomicExch(a.status, 1);
        } else {
            const uint8_t flags = *w_flags;
            const bool filled = flags & 1;
            for (int k = lane; k < a.S; k += 32) f_state[k] = w_state[k];
            for (int k = lane; k < a.N * a.OBS; k += 32) f_obs[k] = w_obs[k];
            for (int e = lane; e < a.N * a.A; e += 32) {
                const int n = e / a.A, j = e - n * a.A;
                f_avail[e] = (int)((w_avail[n] >> j) & 1u);
                f_onehot[e] = (filled && j == (int)w_act[n]) ? 1.0f : 0.0f;
            }
            for (int n = lane; n < a.N; n += 32) f_act[n] = (long long)w_act[n];
            if (lane == 0) {
                *f_reward = *w_reward;
                *f_term = (flags >> 1) & 1;
                *f_filled = filled ? 1 : 0;
            }
        }
    }
}
