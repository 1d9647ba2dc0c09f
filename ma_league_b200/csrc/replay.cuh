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
