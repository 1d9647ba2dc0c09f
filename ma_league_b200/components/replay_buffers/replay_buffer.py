"""Device-resident episode ring buffer with the reference's API
(marl/components/replay_buffers/replay_buffer.py:5-59) on top of packed episode records."""
import numpy as np
import torch as th

from ... import _native as nat
from ..episode_batch import EpisodeBatch


class ReplayBuffer(EpisodeBatch):
    def __init__(self, scheme, groups, buffer_size: int, max_seq_length: int, preprocess=None, device="cpu"):
        super().__init__(scheme, groups, buffer_size, max_seq_length, preprocess=preprocess, device=device)
        self.buffer_size = buffer_size
        self.buffer_index = 0
        self.episodes_in_buffer = 0

    def insert_episode_batch(self, ep_batch: EpisodeBatch):
        """replay_buffer.py:22-41: write n episodes at the ring cursor, splitting on wrap-around."""
        n = ep_batch.batch_size
        if self.buffer_index + n <= self.buffer_size:
            if not self._insert_records(ep_batch):
                self.update(ep_batch.data.transition_data, slice(self.buffer_index, self.buffer_index + n),
                            slice(0, ep_batch.max_seq_length), mark_filled=False)
                self.update(ep_batch.data.episode_data, slice(self.buffer_index, self.buffer_index + n))
            self.buffer_index = self.buffer_index + n
            self.episodes_in_buffer = max(self.episodes_in_buffer, self.buffer_index)
            self.buffer_index = self.buffer_index % self.buffer_size
            assert self.buffer_index < self.buffer_size
        else:
            buffer_left = self.buffer_size - self.buffer_index
            self.insert_episode_batch(ep_batch[0:buffer_left, :])
            self.insert_episode_batch(ep_batch[buffer_left:, :])

    def _insert_records(self, ep_batch) -> bool:
        """Fast path: same record layout on the same CUDA device -> one bulk record copy into the ring slots."""
        src = _record_source(ep_batch)
        if src is None or self._layout is None or not self._layout.same_as(src[0]):
            return False
        layout, base, first = src
        if self._storage.device.type != "cuda" or base.device != self._storage.device:
            return False
        rb = layout.record_bytes
        n = ep_batch.batch_size
        dst = self._storage[self.buffer_index * rb:]
        srcv = base[first * rb:]
        with th.cuda.device(self._storage.device):
            nat.check(nat.lib().mal_record_copy(nat.ptr(dst), rb, None, nat.ptr(srcv), rb, None, n, rb,
                                                nat.current_stream(self._storage.device)), "mal_record_copy")
        return True

    def can_sample(self, batch_size: int) -> bool:
        return self.episodes_in_buffer >= batch_size

    def sample(self, batch_size: int) -> EpisodeBatch:
        """replay_buffer.py:46-53: whole buffer (views) if it holds exactly one batch, else uniform w/o replacement."""
        assert self.can_sample(batch_size)
        if self.episodes_in_buffer == batch_size:
            return self[:batch_size]
        ep_ids = np.random.choice(self.episodes_in_buffer, batch_size, replace=False)
        return self[ep_ids]

    def sample_into(self, out: EpisodeBatch) -> EpisodeBatch:
        """`sample(out.batch_size)` written into the caller's packed batch `out` (same scheme, on this device) instead of
        a fresh one: same episode ids as `sample` for the same numpy RNG state, one bulk record-copy launch, and
        stable device addresses from step to step (so the learner's captured CUDA graph can be replayed)."""
        n = out.batch_size
        assert self.can_sample(n)
        if self._layout is None or out._layout is None or not self._layout.same_as(out._layout):
            raise nat.MalError("sample_into needs packed batches with the buffer's record layout")
        if self._storage.device.type != "cuda" or out._storage.device != self._storage.device:
            raise nat.MalError("sample_into: the buffer and the output batch must live on the same CUDA device")
        if self.episodes_in_buffer == n:
            ids = np.arange(n, dtype=np.int64)       # replay_buffer.py:48-49 returns the first n episodes
        else:
            ids = np.random.choice(self.episodes_in_buffer, n, replace=False).astype(np.int64)
        dev = self._storage.device
        ids_d = th.from_numpy(ids).to(dev, non_blocking=True)
        rb = self._layout.record_bytes
        with th.cuda.device(dev):
            nat.check(nat.lib().mal_record_copy(nat.ptr(out._storage), rb, None, nat.ptr(self._storage), rb,
                                                nat.ptr(ids_d), n, rb, nat.current_stream(dev)), "mal_record_copy")
        return out

    def __repr__(self):
        return "ReplayBuffer. {}/{} episodes. Keys:{} Groups:{}".format(
            self.episodes_in_buffer, self.buffer_size, self.scheme.keys(), self.groups.keys())


def _record_source(ep_batch):
    """(layout, uint8 base storage, first record) if `ep_batch` is a packed batch or a contiguous batch-slice view
    of one (time untouched); None otherwise."""
    if getattr(ep_batch, "_layout", None) is not None:
        return ep_batch._layout, ep_batch._storage, 0
    parent = getattr(ep_batch, "_parent_records", None)
    return parent
