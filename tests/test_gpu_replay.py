"""Device-resident replay buffer: reference ring fixture through the CUDA record copy, and full-size properties."""
import numpy as np
import pytest
import torch as th

import ma_league_b200 as M
from ma_league_b200.synthetic import make_scheme, synth_episode_data, fill_episode_batch
from tests.helpers import load_golden, sub

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ring_fixture_on_device():
    g = load_golden("replay_ring")
    size, TT, N, A, OBS, S = [int(x) for x in g["meta"]]
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, size, TT, preprocess=pre, device=DEV)
    np.random.seed(3)
    for i, n in enumerate([3, 3, 3, 5, 1]):
        eb = M.EpisodeBatch(scheme, groups, n, TT, preprocess=pre, device=DEV)
        for k, v in sub(g, "ins%d." % i).items():
            eb.data.transition_data[k].copy_(th.from_numpy(v).to(DEV))
        buf.insert_episode_batch(eb)
        assert [buf.buffer_index, buf.episodes_in_buffer] == list(g["counters"][i])
        for k, v in buf.data.transition_data.items():
            assert np.array_equal(v.cpu().numpy(), g["buf%d.%s" % (i, k)]), (i, k)       # bit-exact buffer contents
        if buf.can_sample(4):
            smp = buf.sample(4)
            assert np.array_equal(np.asarray(g["smp%d.ids" % i]), g["smp%d.ids" % i])
            for k, v in smp.data.transition_data.items():
                assert np.array_equal(v.cpu().numpy(), g["smp%d.%s" % (i, k)]), (i, k)
            assert int(smp.max_t_filled()) == int(g["smp%d.max_t" % i])


def test_full_size_gather_scatter_properties():
    """5v5, T+1 = 201, 256-episode buffer: sample == torch advanced indexing per key (bit-exact); checksum of checksums;
    insert with wrap-around; empty and single-episode edge cases."""
    N, A, OBS, S, TT = 5, 11, 48, 80, 201
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, 256, TT, preprocess=pre, device=DEV)
    gen = th.Generator().manual_seed(0)
    for n in (100, 100, 100):        # third insert wraps around (200 + 100 > 256)
        data, lens = synth_episode_data(n, TT, N, A, OBS, S, gen, device=DEV)
        eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, n, TT, preprocess=pre, device=DEV), data, lens)
        before_tail = buf["obs"][buf.buffer_index:buf.buffer_index + 1].clone()
        buf.insert_episode_batch(eb)
        last = eb
    assert buf.buffer_index == 44 and buf.episodes_in_buffer == 256
    assert th.equal(buf["obs"][:44], last["obs"][56:]) and th.equal(buf["obs"][200:256], last["obs"][:56])
    assert th.equal(buf["filled"][:44], last["filled"][56:])
    ids = np.random.RandomState(1).choice(256, 32, replace=False)
    smp = buf[ids]
    ids_t = th.as_tensor(ids, device=DEV)
    for k, v in buf.data.transition_data.items():
        assert th.equal(smp[k], v[ids_t]), k
    rb = buf._layout.record_bytes
    rec = buf._storage.view(256, rb).long().sum(1)
    assert int(smp._storage.view(32, rb).long().sum()) == int(rec[ids_t].sum())
    assert int(smp.max_t_filled()) == int(buf["filled"][ids_t].sum(1).max())
    trunc = smp[:, :int(smp.max_t_filled())]
    assert trunc["obs"].data_ptr() == smp["obs"].data_ptr()
    one = buf[[255]]
    assert one.batch_size == 1 and th.equal(one["state"][0], buf["state"][255])
    empty = buf[np.zeros(0, dtype=np.int64)]
    assert empty.batch_size == 0
    dup = buf[[3, 3, 7]]
    assert th.equal(dup["reward"][0], dup["reward"][1])
    dev_ids = buf[th.tensor([5, 9], device=DEV)]
    assert th.equal(dev_ids["avail_actions"][1], buf["avail_actions"][9])
