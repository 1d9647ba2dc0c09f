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
            # sample() must sit at the same position of numpy's legacy stream as the reference's np.random.choice
            # (replay_buffer.py:52): predict the draw, rewind, sample, and compare ids AND the stream position after
            st0 = np.random.get_state()
            predicted = np.random.choice(buf.episodes_in_buffer, 4, replace=False)
            st1 = np.random.get_state()
            np.random.set_state(st0)
            smp = buf.sample(4)
            assert np.array_equal(predicted, g["smp%d.ids" % i]), (i, predicted)
            after = np.random.get_state()
            assert after[2] == st1[2] and np.array_equal(after[1], st1[1])      # exactly one choice() was consumed
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


def test_train_from_buffer_equals_sample_truncate_train():
    """the fused input pipeline (sample into a persistent batch, mask instead of truncate, no host
    sync) lands on the same update as the reference's sample -> max_t_filled -> truncate -> train sequence."""
    import numpy as np
    from tests.gpu_helpers import seeded_system, np_params
    from tests.helpers import assert_close

    def run(fused, truncate=False):
        s = seeded_system(3, 6, 14, "qmix", True, seed=31, buffer_size=24, learner_log_interval=0)
        gen = th.Generator().manual_seed(77)
        from ma_league_b200.synthetic import synth_episode_data, fill_episode_batch
        import ma_league_b200 as M
        for _ in range(4):      # 24 episodes of varying length (none of full length: truncation is exercised)
            data, lens = synth_episode_data(6, 14, 3, 9, 32, 48, gen, var_len=True, device="cuda")
            lens = th.clamp(lens, max=10)
            data["terminated"].zero_()
            data["terminated"][th.arange(6), lens - 1] = 1
            eb = fill_episode_batch(M.EpisodeBatch(s.scheme, s.groups, 6, 14, preprocess=s.pre, device="cuda"), data, lens)
            s.buf.insert_episode_batch(eb)
        np.random.seed(5)
        for i in range(3):
            if fused:
                b = s.learner.train_from_buffer(s.buf, 6, t_env=i, episode_num=i, truncate=truncate)
                assert b.max_seq_length == (14 if not truncate else int(b.max_t_filled()))
            else:
                smp = s.buf.sample(6)
                mt = int(smp.max_t_filled())
                assert mt < 14
                s.learner.train(smp[:, :mt], t_env=i, episode_num=i)
        th.cuda.synchronize()
        return np_params(s.mac.agent), np_params(s.learner.mixer), {k: v[0] for k, v in s.logger.stats.items()}

    ref = run(False)
    for variant in (run(True, False), run(True, True)):
        for x, y in zip(variant[:2], ref[:2]):
            for k in y:
                assert_close(x[k], y[k], 1e-5, k)
        for k, v in ref[2].items():
            assert abs(variant[2][k] - v) <= 1e-5 * max(1.0, abs(v)), (k, variant[2][k], v)


def test_train_from_buffer_against_oracle():
    """SURVEY.md 8(f3) against the ORACLE (not against the repo's own sample -> truncate -> train): the fused input
    pipeline samples with numpy's stream (ids predicted from the same state), gathers into its staging batch and trains
    over the full sequence length with the padding masked; the oracle runs the reference's sequence
    (ma_experiment.py:231-241: sample -> max_t_filled -> truncate -> train) on the SAME episode ids in fp64."""
    from oracle import np_oracle as O
    from tests.gpu_helpers import seeded_system, np_params, np_batch
    from tests.helpers import assert_close
    clip = 0.5
    s = seeded_system(3, 6, 14, "qmix", True, seed=33, buffer_size=24, learner_log_interval=0, clip=clip)
    gen = th.Generator().manual_seed(78)
    for _ in range(4):      # 24 episodes, none of full length: the truncation bound is < max_seq_length
        data, lens = synth_episode_data(6, 14, 3, 9, 32, 48, gen, var_len=True, device="cuda")
        lens = th.clamp(lens, max=10)
        data["terminated"].zero_()
        data["terminated"][th.arange(6), lens - 1] = 1
        eb = fill_episode_batch(M.EpisodeBatch(s.scheme, s.groups, 6, 14, preprocess=s.pre, device="cuda"), data, lens)
        s.buf.insert_episode_batch(eb)
    host_buf = np_batch(s.buf)
    L = s.learner
    a = s.args
    for variant in ("masked", "truncate"):
        p_agent, p_tagent = np_params(s.mac.agent), np_params(L.target_mac.agent)
        p_mixer, p_tmixer = np_params(L.mixer), np_params(L.target_mixer)
        sq0 = L.optimiser.flat_sq.cpu().numpy().astype(np.float64)
        np.random.seed(9)
        ids = np.random.choice(s.buf.episodes_in_buffer, 6, replace=False)
        np.random.seed(9)
        b = L.train_from_buffer(s.buf, 6, t_env=0, episode_num=0, truncate=(variant == "truncate"))
        th.cuda.synchronize()
        smp = {k: v[ids] for k, v in host_buf.items()}
        mt = O.max_t_filled(smp["filled"])
        assert mt < 14 and b.max_seq_length == (mt if variant == "truncate" else 14)
        smp = {k: v[:, :mt] for k, v in smp.items()}
        ref = O.learner_forward_backward(p_agent, p_tagent, p_mixer, p_tmixer, smp, mixer="qmix", double_q=True,
                                         gamma=a.gamma, dtype=np.float64)
        g_ref = np.concatenate([v.ravel() for v in list(ref["agent_grads"].values()) + list(ref["mixer_grads"].values())])
        norm, (g_clip,) = O.clip_grad_norm([g_ref], clip)
        assert norm > clip
        st = {k: v[0] for k, v in s.logger.stats.items()}
        assert abs(st["home_qlearner_loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"])
        assert abs(st["home_qlearner_grad_norm"] - norm) <= 1e-5 * norm
        assert_close(L._grad.cpu().numpy(), g_clip, 1e-5, "clipped gradient (%s)" % variant)
        p0 = np.concatenate([v.ravel() for v in list(p_agent.values()) + list(p_mixer.values())]).astype(np.float64)
        p_ref, sq_ref = O.rmsprop_update(p0, g_clip, sq0, a.lr, a.optim_alpha, a.optim_eps)
        p_new = np.concatenate([v.ravel() for v in list(np_params(s.mac.agent).values()) + list(np_params(L.mixer).values())])
        assert_close(p_new, p_ref, 1e-5, "post-step parameters (%s)" % variant)
        assert_close(p_new - p0, p_ref - p0, 1e-3, "parameter update (%s)" % variant)
        assert_close(L.optimiser.flat_sq.cpu().numpy(), sq_ref, 1e-5, "square_avg (%s)" % variant)


@pytest.mark.parametrize("N,B,TT", [(5, 32, 201), (3, 4, 7), (20, 3, 5), (26, 2, 4)])
def test_wire_records_round_trip_bit_exact(N, B, TT):
    """Compact wire form of an episode batch (host <-> device path of a host-resident replay buffer): pack -> (host) ->
    unpack restores every field bit for bit, incl. variable-length episodes; values that do not fit are reported."""
    A, OBS, S = 6 + N, 8 + 8 * N, 16 * N
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    gen = th.Generator().manual_seed(N)
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=True, device=DEV)
    eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, B, TT, preprocess=pre, device=DEV), data, lens)
    wire = eb.to_wire()
    rb, wb = eb._layout.record_bytes, eb.wire_bytes()
    assert wire.shape == (B, wb) and wb < 0.8 * rb                        # at least 20 % fewer bytes on the wire
    host = wire.cpu().pin_memory()                                         # what a host-resident buffer keeps
    out = M.EpisodeBatch(scheme, groups, B, TT, preprocess=pre, device=DEV)
    out._storage.fill_(0xAB)                                               # every field must be overwritten
    out.load_wire(host.to(DEV, non_blocking=True))
    for k, v in eb.data.transition_data.items():
        assert th.equal(out[k], v), k
    assert int(out.max_t_filled()) == int(eb.max_t_filled())
    # the learner sees the same batch: identical loss statistics after one forward
    bad = M.EpisodeBatch(scheme, groups, B, TT, preprocess=pre, device=DEV)
    bad._storage.copy_(eb._storage)
    bad["avail_actions"][0, 0, 0, 1] = 7                                   # not a 0/1 flag
    with pytest.raises(ValueError):
        bad.to_wire()
    bad._storage.copy_(eb._storage)
    bad["actions_onehot"][0, 0, 0].zero_()                                 # a filled step without its one-hot
    with pytest.raises(ValueError):
        bad.to_wire()
