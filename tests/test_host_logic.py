"""Host-side mirror of the reference API (no GPU): packed EpisodeBatch / ReplayBuffer semantics against the golden
ring fixture, error conventions, registries, epsilon schedule, state_dict keys, no-CPU-fallback behaviour."""
import copy

import numpy as np
import pytest
import torch as th

import ma_league_b200 as M
from ma_league_b200 import _native as nat
from ma_league_b200.components.epsilon_schedules import DecayThenFlatSchedule
from ma_league_b200.exceptions import HiddenStateNotInitialized
from ma_league_b200.flat import ensure_flat
from ma_league_b200.synthetic import make_args, make_scheme
from oracle import np_oracle as O
from tests.helpers import load_golden, sub


def _ring_fixture():
    g = load_golden("replay_ring")
    size, TT, N, A, OBS, S = [int(x) for x in g["meta"]]
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    return g, size, TT, (scheme, groups, pre)


def _batch_from(g, prefix, n, TT, sg):
    scheme, groups, pre = sg
    eb = M.EpisodeBatch(scheme, groups, n, TT, preprocess=pre, device="cpu")
    for k, v in sub(g, prefix).items():
        eb.data.transition_data[k].copy_(th.from_numpy(v))
    return eb


def test_replay_ring_matches_reference_on_host_storage():
    g, size, TT, sg = _ring_fixture()
    scheme, groups, pre = sg
    buf = M.ReplayBuffer(scheme, groups, size, TT, preprocess=pre, device="cpu")
    assert list(buf.data.transition_data.keys()) == ["state", "obs", "actions", "avail_actions", "reward", "terminated",
                                                     "actions_onehot", "filled"]
    np.random.seed(3)
    for i, n in enumerate([3, 3, 3, 5, 1]):
        buf.insert_episode_batch(_batch_from(g, "ins%d." % i, n, TT, sg))
        assert [buf.buffer_index, buf.episodes_in_buffer] == list(g["counters"][i])
        for k, v in buf.data.transition_data.items():
            assert np.array_equal(v.numpy(), g["buf%d.%s" % (i, k)]), (i, k)
        if buf.can_sample(4):
            smp = buf.sample(4)
            assert smp.batch_size == 4 and smp.max_seq_length == TT
            for k, v in smp.data.transition_data.items():
                assert np.array_equal(v.numpy(), g["smp%d.%s" % (i, k)]), (i, k)
            assert int(smp.max_t_filled()) == int(g["smp%d.max_t" % i])
    with pytest.raises(AssertionError):
        M.ReplayBuffer(scheme, groups, 2, TT, preprocess=pre).sample(1)     # replay_buffer.py:47


def test_sample_exact_fill_returns_views():
    g, size, TT, sg = _ring_fixture()
    scheme, groups, pre = sg
    buf = M.ReplayBuffer(scheme, groups, size, TT, preprocess=pre, device="cpu")
    buf.insert_episode_batch(_batch_from(g, "ins0.", 3, TT, sg))
    smp = buf.sample(3)                                # episodes_in_buffer == batch_size -> self[:3] (views)
    assert smp["obs"].data_ptr() == buf["obs"].data_ptr()


def test_episode_batch_indexing_semantics():
    g, size, TT, sg = _ring_fixture()
    scheme, groups, pre = sg
    eb = _batch_from(g, "ins3.", 5, TT, sg)
    assert eb["obs"].shape == (5, TT, 2, 3) and eb["filled"].dtype == th.long and eb["terminated"].dtype == th.uint8
    v = eb[1:3, :2]
    assert v.batch_size == 2 and v.max_seq_length == 2
    assert v["state"].data_ptr() == eb["state"][1:3, :2].data_ptr()          # slices alias the parent
    c = eb[[0, 4]]
    assert c.batch_size == 2 and c["obs"].data_ptr() != eb["obs"].data_ptr()  # index arrays copy
    assert th.equal(c["obs"], eb["obs"][[0, 4]])
    assert eb[2].batch_size == 1
    sb = eb[("obs", "actions")]
    assert set(sb.scheme.keys()) == {"obs", "actions"} and sb["obs"].data_ptr() == eb["obs"].data_ptr()
    with pytest.raises(IndexError):
        eb[0, [0, 2]]                                                        # episode_batch.py:43-44
    with pytest.raises(ValueError):
        eb["nope"]                                                           # :206
    with pytest.raises(KeyError):
        eb[("obs", "nope")]                                                  # :217
    with pytest.raises(KeyError):
        eb.update({"nope": th.zeros(1)})                                     # :173
    with pytest.raises(ValueError):
        eb.update({"state": th.zeros(5, TT, 4)})                             # unsafe reshape, :11


def test_update_marks_filled_and_applies_onehot():
    N, A, OBS, S, TT = 2, 4, 3, 5, 5
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    eb = M.EpisodeBatch(scheme, groups, 1, TT, preprocess=pre)
    eb.update({"state": [[0.5] * S], "avail_actions": [[[1, 0, 1, 1], [1, 1, 0, 0]]], "obs": [[[1.0] * OBS] * N]}, ts=0)
    eb.update({"actions": [[[2], [1]]], "reward": [(0.25,)], "terminated": [(False,)]}, ts=0)
    assert eb["filled"][0, :, 0].tolist() == [1, 0, 0, 0, 0]
    assert np.array_equal(eb["actions_onehot"][0, 0].numpy(), O.onehot(np.array([[2], [1]]), A))
    assert eb["avail_actions"].dtype == th.int32 and eb["actions"].dtype == th.long
    eb.to("cpu")
    assert eb["obs"].shape == (1, TT, N, OBS)


def test_records_are_aligned_and_views_alias_storage():
    N, A, OBS, S, TT = 5, 11, 48, 80, 201
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    eb = M.EpisodeBatch(scheme, groups, 3, TT, preprocess=pre)
    rb = eb._layout.record_bytes
    assert rb % 128 == 0 and rb >= 1773 * TT          # 1773 B per stored step at N=5
    for key, (off, shape, dtype, const) in eb._layout.fields.items():
        assert off % 128 == 0
    eb["reward"][2, 7, 0] = 3.5
    assert eb._storage.view(3, rb)[2].view(th.float32)[eb._layout.fields["reward"][0] // 4 + 7] == 3.5


def test_registries_and_schedule():
    assert set(M.mac_REGISTRY) == {"basic", "ensemble"} and set(M.learner_REGISTRY) == {"q"}
    assert set(M.agent_REGISTRY) == {"rnn", "dqn"} and set(M.action_REGISTRY) == {"epsilon_greedy"}
    s = DecayThenFlatSchedule(1.0, 0.05, 50000, decay="linear")
    for t in (0, 1, 25000, 49999, 50000, 10 ** 7):
        assert s.eval(t) == O.epsilon_linear(1.0, 0.05, 50000, t)


def _cpu_learner(mixer="qmix"):
    N, A, OBS, S = 3, 9, 12, 14
    args = make_args(N, A, S, mixer=mixer, device="cpu")
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, 4, 6, preprocess=pre)
    mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    learner = M.learner_REGISTRY["q"](mac, buf.scheme, None, args, name="home")
    learner.build_optimizer()
    return args, buf, mac, learner


def test_state_dict_keys_and_flat_aliasing_survive_deepcopy_and_load():
    args, buf, mac, learner = _cpu_learner()
    assert list(mac.agent.state_dict().keys()) == list(O.agent_param_shapes(24, 9).keys())
    assert list(learner.mixer.state_dict().keys()) == list(O.qmix_param_shapes(14, 3).keys())
    for k, shp in O.qmix_param_shapes(14, 3).items():
        assert tuple(learner.mixer.state_dict()[k].shape) == shp
    assert learner.name == "home_qlearner_" and mac.input_shape == 12 + 9 + 3
    flat = ensure_flat(mac.agent)
    assert flat.numel() == nat.lib().mal_agent_param_count(24, 9)
    assert mac.agent.fc1.weight.data_ptr() == flat.data_ptr()
    tflat = ensure_flat(learner.target_mac.agent)                 # deepcopy broke the aliasing; it is repaired
    assert tflat.data_ptr() != flat.data_ptr()
    assert learner.target_mac.agent.fc2.bias.data_ptr() == tflat.data_ptr() + 4 * (tflat.numel() - 9)
    sd = {k: th.randn_like(v) for k, v in mac.agent.state_dict().items()}
    mac.load_state_dict(sd)                                       # in-place: views stay attached
    assert ensure_flat(mac.agent).data_ptr() == flat.data_ptr()
    assert th.equal(flat[:64 * 24].view(64, 24), sd["fc1.weight"])
    clone = copy.deepcopy(mac)
    assert th.equal(clone.agent.fc1.weight, mac.agent.fc1.weight)
    mac.agent.trained_steps = 5
    mac.update_trained_steps(7)
    assert mac.agent.trained_steps == 12


def test_optimizer_state_dict_is_torch_rmsprop_format():
    args, buf, mac, learner = _cpu_learner("vdn")
    sd = learner.optimiser.state_dict()
    g = sd["param_groups"][0]
    assert g["lr"] == 5e-4 and g["alpha"] == 0.99 and g["eps"] == 1e-5 and g["momentum"] == 0 and not g["centered"]
    assert set(sd["state"][0].keys()) == {"step", "square_avg"}
    ref = th.optim.RMSprop(learner.parameters(), lr=5e-4, alpha=0.99, eps=1e-5)
    ref.load_state_dict(sd)                                       # loads into the stock optimiser
    sd["state"][0]["square_avg"] = th.full_like(sd["state"][0]["square_avg"], 2.0)
    learner.optimiser.load_state_dict(sd)
    assert float(learner.optimiser.flat_sq[0]) == 2.0
    with pytest.raises(ValueError):
        M.QLearner(mac, buf.scheme, None, make_args(3, 9, 14, mixer="foo", device="cpu"))   # q_learner.py:24


def test_no_cpu_fallback():
    args, buf, mac, learner = _cpu_learner()
    with pytest.raises(HiddenStateNotInitialized):
        mac.forward(buf, 0)
    mac.init_hidden(4)
    assert mac.hidden_states.shape == (4, 3, 64)
    with pytest.raises(nat.MalError):
        mac.forward(buf, 0)                                       # CPU tensors -> raise, never compute on the host
    with pytest.raises(nat.MalError):
        learner.train(buf, 0, 0)
    with pytest.raises(nat.MalError):
        mac.action_selector.select(th.zeros(1, 3, 9), th.ones(1, 3, 9, dtype=th.int32), 0)


def _balanced_pieces(worker, D, chains, TT):
    """The pieces of one worker of k_gru_fwd9's balanced mode, in the order the kernel runs them (csrc/gru_rec.cuh: the
    `for (int pend = hi; pend > lo;)` loop): (chain, t_begin, t_end)."""
    lo = D * worker
    hi = min(lo + D, chains * TT)
    out, pend = [], hi
    while pend > lo:
        chain = (pend - 1) // TT
        pbeg = max(lo, chain * TT)
        out.append((chain, pbeg - chain * TT, pend - chain * TT))
        pend = pbeg
    return out


@pytest.mark.parametrize("R,TT,sms", [(160, 201, 148), (160, 61, 148), (149, 201, 148), (222, 201, 148), (200, 33, 148),
                                      (160, 201, 132), (75, 40, 66)])
def test_balanced_recurrence_schedule_properties(R, TT, sms):
    """Host restatement of the balanced forward-recurrence schedule (launch_gru_fwd in csrc/mal_b200.cu picks it for
    2 * SMs < chains <= 3 * SMs): every chain-step is run exactly once, a chain is split between at most two workers, the
    head piece (the one that signals) is the FIRST thing its worker runs and the tail piece (the one that waits) the LAST
    thing its worker runs, and a worker with a full share never reaches its tail before the head has ended -- so, when all
    workers advance at the same pace, the hand-over flag is set before it is polled and nobody idles; the makespan is D."""
    chains, workers = 2 * R, 2 * sms
    assert workers < chains <= 3 * sms                      # the launcher's window
    D = -(-chains * TT // workers)
    assert D >= TT                                          # a chain cannot span three workers
    covered = np.zeros((chains, TT), np.int32)
    head_end, tails = {}, {}
    for w in range(workers):
        pieces = _balanced_pieces(w, D, chains, TT)
        clock = 0
        for k, (c, tb, te) in enumerate(pieces):
            assert 0 <= tb < te <= TT
            covered[c, tb:te] += 1
            if te < TT:                                     # head of a split chain: signals when done
                assert tb == 0 and k == 0                   # ... and is the first piece of its worker (starts at time 0)
                head_end[c] = te
            if tb > 0:                                      # tail of a split chain: waits for the head
                assert te == TT and k == len(pieces) - 1    # ... and is the last piece of its worker
                tails[c] = (w, clock, te - tb, sum(e - b for _, b, e in pieces))
            clock += te - tb
        assert clock <= D
    assert (covered == 1).all()
    assert set(head_end) == set(tails)                      # every split chain has exactly one producer and one consumer
    for c, (w, start, length, total) in tails.items():
        if total == D:                                      # a full worker never waits: the head ended before it gets there
            assert start >= head_end[c], (c, w, head_end[c], start)
        # a short worker (the clipped end of the sequence) may poll its flag, but still ends within the makespan
        assert max(start, head_end[c]) + length <= D
