"""Act-select parity: bit-exact action indices against reference fixtures (injected draws) and against torch's own
CUDA RNG stream (Philox mode), plus BasicMAC.forward / select_actions numerics."""
import numpy as np
import pytest
import torch as th
from torch.distributions import Categorical

import ma_league_b200 as M
from ma_league_b200.synthetic import make_args, make_scheme
from oracle import np_oracle as O
from tests.gpu_helpers import to_sd, build_system
from tests.helpers import load_golden, sub, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _selector(**kw):
    return M.EpsilonGreedyActionSelector(make_args(5, 11, 80, device=DEV, **kw))


def test_select_fixture_bit_exact():
    g = load_golden("select_eps_greedy")
    sel = _selector()
    for ci in range(int(g["n_cases"])):
        p = "c%d." % ci
        q = th.from_numpy(g[p + "q"]).to(DEV)
        avail = th.from_numpy(g[p + "avail"]).to(DEV)
        picked, greedy = sel.select(q, avail, int(g[p + "t_env"]), bool(g[p + "test_mode"]),
                                    u=th.from_numpy(g[p + "u"]), e=th.from_numpy(g[p + "e"]))
        assert sel.epsilon == float(g[p + "eps"])
        assert picked.dtype == th.long and greedy.dtype == th.long
        assert np.array_equal(picked.cpu().numpy(), g[p + "picked"]), ci
        assert np.array_equal(greedy.cpu().numpy(), g[p + "greedy"]), ci


def _reference_select_torch(q, avail, eps):
    """The reference's op sequence (action_selectors.py:52-61) on the current device/generator."""
    masked = q.clone()
    masked[avail == 0.0] = -float("inf")
    rnd = th.rand_like(q[:, :, 0])
    pick_random = (rnd < eps).long()
    random_actions = Categorical(avail.float()).sample().long()
    return pick_random * random_actions + (1 - pick_random) * masked.max(dim=2)[1], 1 - pick_random


@pytest.mark.parametrize("bs,N,A", [(1, 5, 11), (32, 5, 11), (128, 10, 16), (1024, 20, 26), (3, 1, 2)])
def test_philox_stream_matches_torch_cuda_generator(bs, N, A):
    gen = th.Generator(device="cpu").manual_seed(bs)
    q = th.randn(bs, N, A, generator=gen).to(DEV)
    avail = (th.rand(bs, N, A, generator=gen) < 0.6).int()
    avail[..., 1] = 1
    avail = avail.to(DEV)
    sel = _selector()
    for eps_t in (0, 30000, 10 ** 7):
        eps = O.epsilon_linear(1.0, 0.05, 50000, eps_t)
        th.manual_seed(1234 + eps_t)
        ref_a, ref_g = _reference_select_torch(q, avail, eps)
        after_ref = th.cuda.default_generators[0].get_offset()
        th.manual_seed(1234 + eps_t)
        a, g = sel.select(q, avail, eps_t)
        assert th.cuda.default_generators[0].get_offset() == after_ref      # same generator advance
        assert th.equal(a, ref_a) and th.equal(g, ref_g)
    # test_mode: epsilon forced to 0 but the stream is still consumed
    th.manual_seed(7)
    ref_a, _ = _reference_select_torch(q, avail, 0.0)
    nxt_ref = th.rand(4, device=DEV)
    th.manual_seed(7)
    a, g = sel.select(q, avail, 0, test_mode=True)
    assert th.equal(a, ref_a) and bool((g == 1).all())
    assert th.equal(th.rand(4, device=DEV), nxt_ref)


def test_select_edge_cases():
    sel = _selector()
    q = th.zeros(2, 3, 4, device=DEV)
    q[0, 0] = th.tensor([1.0, 5.0, 5.0, 0.0])          # tie -> lowest index
    q[0, 1] = th.tensor([9.0, 1.0, 2.0, 3.0])          # best action unavailable
    q[1, 2] = th.tensor([float("nan"), 1.0, 2.0, 3.0])  # NaN wins torch.max
    avail = th.ones(2, 3, 4, dtype=th.int32, device=DEV)
    avail[0, 1, 0] = 0
    u = th.ones(2, 3)
    e = th.ones(6, 4)
    a, g = sel.select(q, avail, 10 ** 7, u=u, e=e)
    ref_a, ref_g = O.eps_greedy_select(q.cpu().numpy(), avail.cpu().numpy(), 0.05, u.numpy(), e.numpy())
    assert a[0, 0] == 1 and a[0, 1] == 3 and a[1, 2] == 0
    assert np.array_equal(a.cpu().numpy()[:1], ref_a[:1])
    avail[1, 1] = 0                                     # no available action: Categorical raises ValueError
    with pytest.raises(ValueError):
        sel.select(q, avail, 0, u=u, e=e)


def test_mac_select_actions_fixture():
    g = load_golden("mac_select_actions")
    bs, TT, N, A, OBS, S = [int(x) for x in g["meta"]]
    s = build_system(N, A, OBS, S, bs, TT, "vdn", True, DEV)
    s.mac.agent.load_state_dict(to_sd(sub(g, "agent."), DEV))
    eb = M.EpisodeBatch(s.scheme, s.groups, bs, TT, preprocess=s.pre, device=DEV)
    for k, v in sub(g, "batch.").items():
        eb.data.transition_data[k].copy_(th.from_numpy(v).to(DEV))
    with pytest.raises(Exception):
        s.mac.forward(eb, 0)                            # HiddenStateNotInitialized
    s.mac.init_hidden(bs)
    for t in range(3):
        acts, greedy = s.mac.select_actions(eb, t_ep=t, t_env=20000 * t, test_mode=False,
                                            u=th.from_numpy(g["t%d.u" % t]), e=th.from_numpy(g["t%d.e" % t]))
        assert s.mac.action_selector.epsilon == float(g["t%d.eps" % t])
        assert_close(s.mac.hidden_states.cpu().numpy(), g["t%d.hidden" % t], 1e-5, "hidden t=%d" % t)
        assert np.array_equal(acts.cpu().numpy(), g["t%d.actions" % t]), t
        assert np.array_equal(greedy.cpu().numpy(), g["t%d.greedy" % t]), t
    # forward() alone reproduces q; a subset `bs` goes through forward + the stand-alone selector
    s.mac.init_hidden(bs)
    for t in range(3):
        q = s.mac.forward(eb, t)
        assert q.shape == (bs, N, A)
        assert_close(q.cpu().numpy(), g["t%d.q" % t], 1e-5, "q t=%d" % t)
    s.mac.init_hidden(bs)
    acts, _ = s.mac.select_actions(eb, 0, 0, bs=[0, 2], u=th.from_numpy(g["t0.u"])[[0, 2]],
                                   e=th.from_numpy(g["t0.e"]).view(bs, N, A)[[0, 2]].reshape(-1, A))
    assert acts.shape == (2, N)
    # a subset of environments sees exactly the actions the full call gave those rows (same draws, same Q-values)
    assert np.array_equal(acts.cpu().numpy(), g["t0.actions"][[0, 2]])
    s.mac.init_hidden(bs)
    acts1, greedy1 = s.mac.select_actions(eb, 0, 0, bs=slice(1, 3), u=th.from_numpy(g["t0.u"])[1:3],
                                          e=th.from_numpy(g["t0.e"]).view(bs, N, A)[1:3].reshape(-1, A))
    assert np.array_equal(acts1.cpu().numpy(), g["t0.actions"][1:3])
    assert np.array_equal(greedy1.cpu().numpy(), g["t0.greedy"][1:3])


def test_agent_and_mixer_modules_standalone():
    """DRQNAgentNetwork.forward(inputs, hidden) and QMixer/VDNMixer.forward against the oracle."""
    N, A, OBS, S, B, T = 4, 10, 9, 13, 3, 5
    s = build_system(N, A, OBS, S, B, T + 1, "qmix", True, DEV)
    gen = th.Generator().manual_seed(0)
    x = th.randn(B * N, OBS + A + N, generator=gen)
    h = th.randn(B * N, 64, generator=gen)
    q, hn = s.mac.agent(x.to(DEV), h.to(DEV))
    p = {k: v.cpu().numpy() for k, v in s.mac.agent.state_dict().items()}
    rq, rh, _ = O.drqn_step(p, x.numpy().astype(np.float64), h.numpy().astype(np.float64))
    assert_close(q.cpu().numpy(), rq, 1e-5, "agent q")
    assert_close(hn.cpu().numpy(), rh, 1e-5, "agent h")
    qs = th.randn(B, T, N, generator=gen)
    st = th.randn(B, T + 1, S, generator=gen)
    out = s.learner.mixer(qs.to(DEV), st.to(DEV)[:, :-1])
    mp = {k: v.cpu().numpy().astype(np.float64) for k, v in s.learner.mixer.state_dict().items()}
    ref, _ = O.qmix_forward(mp, qs.numpy().astype(np.float64), st.numpy().astype(np.float64)[:, :-1])
    assert out.shape == (B, T, 1)
    assert_close(out.cpu().numpy(), ref, 1e-5, "qmix")
    v = M.VDNMixer()(qs.to(DEV), None)
    assert_close(v.cpu().numpy(), qs.numpy().sum(2, keepdims=True), 1e-6, "vdn")


def test_ensemble_mac_matches_oracle_per_agent():
    """EnsembleMAC (ensemble_agent_controller.py:46-57): agents 1 and 3 infer with their own networks, the others with
    the native one; Q-values, per-network hidden states and greedy actions are checked over three steps against the
    numpy oracle's single-network step applied agent by agent."""
    from oracle import np_oracle as O
    from tests.gpu_helpers import np_params, np_batch
    from ma_league_b200.synthetic import make_args, make_scheme, synth_episode_data, fill_episode_batch
    N, A, OBS, S, B, TT = 4, 10, 40, 64, 3, 6
    th.manual_seed(3)
    args = make_args(N, A, S, mixer="vdn", device=DEV)
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, B, TT, preprocess=pre, device=DEV)
    mac = M.mac_REGISTRY["ensemble"](buf.scheme, groups, args)
    donors = {aid: M.mac_REGISTRY["basic"](buf.scheme, groups, args) for aid in (1, 3)}      # differently initialised nets
    mac.load_state_dict(ensemble={aid: d.agent.state_dict() for aid, d in donors.items()})
    assert sorted(mac.ensemble_ids) == [1, 3] and mac.n_native_agents == 2 and mac.native_agents_ids == [0, 2]
    gen = th.Generator().manual_seed(4)
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=False, device=DEV)
    eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, B, TT, preprocess=pre, device=DEV), data, lens)
    nb = np_batch(eb)
    p_nat = np_params(mac.agent)
    p_ens = {aid: np_params(mac.ensemble[aid]) for aid in (1, 3)}
    with pytest.raises(Exception):
        mac.forward(eb, 0)                               # HiddenStateNotInitialized
    mac.init_hidden(B)
    h_nat = np.zeros((B * N, 64), np.float32)
    h_ens = {aid: np.zeros((B, 64), np.float32) for aid in (1, 3)}
    for t in range(3):
        acts, greedy = mac.select_actions(eb, t_ep=t, t_env=0, test_mode=True)
        inp = O.build_inputs(nb["obs"], nb["actions_onehot"], t)                     # [B*N, D_in]
        q_ref, h_nat, _ = O.drqn_step(p_nat, inp, h_nat)
        q_ref = q_ref.reshape(B, N, A).copy()
        for aid in (1, 3):
            qa, h_ens[aid], _ = O.drqn_step(p_ens[aid], inp.reshape(B, N, -1)[:, aid], h_ens[aid])
            q_ref[:, aid] = qa
        masked = np.where(nb["avail_actions"][:, t] == 0, -np.inf, q_ref)
        assert np.array_equal(acts.cpu().numpy(), masked.argmax(-1)), t
        assert bool((greedy == 1).all())
        assert_close(mac.native_hidden_states.reshape(B * N, 64).cpu().numpy(), h_nat, 1e-5, "native hidden t=%d" % t)
        for aid in (1, 3):
            assert_close(mac.ensemble_hidden_states[aid].reshape(B, 64).cpu().numpy(), h_ens[aid], 1e-5, "ensemble hidden")
    # without ensemble members it is BasicMAC
    plain = M.mac_REGISTRY["ensemble"](buf.scheme, groups, args)
    plain.load_state_dict(agent=mac.agent.state_dict())
    base = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    base.load_state(mac)
    plain.init_hidden(B); base.init_hidden(B)
    a1, _ = plain.select_actions(eb, 0, 0, test_mode=True)
    a2, _ = base.select_actions(eb, 0, 0, test_mode=True)
    assert th.equal(a1, a2)


def test_dqn_agent_forward_fixture_and_select_actions():
    """DQNAgentNetwork.forward against the reference fixture (dqn_agent.py:34-37), and BasicMAC with agent="dqn":
    forward / select_actions against the numpy oracle over three steps (greedy and exploring picks, injected draws)."""
    from tests.gpu_helpers import np_params, np_batch
    from ma_league_b200.synthetic import synth_episode_data, fill_episode_batch
    g = load_golden("dqn_agent")
    rows, d_in, A = [int(x) for x in g["meta"]]
    args = make_args(3, A, 10, device=DEV, agent="dqn", batch_size=rows)
    net = M.agent_REGISTRY["dqn"](d_in, args)
    net.load_state_dict(to_sd(sub(g, "agent."), DEV))
    hidden = net.init_hidden()
    assert tuple(hidden.shape) == tuple(int(x) for x in g["hidden_shape"])
    q, h_out = net(th.from_numpy(g["x"]).to(DEV), hidden)
    assert h_out is hidden                                # passed through untouched
    assert_close(q.cpu().numpy(), g["q"], 1e-5, "dqn q vs reference fixture")
    # through the controller
    N, A, OBS, S, B, TT = 4, 10, 40, 64, 5, 6
    s = build_system(N, A, OBS, S, B, TT, "vdn", True, DEV, agent="dqn")
    gen = th.Generator().manual_seed(9)
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=False, device=DEV)
    eb = fill_episode_batch(M.EpisodeBatch(s.scheme, s.groups, B, TT, preprocess=s.pre, device=DEV), data, lens)
    nb, p = np_batch(eb), np_params(s.mac.agent)
    with pytest.raises(Exception):
        s.mac.forward(eb, 0)                              # HiddenStateNotInitialized, as for the recurrent agent
    s.mac.init_hidden(B)
    for t in range(3):
        u, e = th.rand(B, N, generator=gen), th.empty(B * N, A).exponential_(generator=gen)
        q_ref, _, _ = O.dqn_step({k: v.astype(np.float64) for k, v in p.items()}, O.build_inputs(nb["obs"], nb["actions_onehot"], t).astype(np.float64))
        q = s.mac.forward(eb, t)
        assert_close(q.cpu().numpy(), q_ref.reshape(B, N, A), 1e-5, "dqn q t=%d" % t)
        acts, greedy = s.mac.select_actions(eb, t_ep=t, t_env=30000, u=u, e=e)
        ra, rg = O.eps_greedy_select(q_ref.reshape(B, N, A), nb["avail_actions"][:, t], s.mac.action_selector.epsilon, u.numpy(), e.numpy())
        assert np.array_equal(acts.cpu().numpy(), ra) and np.array_equal(greedy.cpu().numpy(), rg)
        assert 0 < int(rg.sum()) < rg.size                # both branches taken
