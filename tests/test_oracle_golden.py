"""Pin oracle/np_oracle.py against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import load_golden, sub, assert_close, rel_err

LEARNER_CASES = ["learner_qmix_3v3", "learner_vdn_2v2", "learner_qmix_nodouble", "learner_qmix_dqn"]


def _run(g, dtype):
    B, TT, N, A, OBS, S, is_qmix, double_q, layers, steps = [int(x) for x in g["meta"]]
    gamma, lr, alpha, eps, clip = [float(x) for x in g["hyper"]]
    batch = sub(g, "batch.")
    res = O.learner_forward_backward(sub(g, "agent0."), sub(g, "tagent0."),
                                     sub(g, "mixer0.") if is_qmix else None,
                                     sub(g, "tmixer0.") if is_qmix else None, batch,
                                     mixer="qmix" if is_qmix else "vdn", double_q=bool(double_q), gamma=gamma,
                                     dtype=dtype)
    return res, (gamma, lr, alpha, eps, clip, steps, is_qmix, double_q)


@pytest.mark.parametrize("case", LEARNER_CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forward_and_grads_match_reference(case, dtype):
    g = load_golden(case)
    res, _ = _run(g, dtype)
    tol = 1e-5   # north_star tolerance; the fp64 run shows the residual is the reference's own fp32 rounding
    for k in ["mac_out", "target_mac_out", "chosen", "target_max", "q_tot", "target_q_tot"]:
        assert_close(res[k], g[k], tol, k)
    assert np.array_equal(res["argmax"], g["argmax"])
    assert abs(float(res["loss"]) - g["stat.loss"]) <= tol * abs(g["stat.loss"])
    for k, v in res["agent_grads"].items():
        assert_close(v, g["grad.agent." + k], tol, "grad " + k)
    for k, v in res["mixer_grads"].items():
        assert_close(v, g["grad.mixer." + k], tol, "grad mixer " + k)
    st = res["stats"]
    assert st["trained_steps"] * 1 >= 1
    for k in ["td_error_abs", "q_taken_mean", "target_mean"]:
        assert abs(float(st[k]) - g["stat." + k]) <= 1e-5 * max(1.0, abs(g["stat." + k])), k


@pytest.mark.parametrize("case", LEARNER_CASES)
def test_clip_and_rmsprop_two_steps_match_reference(case):
    g = load_golden(case)
    _, (gamma, lr, alpha, eps, clip, steps, is_qmix, double_q) = _run(g, np.float32)
    ap, tp = sub(g, "agent0."), sub(g, "tagent0.")
    mp, tmp = (sub(g, "mixer0."), sub(g, "tmixer0.")) if is_qmix else (None, None)
    batch = sub(g, "batch.")
    sq_a = {k: np.zeros_like(v) for k, v in ap.items()}
    sq_m = {k: np.zeros_like(v) for k, v in mp.items()} if is_qmix else {}
    for step in range(steps):
        res = O.learner_forward_backward(ap, tp, mp, tmp, batch, mixer="qmix" if is_qmix else "vdn",
                                         double_q=bool(double_q), gamma=gamma, dtype=np.float32)
        keys_a, keys_m = list(ap.keys()), list(mp.keys()) if is_qmix else []
        grads = [res["agent_grads"][k] for k in keys_a] + [res["mixer_grads"][k] for k in keys_m]
        norm, clipped = O.clip_grad_norm(grads, clip)
        if step == 0:
            assert abs(float(norm) - float(g["grad_norm"])) <= 1e-5 * float(g["grad_norm"])
        for k, gr in zip(keys_a, clipped[:len(keys_a)]):
            ap[k], sq_a[k] = O.rmsprop_update(ap[k], gr, sq_a[k], lr, alpha, eps)
        for k, gr in zip(keys_m, clipped[len(keys_a):]):
            mp[k], sq_m[k] = O.rmsprop_update(mp[k], gr, sq_m[k], lr, alpha, eps)
    # RMSprop's first steps are sign-like (g/sqrt(0.01 g^2)): compare the parameter DELTA norm-wise, loosely
    for k in ap:
        assert_close(ap[k], g["agentK." + k], 1e-5, "post-step " + k)
        assert rel_err(sq_a[k], g["sqavg.agent." + k]) < 1e-4, k
    for k in (mp or {}):
        assert_close(mp[k], g["mixerK." + k], 1e-5, "post-step mixer " + k)
    assert int(g["trained_steps"]) == steps * res["stats"]["trained_steps"]


def test_eps_greedy_select_bit_exact():
    g = load_golden("select_eps_greedy")
    for ci in range(int(g["n_cases"])):
        p = "c%d." % ci
        eps = float(g[p + "eps"])
        if int(g[p + "test_mode"]):
            assert eps == 0.0
        else:
            assert eps == O.epsilon_linear(1.0, 0.05, 50000, int(g[p + "t_env"]))
        picked, greedy = O.eps_greedy_select(g[p + "q"], g[p + "avail"], eps, g[p + "u"], g[p + "e"])
        assert np.array_equal(picked, g[p + "picked"]), ci
        assert np.array_equal(greedy, g[p + "greedy"]), ci


def test_mac_select_actions_steps():
    g = load_golden("mac_select_actions")
    bs, TT, N, A, OBS, S = [int(x) for x in g["meta"]]
    p = sub(g, "agent.")
    batch = sub(g, "batch.")
    h = np.zeros((bs * N, 64), np.float32)
    for t in range(3):
        q, h, _ = O.drqn_step(p, O.build_inputs(batch["obs"], batch["actions_onehot"], t), h)
        assert_close(q.reshape(bs, N, A), g["t%d.q" % t], 1e-5, "q t=%d" % t)
        assert_close(h, g["t%d.hidden" % t], 1e-5, "h t=%d" % t)
        # select on the reference's own q so that index equality is exact
        picked, greedy = O.eps_greedy_select(g["t%d.q" % t], batch["avail_actions"][:, t], float(g["t%d.eps" % t]),
                                             g["t%d.u" % t], g["t%d.e" % t])
        assert np.array_equal(picked, g["t%d.actions" % t])
        assert np.array_equal(greedy, g["t%d.greedy" % t])


def test_ring_buffer_indices():
    g = load_golden("replay_ring")
    size = int(g["meta"][0])
    idx, filled = 0, 0
    buf = None
    for i, n in enumerate([3, 3, 3, 5, 1]):
        slots, idx, filled = O.ring_insert_slots(idx, filled, size, n)
        assert [idx, filled] == list(g["counters"][i])
        ins = sub(g, "ins%d." % i)
        if buf is None:
            buf = {k: np.zeros((size,) + v.shape[1:], v.dtype) for k, v in ins.items()}
        for k, v in ins.items():
            buf[k][slots] = v
        for k in buf:
            assert np.array_equal(buf[k], g["buf%d.%s" % (i, k)]), (i, k)
        if ("smp%d.ids" % i) in g:
            ids = g["smp%d.ids" % i]
            for k in buf:
                assert np.array_equal(buf[k][ids], g["smp%d.%s" % (i, k)])
            assert O.max_t_filled(buf["filled"][ids]) == int(g["smp%d.max_t" % i])


def test_onehot():
    a = np.array([[[2], [0]]])
    assert np.array_equal(O.onehot(a, 3), np.array([[[0, 0, 1], [1, 0, 0]]], np.float32))


# ---- oracle/torch_port.py: the CPU port used as the baseline arm issues the reference's op sequence
@pytest.mark.parametrize("case", LEARNER_CASES)
def test_torch_port_two_steps_match_reference(case):
    import torch as th
    from oracle import torch_port as TP
    g = load_golden(case)
    B, TT, N, A, OBS, S, is_qmix, double_q, layers, steps = [int(x) for x in g["meta"]]
    gamma, lr, alpha, eps, clip = [float(x) for x in g["hyper"]]
    L = TP.TorchPortLearner(sub(g, "agent0."), sub(g, "tagent0."), sub(g, "mixer0."), sub(g, "tmixer0."),
                            mixer="qmix" if is_qmix else "vdn", double_q=bool(double_q), gamma=gamma, lr=lr,
                            alpha=alpha, eps=eps, clip=clip)
    batch = {k: th.from_numpy(v.copy()) for k, v in sub(g, "batch.").items()}
    for i in range(steps):
        out = L.train(batch)
        if i == 0:
            assert abs(out["loss"] - g["stat.loss"]) <= 1e-6 * abs(g["stat.loss"])
            assert abs(out["grad_norm"] - float(g["grad_norm"])) <= 1e-6 * float(g["grad_norm"])
    for k, v in L.ap.items():
        assert np.array_equal(v.detach().numpy(), g["agentK." + k]), k      # same ops -> bit-identical on this host
    for k, v in L.mp.items():
        assert np.array_equal(v.detach().numpy(), g["mixerK." + k]), k


def test_dqn_agent_forward_matches_reference():
    """DQNAgentNetwork.forward (marl/modules/agents/dqn_agent.py:34-37) and its state_dict layout."""
    g = load_golden("dqn_agent")
    rows, d_in, A = [int(x) for x in g["meta"]]
    p = sub(g, "agent.")
    assert list(p) == list(O.dqn_agent_param_shapes(d_in, A)) and all(p[k].shape == v for k, v in O.dqn_agent_param_shapes(d_in, A).items())
    q, h, _ = O.dqn_step(p, g["x"])
    assert h is None and rel_err(q, g["q"]) < 1e-6
    assert list(g["hidden_shape"]) == [rows, 1, 1]               # dqn_agent.py:27-32: a placeholder, passed through
