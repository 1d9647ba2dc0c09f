"""Parity of the CUDA learner step (through the C ABI) with the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch as th

from oracle import np_oracle as O
from tests.gpu_helpers import system_from_golden, seeded_system, np_params, np_batch, split_grad
from tests.helpers import load_golden, assert_close, rel_err, sub

pytestmark = pytest.mark.gpu
CASES = ["learner_qmix_3v3", "learner_vdn_2v2", "learner_qmix_nodouble", "learner_qmix_dqn"]
TOL = 1e-5   # north_star: 1e-5 relative (fp32) on Q-values, mixer outputs, loss and gradients


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference_fixture(case):
    g = load_golden(case)
    s = system_from_golden(g)
    s.learner.save_q = True
    s.learner.forward_only(s.batch)
    it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
    for k in ["mac_out", "target_mac_out", "chosen", "target_max", "q_tot", "target_q_tot"]:
        assert_close(it[k], g[k], TOL, k)
    assert np.array_equal(it["argmax"].astype(np.int64), g["argmax"])          # bit-exact indices
    sc = it["scalars"]
    assert abs(sc[1] - g["stat.loss"]) <= TOL * abs(g["stat.loss"])
    assert abs(sc[2] - g["stat.td_error_abs"]) <= TOL * max(1, abs(g["stat.td_error_abs"]))
    assert abs(sc[3] - g["stat.q_taken_mean"]) <= TOL * max(1, abs(g["stat.q_taken_mean"]))
    assert abs(sc[4] - g["stat.target_mean"]) <= TOL * max(1, abs(g["stat.target_mean"]))


@pytest.mark.parametrize("case", CASES)
def test_gradients_match_reference_fixture(case):
    g = load_golden(case)
    s = system_from_golden(g)
    grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
    for k, v in grads.items():
        assert_close(v, g["grad." + k], TOL, "grad " + k)


@pytest.mark.parametrize("case", CASES)
def test_train_steps_match_reference_fixture(case):
    g = load_golden(case)
    s = system_from_golden(g)
    s.learner.args.learner_log_interval = 0
    for i in range(s.steps):
        s.learner.train(s.batch, t_env=i, episode_num=i)
        if i == 0:
            st = {k: v[0] for k, v in s.logger.stats.items()}
            assert abs(st["home_qlearner_grad_norm"] - float(g["grad_norm"])) <= TOL * float(g["grad_norm"])
            assert abs(st["home_qlearner_loss"] - g["stat.loss"]) <= TOL * abs(g["stat.loss"])
            s.learner.args.learner_log_interval = 10 ** 9
    for k, v in np_params(s.mac.agent).items():
        assert_close(v, g["agentK." + k], TOL, "post-step " + k)
    if "mixerK.V.2.bias" in g:
        for k, v in np_params(s.learner.mixer).items():
            assert_close(v, g["mixerK." + k], TOL, "post-step mixer " + k)
    sq = s.learner.optimiser.state_dict()["state"]
    names = ["agent." + k for k, _ in s.mac.agent.named_parameters()]
    for i, n in enumerate(names):
        assert rel_err(sq[i]["square_avg"].cpu().numpy(), g["sqavg." + n]) < 1e-4, n
    assert s.mac.agent.trained_steps == int(g["trained_steps"])
    assert all(p.grad is not None for p in s.learner.parameters())


def _oracle_run(s, mixer, double_q, dtype=np.float64):
    L = s.learner
    return O.learner_forward_backward(np_params(s.mac.agent), np_params(L.target_mac.agent),
                                      np_params(L.mixer) if mixer == "qmix" else None,
                                      np_params(L.target_mixer) if mixer == "qmix" else None, np_batch(s.batch),
                                      mixer=mixer, double_q=double_q, gamma=s.args.gamma, dtype=dtype)


@pytest.mark.parametrize("N,B,TT,mixer,double_q,layers", [
    (5, 8, 21, "qmix", True, 2), (5, 8, 21, "vdn", True, 2), (3, 5, 12, "qmix", False, 1), (10, 3, 7, "qmix", True, 2),
    (2, 1, 2, "qmix", True, 2), (20, 2, 5, "qmix", True, 2),
    (26, 1, 3, "qmix", True, 2),      # n_actions = 32 = MAL_MAX_ACTIONS, widest agent input (246 columns)
    (1, 2, 3, "vdn", False, 2),       # a single agent
])
def test_against_oracle_seeded(N, B, TT, mixer, double_q, layers):
    s = seeded_system(N, B, TT, mixer, double_q, seed=N + B, hypernet_layers=layers)
    s.learner.save_q = True
    grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
    it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
    ref = _oracle_run(s, mixer, double_q)
    assert_close(it["mac_out"], ref["mac_out"], TOL, "mac_out")
    assert_close(it["target_mac_out"], ref["target_mac_out"], TOL, "target_mac_out")
    assert_close(it["hout"], ref["hout"], TOL, "hidden states")
    assert_close(it["chosen"], ref["chosen"], TOL, "chosen")
    # near-ties of the fp32 argmax may legitimately flip; none are expected on N(0,1) data
    assert np.array_equal(it["argmax"].astype(np.int64), ref["argmax"])
    assert_close(it["target_max"], ref["target_max"], TOL, "target_max")
    assert_close(it["mask"], ref["mask"], 0, "mask")
    assert_close(it["q_tot"], ref["q_tot"], TOL, "q_tot")
    assert_close(it["targets"], ref["targets"], TOL, "targets")
    assert abs(it["scalars"][1] - ref["loss"]) <= TOL * abs(ref["loss"])
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)
    for k, v in ref["mixer_grads"].items():
        assert_close(grads["mixer." + k], v, TOL, "grad mixer " + k)
    assert int(it["scalars"][6:7].view(np.int32)[0]) == ref["stats"]["trained_steps"]


def test_full_size_metric_config_against_oracle():
    """BASELINE.json metric shape: QMIX, B=32, T=200 (201 stored steps), 5v5, double-Q."""
    s = seeded_system(5, 32, 201, "qmix", True, seed=11)
    grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
    it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
    ref = _oracle_run(s, "qmix", True, dtype=np.float64)
    assert_close(it["hout"], ref["hout"], TOL, "hidden states")
    assert_close(it["q_tot"], ref["q_tot"], TOL, "q_tot")
    flips = int((it["argmax"].astype(np.int64) != ref["argmax"]).sum())
    assert flips == 0, "%d argmax flips" % flips
    assert abs(it["scalars"][1] - ref["loss"]) <= TOL * abs(ref["loss"])
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)
    for k, v in ref["mixer_grads"].items():
        assert_close(grads["mixer." + k], v, TOL, "grad mixer " + k)


@pytest.mark.parametrize("mode", [2, 0])
def test_both_tensor_core_gemm_kernels_inside_the_learner(mode):
    """The launch heuristics pick the warp-specialised pipelined GEMM and the tensor-core weight-gradient reductions
    only for long runs (10v10 / 20v20 batches); force each variant (2: pipelined GEMM + tcgen05 reductions, 0: one
    tile at a time + FFMA reductions) through a whole forward/backward at the metric's shape and hold both to the
    oracle."""
    from ma_league_b200 import _native as nat
    s = seeded_system(5, 32, 201, "qmix", True, seed=13)
    ref = _oracle_run(s, "qmix", True, dtype=np.float64)
    nat.check(nat.lib().mal_set_option(b"tc_pipelined", mode), "mal_set_option")
    nat.check(nat.lib().mal_set_option(b"reduce_tc", mode), "mal_set_option")     # 2: tcgen05 split-M reductions, 0: FFMA
    try:
        grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
        it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
    finally:
        nat.check(nat.lib().mal_set_option(b"tc_pipelined", 1), "mal_set_option")
        nat.check(nat.lib().mal_set_option(b"reduce_tc", 1), "mal_set_option")
    assert_close(it["q_tot"], ref["q_tot"], TOL, "q_tot")
    assert int((it["argmax"].astype(np.int64) != ref["argmax"]).sum()) == 0
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)
    for k, v in ref["mixer_grads"].items():
        assert_close(grads["mixer." + k], v, TOL, "grad mixer " + k)


@pytest.mark.parametrize("fc1_fused", [1, 0])
@pytest.mark.parametrize("N,B,TT", [(10, 4, 9), (20, 3, 6), (26, 2, 4)])
def test_tensor_core_reductions_forced_on_small_wide_problems(N, B, TT, fc1_fused):
    """fc1 weight gradients with d_in > 64 take the transposed (SWAP) orientation of k_reduce_tc, which the launch
    heuristic only selects for long row chunks: force the tensor-core reductions and the pipelined GEMM on small
    batches of the wide configs and hold every gradient to the oracle."""
    from ma_league_b200 import _native as nat
    s = seeded_system(N, B, TT, "qmix", True, seed=N + TT)
    ref = _oracle_run(s, "qmix", True, dtype=np.float64)
    nat.check(nat.lib().mal_set_option(b"reduce_tc", 2), "mal_set_option")
    nat.check(nat.lib().mal_set_option(b"tc_pipelined", 2), "mal_set_option")
    nat.check(nat.lib().mal_set_option(b"fc1_fused", fc1_fused), "mal_set_option")   # 1: k_reduce_fc1 (full input width per CTA), 0: SWAP tiles of k_reduce_tc3
    n_fc1 = nat.lib().mal_stat(b"reduce_fc1")
    try:
        grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
    finally:
        nat.check(nat.lib().mal_set_option(b"reduce_tc", 1), "mal_set_option")
        nat.check(nat.lib().mal_set_option(b"tc_pipelined", 1), "mal_set_option")
        nat.check(nat.lib().mal_set_option(b"fc1_fused", 1), "mal_set_option")
    d_in = s.mac.agent.fc1.weight.shape[1]
    assert (nat.lib().mal_stat(b"reduce_fc1") - n_fc1 >= 1) == bool(fc1_fused and d_in <= 256), d_in
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)
    for k, v in ref["mixer_grads"].items():
        assert_close(grads["mixer." + k], v, TOL, "grad mixer " + k)


def test_cuda_graph_replay_equals_eager_launches():
    """train() replays a captured CUDA graph from the third sighting of a batch on; parameters, optimiser state and
    logged statistics must be bit-identical to eager launches, including when two batches alternate."""
    def run(graphs):
        s = seeded_system(3, 4, 9, "qmix", True, seed=21, learner_log_interval=0)
        s.learner.use_graphs = graphs
        s2 = seeded_system(3, 4, 9, "qmix", True, seed=22)          # a second batch at other addresses
        for i in range(8):
            s.learner.train(s.batch if i % 2 == 0 else s2.batch, t_env=i, episode_num=i)
        th.cuda.synchronize()
        if graphs:
            assert sum(1 for v in s.learner._graphs.values() if v) == 2
        return np_params(s.mac.agent), np_params(s.learner.mixer), s.learner.optimiser.flat_sq.cpu().numpy(), \
            {k: v[0] for k, v in s.logger.stats.items()}, s.mac.agent.trained_steps
    a, b = run(True), run(False)
    for x, y in zip(a[:2], b[:2]):
        for k in x:
            assert np.array_equal(x[k], y[k]), k
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3] and a[4] == b[4]


@pytest.mark.parametrize("graphs", [False, True])
def test_programmatic_dependent_launch_does_not_change_results(graphs):
    """The main kernel chain is launched with programmatic-serialization edges (kernels start under their predecessor's
    tail and wait before touching its products): parameters, optimiser state and statistics must be bit-identical to
    fully serialised launches, eagerly and through the captured graph, at the metric's size (every kernel many CTAs)."""
    from ma_league_b200 import _native as nat

    def run(pdl):
        nat.check(nat.lib().mal_set_option(b"pdl", pdl), "mal_set_option")
        try:
            s = seeded_system(5, 32, 201, "qmix", True, seed=31, learner_log_interval=0)
            s.learner.use_graphs = graphs
            for i in range(4):
                s.learner.train(s.batch, t_env=i, episode_num=i)
            th.cuda.synchronize()
            return np_params(s.mac.agent), np_params(s.learner.mixer), s.learner.optimiser.flat_sq.cpu().numpy(), \
                {k: v[0] for k, v in s.logger.stats.items()}
        finally:
            nat.check(nat.lib().mal_set_option(b"pdl", 1), "mal_set_option")
    a, b = run(1), run(0)
    for x, y in zip(a[:2], b[:2]):
        for k in x:
            assert np.array_equal(x[k], y[k]), k
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3]


def test_properties_and_determinism_at_full_size():
    s = seeded_system(5, 32, 201, "qmix", True, seed=5)
    g1 = s.learner.forward_backward(s.batch).clone()
    g2 = s.learner.forward_backward(s.batch).clone()
    assert th.equal(g1, g2)                                    # fixed summation order: bit-reproducible
    # masked-out padding must not influence anything: scramble data after each episode's end
    pad = (s.batch["filled"][..., 0] == 0)
    assert bool(pad.any())
    s.batch["obs"][pad] = 7.0
    s.batch["state"][pad] = -3.0
    s.batch["reward"][pad] = 123.0
    g3 = s.learner.forward_backward(s.batch)
    assert rel_err(g3.cpu().numpy(), g1.cpu().numpy()) < 1e-6
    # linearity of the loss gradient in the reward scale is NOT expected (targets), but gamma = 0 & zero reward => td = q
    s2 = seeded_system(5, 4, 9, "vdn", True, seed=6, gamma=0.0)
    s2.batch["reward"].zero_()
    s2.learner.forward_only(s2.batch)
    it = s2.learner.intermediates(s2.batch)
    assert th.allclose(it["td"], it["q_tot"])
    assert th.allclose(it["q_tot"].squeeze(-1), it["chosen"].sum(-1), atol=1e-6)


def test_target_update_and_truncated_views():
    s = seeded_system(3, 6, 15, "qmix", True, seed=9)
    L = s.learner
    L.args.target_update_interval = 1
    before = np_params(L.target_mac.agent)
    view = s.batch[:, :int(s.batch.max_t_filled())]          # ma_experiment.py:235-236
    assert view.max_seq_length == 15
    short = s.batch[1:5, :9]                                  # strided views (batch + time slices) feed the kernels
    L.train(short, t_env=0, episode_num=1)
    assert any("Updated" in m for m in s.logger.infos)
    after = np_params(L.target_mac.agent)
    online = np_params(s.mac.agent)
    for k in after:
        assert np.array_equal(after[k], online[k]) and not np.array_equal(after[k], before[k])
    for k, v in np_params(L.target_mixer).items():
        assert np.array_equal(v, np_params(L.mixer)[k])
    # the strided-view step equals the same step on a compact copy
    s2 = seeded_system(3, 6, 15, "qmix", True, seed=9)
    ref = O.learner_forward_backward(np_params(s2.mac.agent), np_params(s2.learner.target_mac.agent),
                                     np_params(s2.learner.mixer), np_params(s2.learner.target_mixer),
                                     {k: v[1:5, :9] for k, v in np_batch(s2.batch).items()}, mixer="qmix",
                                     double_q=True, gamma=0.99, dtype=np.float64)
    grads = split_grad(s2.learner.forward_backward(s2.batch[1:5, :9]), s2.learner)
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)


def test_checkpoint_roundtrip(tmp_path):
    s = seeded_system(3, 4, 8, "qmix", True, seed=2)
    s.learner.train(s.batch, 0, 0)
    s.learner.save_models(str(tmp_path))
    for f in ("home_qlearner_agent.th", "home_qlearner_mixer.th", "home_qlearner_opt.th"):   # q_learner.py:133-137
        assert (tmp_path / f).exists()
    s2 = seeded_system(3, 4, 8, "qmix", True, seed=3)
    s2.learner.load_models(str(tmp_path))
    for k, v in np_params(s.mac.agent).items():
        assert np.array_equal(v, np_params(s2.mac.agent)[k])
        assert np.array_equal(v, np_params(s2.learner.target_mac.agent)[k])
    assert th.equal(s.learner.optimiser.flat_sq, s2.learner.optimiser.flat_sq)
    s.learner.train(s.batch, 1, 1)
    s2.learner.target_mixer.load_state_dict(s.learner.target_mixer.state_dict())
    s2.learner.target_mac.load_state(s.learner.target_mac)
    s2.learner.train(s.batch, 1, 1)
    for k, v in np_params(s.mac.agent).items():
        assert np.array_equal(v, np_params(s2.mac.agent)[k])


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at their REAL shape with the DEFAULT launch heuristics (nothing forced): the kernels the bench
# lines of these workloads actually take (k_linear_tc2, k_reduce_tc incl. SWAP, pipelined k_agent_in_tc at d_in > 64)
# are held to the fp64 oracle, and the test asserts which flavours ran (mal_stat counters).
# ------------------------------------------------------------------------------------------------------------------
_FLAVOURS = ("linear_tc2", "linear_tc", "reduce_tc", "reduce_tc_swap", "reduce_ffma", "agent_in_fused", "rec_tc", "rec_tc_bwd", "reduce_fc1")


def _discrete_choices(it, ref, mixer, E=32, HE=64, tol=1e-5):
    """The implementation's branch at every discontinuity of the loss (ReLU derivative of fc1 / hypernet hidden layers /
    V, sign of the abs() inputs), after checking that wherever it differs from the fp64 oracle's branch the oracle's
    pre-activation lies within `tol` (relative to the tensor's rms) of the discontinuity.  Returns (overrides, n_diff)."""
    pre = ref["pre"]
    out, n_diff = {}, 0

    def take(name, impl_positive, impl_sign=None):
        nonlocal n_diff
        p = pre[name]
        scale = max(float(np.sqrt(np.mean(p ** 2))), 1e-3)
        if impl_sign is None:
            differ = impl_positive != (p > 0)
        else:
            differ = impl_sign != np.sign(p)
        if differ.any():
            worst = np.abs(p[differ]).max()
            assert worst <= tol * scale, "%s: %d branch differences, one at distance %.3e from the discontinuity (rms %.3e)" % (
                name, int(differ.sum()), worst, scale)
            assert differ.sum() <= 2 + p.size // 500000, "%s: %d branch differences" % (name, int(differ.sum()))
            n_diff += int(differ.sum())
        return impl_positive if impl_sign is None else impl_sign

    out["x_mask"] = take("x", it["x"] > 0)
    if mixer == "qmix":
        N_E = pre["a1"].shape[1]
        a2, y1 = it["a2"], it["y1"]
        out["a1_sign"] = take("a1", None, np.sign(a2[:, :N_E]))
        out["af_sign"] = take("af", None, np.sign(a2[:, N_E:N_E + E]))
        if pre["h1"] is not None:
            out["h1_mask"] = take("h1", y1[:, :HE] > 0)
            out["hf_mask"] = take("hf", y1[:, HE:2 * HE] > 0)
            out["v1_mask"] = take("v1", y1[:, 2 * HE + E:2 * HE + 2 * E] > 0)
        else:
            out["v1_mask"] = take("v1", y1[:, E:2 * E] > 0)
    return out, n_diff


def _check_against_oracle_full(s, mixer, gap_tol=1e-5):
    """Forward intermediates, loss and every gradient tensor against the fp64 oracle at 1e-5.  The DISCRETE choices inside
    the loss (double-Q arg-max, ReLU / abs derivatives) may differ from the fp64 oracle's only on verified near-ties; the
    oracle is then re-evaluated on the implementation's choices (np_oracle.learner_forward_backward docstring)."""
    from ma_league_b200 import _native as nat
    lib = nat.lib()
    before = {k: lib.mal_stat(k.encode()) for k in _FLAVOURS}
    grads = split_grad(s.learner.forward_backward(s.batch), s.learner)
    th.cuda.synchronize()
    ran = {k: lib.mal_stat(k.encode()) - before[k] for k in _FLAVOURS}
    it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
    ref = _oracle_run(s, mixer, True)
    assert_close(it["hout"], ref["hout"], TOL, "hidden states")
    assert_close(it["chosen"], ref["chosen"], TOL, "chosen")
    flips = np.argwhere(it["argmax"].astype(np.int64) != ref["argmax"])
    if len(flips):
        # an fp32 arg-max may only differ from the fp64 one on a genuine near-tie of the masked online Q-values
        oq = ref["mac_out"][:, 1:].copy()
        oq[np_batch(s.batch)["avail_actions"][:, 1:] == 0] = O.NEG_MASK
        for b, t, n in flips:
            row = oq[b, t, n]
            gap = abs(row[ref["argmax"][b, t, n]] - row[it["argmax"][b, t, n]])
            assert gap <= gap_tol * max(1.0, np.abs(row[row > O.NEG_MASK]).max()), ("argmax flip without a near-tie", b, t, n, gap)
        assert len(flips) <= 1 + it["argmax"].size // 100000, "%d argmax flips" % len(flips)
    discrete, n_diff = _discrete_choices(it, ref, mixer)
    if len(flips) or n_diff:
        L = s.learner
        ref = O.learner_forward_backward(np_params(s.mac.agent), np_params(L.target_mac.agent),
                                         np_params(L.mixer) if mixer == "qmix" else None,
                                         np_params(L.target_mixer) if mixer == "qmix" else None, np_batch(s.batch),
                                         mixer=mixer, double_q=True, gamma=s.args.gamma, dtype=np.float64,
                                         argmax_override=it["argmax"].astype(np.int64), discrete=discrete)
    assert_close(it["target_max"], ref["target_max"], TOL, "target_max")
    assert_close(it["mask"], ref["mask"], 0, "mask")
    assert_close(it["q_tot"], ref["q_tot"], TOL, "q_tot")
    assert_close(it["target_q_tot"], ref["target_q_tot"], TOL, "target_q_tot")
    assert abs(it["scalars"][1] - ref["loss"]) <= TOL * abs(ref["loss"])
    errs = {}
    for k, v in ref["agent_grads"].items():
        errs["agent." + k] = (rel_err(grads["agent." + k], v), )
    for k, v in ref["mixer_grads"].items():
        errs["mixer." + k] = (rel_err(grads["mixer." + k], v), )
    print("discrete differences: argmax %d, relu/abs %d; worst gradient errors: %s" % (
        len(flips), n_diff, sorted(((round(v[0], 9), k) for k, v in errs.items()), reverse=True)[:4]))
    for k, v in ref["agent_grads"].items():
        assert_close(grads["agent." + k], v, TOL, "grad " + k)
    for k, v in ref["mixer_grads"].items():
        assert_close(grads["mixer." + k], v, TOL, "grad mixer " + k)
    assert int(it["scalars"][6:7].view(np.int32)[0]) == ref["stats"]["trained_steps"]
    return ran


@pytest.mark.parametrize("bwd_tc", [0, 1])
@pytest.mark.parametrize("N,B,TT,mixer", [(5, 32, 201, "qmix"), (20, 8, 12, "qmix"), (3, 96, 7, "vdn"), (5, 64, 40, "qmix")])
def test_tensor_core_recurrence_forced(N, B, TT, mixer, bwd_tc):
    """k_gru_fwd_tc (one tcgen05 GEMM per timestep for a tile of up to 128 chains, gi in the tiled layout written by
    k_agent_in_tc) is what the heuristics pick from 8 192 chains on (20v20 / B=1024); force it on shapes the numpy
    oracle affords -- R = 160 (5 groups: tiles of 1), 160, 288 (9 groups), 320 (10 groups) -- and hold hidden states,
    Q-values, loss and every gradient to the fp64 oracle."""
    from ma_league_b200 import _native as nat
    assert (N * B) % 32 == 0
    nat.check(nat.lib().mal_set_option(b"rec_tc", 2), "mal_set_option")
    nat.check(nat.lib().mal_set_option(b"rec_tc_bwd", bwd_tc), "mal_set_option")     # 1: BPTT on the tensor cores too (A operand in tensor memory, tiled gates)
    try:
        ran = _check_against_oracle_full(seeded_system(N, B, TT, mixer, True, seed=50 + N), mixer)
    finally:
        nat.check(nat.lib().mal_set_option(b"rec_tc", 1), "mal_set_option")
        nat.check(nat.lib().mal_set_option(b"rec_tc_bwd", 0), "mal_set_option")
    assert ran["rec_tc"] >= 1 and (ran["rec_tc_bwd"] >= 1) == bool(bwd_tc), ran


def test_config1_qmix_3v3_b32_full_shape():
    """BASELINE.json configs[0]: QMIX 3v3, B=32, T=200."""
    ran = _check_against_oracle_full(seeded_system(3, 32, 201, "qmix", True, seed=41), "qmix")
    assert ran["agent_in_fused"] == 1


def test_config2_vdn_5v5_b32_full_shape():
    """BASELINE.json configs[1]: VDN 5v5, B=32, T=200."""
    ran = _check_against_oracle_full(seeded_system(5, 32, 201, "vdn", True, seed=42), "vdn")
    assert ran["agent_in_fused"] == 1


def test_config3_qmix_10v10_b128_full_shape_default_heuristics():
    """BASELINE.json configs[2]: QMIX 10v10, B=128, T=200, double-Q with avail masks (R = 1 280 rows per net, d_in = 114:
    two k-chunks in the agent-input kernel; the pipelined GEMM and the tensor-core reductions, incl. the transposed
    SWAP orientation for fc1, are what the launch heuristics pick at this size)."""
    ran = _check_against_oracle_full(seeded_system(10, 128, 201, "qmix", True, seed=43), "qmix")
    assert ran["agent_in_fused"] == 1 and ran["linear_tc2"] >= 2, ran
    assert ran["reduce_tc"] >= 3 and ran["reduce_tc_swap"] >= 1 and ran["reduce_ffma"] == 0, ran


def test_config5_dims_qmix_20v20_b64_default_heuristics():
    """BASELINE.json configs[4] dimensions (20v20: N=20, A=26, OBS=168, S=320, d_in=214 -> four k-chunks) at the largest
    batch the numpy oracle affords in seconds (B=64: the same 1 280 rows per net as config 3, enough rows per CTA for
    the launch heuristics to pick exactly the kernels of the B=1024 bench line)."""
    ran = _check_against_oracle_full(seeded_system(20, 64, 201, "qmix", True, seed=44), "qmix")
    assert ran["agent_in_fused"] == 1 and ran["linear_tc2"] >= 2, ran
    assert ran["reduce_tc"] >= 3 and ran["reduce_fc1"] >= 1 and ran["reduce_ffma"] == 0, ran


def test_nan_in_target_max_propagates_like_torch():
    """q_learner.py:65-78 / SURVEY a5: NaN wins torch.max.  A NaN row of the target net's fc2.weight makes action 2 of
    every target Q NaN: with double_q=False the arg-max is 2 wherever action 2 is available (and the target max NaN),
    elsewhere the usual masked max; with double_q=True the online arg-max is untouched and NaN appears exactly where
    it picked action 2."""
    for double_q in (False, True):
        s = seeded_system(4, 6, 11, "qmix", double_q, seed=51)
        with th.no_grad():
            s.learner.target_mac.agent.fc2.weight[2].fill_(float("nan"))
        s.learner.forward_only(s.batch)
        it = {k: v.cpu().numpy() for k, v in s.learner.intermediates(s.batch).items()}
        with np.errstate(invalid="ignore"):
            ref = _oracle_run(s, "qmix", double_q)
        assert np.array_equal(it["argmax"].astype(np.int64), ref["argmax"])
        assert np.array_equal(np.isnan(it["target_max"]), np.isnan(ref["target_max"]))
        assert np.isnan(it["target_max"]).any() and not np.isnan(it["target_max"]).all()
        ok = ~np.isnan(ref["target_max"])
        assert_close(it["target_max"][ok], ref["target_max"][ok], TOL, "finite target_max")
        if not double_q:
            av2 = np_batch(s.batch)["avail_actions"][:, 1:, :, 2] != 0
            assert np.array_equal(it["argmax"] == 2, av2)            # NaN beats every finite Q, masked rows stay finite
        assert np.isnan(it["scalars"][1])                            # the loss is NaN, as in the reference


def test_frozen_agent_trains_only_the_mixer():
    """`freeze_agent_weights()` / args.freeze_native (multi_agent_controller.py:74-76): the reference gives
    frozen agent parameters no gradient, leaves them out of clip_grad_norm_ and RMSprop skips them.  Checked against the
    torch port (same ATen ops as the reference) with requires_grad=False agent parameters over three steps."""
    from oracle import torch_port as TP
    s = seeded_system(3, 6, 10, "qmix", True, seed=61, learner_log_interval=0, clip=0.5)
    L = s.learner
    port = TP.TorchPortLearner(np_params(s.mac.agent), np_params(L.target_mac.agent), np_params(L.mixer),
                               np_params(L.target_mixer), mixer="qmix", double_q=True, gamma=s.args.gamma, lr=s.args.lr,
                               alpha=s.args.optim_alpha, eps=s.args.optim_eps, clip=0.5)
    for p in port.ap.values():
        p.requires_grad_(False)
    s.mac.freeze_agent_weights()
    agent0 = np_params(s.mac.agent)
    cpu_batch = {k: v.cpu() for k, v in s.batch.data.transition_data.items()}
    for i in range(3):
        L.train(s.batch, t_env=i, episode_num=i)
        last = port.train(cpu_batch)
        assert abs(s.logger.stats["home_qlearner_grad_norm"][0] - last["grad_norm"]) <= TOL * last["grad_norm"]
    for k, v in np_params(s.mac.agent).items():
        assert np.array_equal(v, agent0[k]), k                        # untouched
    for k, v in np_params(L.mixer).items():
        assert_close(v, port.mp[k].detach().numpy(), TOL, "mixer " + k)
    n_agent = sum(p.numel() for p in s.mac.parameters())
    assert not L.optimiser.flat_sq[:n_agent].any() and L.optimiser.flat_sq[n_agent:].any()
    assert all(p.grad is None for p in s.mac.parameters()) and all(p.grad is not None for p in L.mixer.parameters())
    # partial freezes are rejected, not silently trained
    next(iter(s.mac.parameters())).requires_grad = True
    with pytest.raises(Exception):
        L.train(s.batch, t_env=9, episode_num=9)


@pytest.mark.parametrize("N,B,TT,mixer", [(5, 32, 201, "qmix"), (3, 4, 9, "vdn"), (10, 6, 12, "qmix")])
def test_dqn_agent_learner_against_oracle(N, B, TT, mixer):
    """the feed-forward DQNAgentNetwork (dqn_agent.py:9-37) behind the "dqn" registry key through the
    same learner step (fc1 GEMM -> Q head -> mixer / TD -> fc2 / fc1 gradients, no recurrence), against the fp64 oracle
    (pinned to the reference by tests/golden/learner_qmix_dqn.npz), incl. the metric's B=32, T=200 shape."""
    s = seeded_system(N, B, TT, mixer, True, seed=70 + N, agent="dqn")
    assert type(s.mac.agent).__name__ == "DQNAgentNetwork" and list(s.mac.agent.state_dict()) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    _check_against_oracle_full(s, mixer)
    s.learner.train(s.batch, 0, 0)                       # the full step (clip + RMSprop) runs on this agent too
    th.cuda.synchronize()


@pytest.mark.parametrize("graphs", [False, True])
@pytest.mark.parametrize("mixer", ["qmix", "vdn"])
def test_balanced_forward_recurrence_is_bit_identical(mixer, graphs):
    """k_gru_fwd9's balanced mode (2 x SMs equal workers over the chain-major step sequence; a chain changes workers once and
    hands its hidden state over through global memory + a flag) runs every chain's steps in the same order with the same
    arithmetic: at the metric's size (320 chains > 296 workers) parameters, optimiser state and statistics must be
    bit-identical to one chain per CTA, eagerly and through the captured graph, over several steps (the hand-over flags are
    left clear by their consumers)."""
    from ma_league_b200 import _native as nat

    def run(balance):
        nat.check(nat.lib().mal_set_option(b"gru_balance", balance), "mal_set_option")
        try:
            s = seeded_system(5, 32, 201, mixer, True, seed=37, learner_log_interval=0)
            s.learner.use_graphs = graphs
            for i in range(5):
                s.learner.train(s.batch, t_env=i, episode_num=i)
            th.cuda.synchronize()
            return np_params(s.mac.agent), np_params(s.learner.mixer) if s.learner.mixer is not None else {}, \
                s.learner.optimiser.flat_sq.cpu().numpy(), {k: v[0] for k, v in s.logger.stats.items()}
        finally:
            nat.check(nat.lib().mal_set_option(b"gru_balance", 1), "mal_set_option")
    a, b = run(2), run(0)
    for x, y in zip(a[:2], b[:2]):
        for k in x:
            assert np.array_equal(x[k], y[k]), k
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3]
