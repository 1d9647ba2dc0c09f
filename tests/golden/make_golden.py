"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own hot-path classes from ``/root/reference/src`` (recipe: SURVEY.md appendix B),
drives them on small seeded synthetic episodes on CPU and stores inputs, parameters and every intermediate the
parity tests compare (``mac_out``, chosen Q, masked target max, q_tot, loss, gradients before clipping,
post-step parameters and RMSprop state, selected actions for captured random draws, ring-buffer indices).
The fixtures pin ``oracle/np_oracle.py`` (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import warnings
from types import SimpleNamespace as SN

import numpy as np
import torch as th

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def ref_imports():
    sys.path.insert(0, REF)
    from marl.components.episode_batch import EpisodeBatch
    from marl.components.replay_buffers import ReplayBuffer
    from marl.components.transforms import OneHot
    from marl.components.action_selectors import EpsilonGreedyActionSelector
    from marl.controllers import REGISTRY as mac_REGISTRY
    from marl.learners import REGISTRY as le_REGISTRY
    return SN(EpisodeBatch=EpisodeBatch, ReplayBuffer=ReplayBuffer, OneHot=OneHot,
              Selector=EpsilonGreedyActionSelector, mac=mac_REGISTRY, learner=le_REGISTRY)


class NullLogger:
    def log_stat(self, *a):
        pass

    def info(self, *a):
        pass


def make_args(N, A, S, mixer, double_q, hypernet_layers=2, grad_norm_clip=10.0):
    return SN(n_agents=N, n_actions=A, state_shape=S, agent_output_type="q", action_selector="epsilon_greedy",
              freeze_native=False, agent="rnn", obs_last_action=True, obs_agent_id=True, rnn_hidden_dim=64,
              device=th.device("cpu"), epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000,
              mixer=mixer, double_q=double_q, gamma=0.99, grad_norm_clip=grad_norm_clip,
              target_update_interval=200, learner_log_interval=10 ** 9, lr=5e-4, optim_alpha=0.99, optim_eps=1e-5,
              mixing_embed_dim=32, hypernet_layers=hypernet_layers, hypernet_embed=64, mac="basic")


def make_scheme(R, N, A, OBS, S):
    scheme = {"state": {"vshape": S}, "obs": {"vshape": OBS, "group": "agents"},
              "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
              "avail_actions": {"vshape": (A,), "group": "agents", "dtype": th.int},
              "reward": {"vshape": (1,)}, "terminated": {"vshape": (1,), "dtype": th.uint8}}
    groups = {"agents": N}
    pre = {"actions": ("actions_onehot", [R.OneHot(out_dim=A)])}
    return scheme, groups, pre


def synth_episodes(R, scheme, groups, pre, B, TT, N, A, OBS, S, gen, var_len=True):
    """Synthetic episodes per SURVEY.md 8(d): N(0,1) obs/state/reward, Bernoulli(0.7) avail with action 0 forced."""
    eb = R.EpisodeBatch(scheme, groups, B, TT, preprocess=pre)
    avail = (th.rand(B, TT, N, A, generator=gen) < 0.7).int()
    avail[..., 0] = 1
    e = th.empty(B, TT, N, A).exponential_(generator=gen)
    actions = (avail.float() / e).argmax(-1, keepdim=True)
    T = TT - 1
    lens = th.randint(max(T // 2, 1), T + 1, (B,), generator=gen) if var_len else th.full((B,), T)
    lens[0] = T  # at least one full-length episode so max_t_filled == TT
    term = th.zeros(B, TT, 1, dtype=th.uint8)
    for b in range(B):
        term[b, lens[b] - 1] = 1
    eb.update({"state": th.randn(B, TT, S, generator=gen), "obs": th.randn(B, TT, N, OBS, generator=gen),
               "actions": actions, "avail_actions": avail, "reward": th.randn(B, TT, 1, generator=gen),
               "terminated": term})
    for b in range(B):  # zero the padding the way a real rollout leaves it (never written)
        L = int(lens[b]) + 1
        for k, v in eb.data.transition_data.items():
            v[b, L:] = 0
    return eb


def batch_to_np(eb):
    return {k: v.detach().numpy().copy() for k, v in eb.data.transition_data.items()}


def sd_to_np(sd, prefix):
    return {prefix + k: v.detach().numpy().copy() for k, v in sd.items()}


def learner_case(R, name, *, B, TT, N, A, OBS, S, mixer, double_q, hypernet_layers=2, seed=0, clip=10.0, steps=2,
                 agent="rnn"):
    th.manual_seed(seed)
    np.random.seed(seed)
    gen = th.Generator().manual_seed(seed + 1)
    args = make_args(N, A, S, mixer, double_q, hypernet_layers, clip)
    args.agent, args.batch_size = agent, B
    scheme, groups, pre = make_scheme(R, N, A, OBS, S)
    buf = R.ReplayBuffer(scheme, groups, B, TT, preprocess=pre, device="cpu")
    mac = R.mac["basic"](buf.scheme, groups, args)
    learner = R.learner["q"](mac, buf.scheme, NullLogger(), args, name="home")
    learner.build_optimizer()
    if agent == "dqn":
        # The reference's BasicMAC.init_hidden crashes for this agent (basic_controller.py:59-60 expands the 3-D
        # placeholder of dqn_agent.py:27-32 with three sizes -> RuntimeError), so QLearner.train cannot run with it as
        # shipped.  The placeholder is never read (dqn_agent.py:34-37 passes it through), so the harness -- not the
        # reference's files -- replaces that one call; everything else is the reference's own code.
        for m in (mac, learner.target_mac):
            m.init_hidden = (lambda bs, m=m: setattr(m, "hidden_states", th.zeros(bs, N, 1)))
    # make the target network differ from the online one (as after some training)
    with th.no_grad():
        for p in learner.target_mac.parameters():
            p.add_(0.05 * th.randn(p.shape, generator=gen))
        if mixer == "qmix":
            for p in learner.target_mixer.parameters():
                p.add_(0.05 * th.randn(p.shape, generator=gen))
    eb = synth_episodes(R, scheme, groups, pre, B, TT, N, A, OBS, S, gen)
    out = {"meta": np.array([B, TT, N, A, OBS, S, int(mixer == "qmix"), int(double_q), hypernet_layers, steps],
                            dtype=np.int64),
           "hyper": np.array([args.gamma, args.lr, args.optim_alpha, args.optim_eps, clip], dtype=np.float64)}
    if agent != "rnn":
        out["agent_kind"] = np.int64(1)
    out.update({"batch." + k: v for k, v in batch_to_np(eb).items()})
    out.update(sd_to_np(mac.agent.state_dict(), "agent0."))
    out.update(sd_to_np(learner.target_mac.agent.state_dict(), "tagent0."))
    if mixer == "qmix":
        out.update(sd_to_np(learner.mixer.state_dict(), "mixer0."))
        out.update(sd_to_np(learner.target_mixer.state_dict(), "tmixer0."))

    # ---- intermediates of the first step, recomputed with the reference's own modules (q_learner.py:46-98)
    with th.no_grad():
        mac.init_hidden(B)
        mo = th.stack([mac.forward(eb, t=t) for t in range(TT)], dim=1)
        learner.target_mac.init_hidden(B)
        tmo = th.stack([learner.target_mac.forward(eb, t=t) for t in range(TT)], dim=1)
        out["mac_out"] = mo.numpy().copy()
        out["target_mac_out"] = tmo.numpy().copy()
        chosen = th.gather(mo[:, :-1], 3, eb["actions"][:, :-1]).squeeze(3)
        t1 = tmo[:, 1:].clone()
        t1[eb["avail_actions"][:, 1:] == 0] = -9999999
        if double_q:
            md = mo.clone()
            md[eb["avail_actions"] == 0] = -9999999
            amax = md[:, 1:].max(dim=3, keepdim=True)[1]
            tmax = th.gather(t1, 3, amax).squeeze(3)
        else:
            tmax, amax = t1.max(dim=3)
        out["chosen"] = chosen.numpy().copy()
        out["target_max"] = tmax.numpy().copy()
        out["argmax"] = amax.reshape(tmax.shape).numpy().copy()
        if mixer == "qmix":
            out["q_tot"] = learner.mixer(chosen, eb["state"][:, :-1]).numpy().copy()
            out["target_q_tot"] = learner.target_mixer(tmax, eb["state"][:, 1:]).numpy().copy()
        else:
            out["q_tot"] = learner.mixer(chosen, eb["state"][:, :-1]).numpy().copy()
            out["target_q_tot"] = learner.target_mixer(tmax, eb["state"][:, 1:]).numpy().copy()

    # ---- the real train() calls; capture unclipped grads, grad norm and loss of step 1 through hooks
    captured = {}
    orig_clip = th.nn.utils.clip_grad_norm_

    def spy_clip(params, max_norm, *a, **k):
        params = list(params)
        if "grads" not in captured:
            captured["grads"] = [p.grad.detach().clone() for p in params]
        norm = orig_clip(params, max_norm, *a, **k)
        captured.setdefault("norm", norm.detach().clone())
        return norm

    th.nn.utils.clip_grad_norm_ = spy_clip
    logged = {}
    learner.logger = SN(log_stat=lambda k, v, t: logged.setdefault(k, float(np.asarray(v))), info=lambda s: None)
    learner.args.learner_log_interval = 0
    learner.log_stats_t = -1
    try:
        for i in range(steps):
            learner.train(eb, t_env=i, episode_num=i)
            if i == 0:
                learner.args.learner_log_interval = 10 ** 9
    finally:
        th.nn.utils.clip_grad_norm_ = orig_clip
    names = ["agent." + k for k, _ in mac.agent.named_parameters()]
    if mixer == "qmix":
        names += ["mixer." + k for k, _ in learner.mixer.named_parameters()]
    for n_, g_ in zip(names, captured["grads"]):
        out["grad." + n_] = g_.numpy().copy()
    out["grad_norm"] = captured["norm"].numpy().copy()
    for k, v in logged.items():
        out["stat." + k.replace("home_qlearner_", "")] = np.float64(v)
    out["trained_steps"] = np.int64(mac.agent.trained_steps)
    out.update(sd_to_np(mac.agent.state_dict(), "agentK."))
    if mixer == "qmix":
        out.update(sd_to_np(learner.mixer.state_dict(), "mixerK."))
    st = learner.optimiser.state_dict()["state"]
    for i, n_ in enumerate(names):
        out["sqavg." + n_] = st[i]["square_avg"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", logged.get("home_qlearner_loss"), "norm", float(captured["norm"]))


def select_case(R, name, *, bs, N, A, seed=0):
    """EpsilonGreedyActionSelector.select with the torch CPU draws captured (action_selectors.py:44-62)."""
    args = make_args(N, A, 4, "vdn", True)
    sel = R.Selector(args)
    gen = th.Generator().manual_seed(seed)
    out = {}
    cases = []
    for ci, (t_env, test_mode) in enumerate([(0, False), (25000, False), (49000, False), (10 ** 6, False),
                                             (0, True)]):
        q = th.randn(bs, N, A, generator=gen)
        if ci == 2:
            q[:, :, 1] = q[:, :, 3]  # exact ties in the greedy argmax
        avail = (th.rand(bs, N, A, generator=gen) < 0.6).int()
        avail[..., A - 1] = 1
        th.manual_seed(1000 + ci)
        picked, greedy = sel.select(q, avail, t_env, test_mode)
        th.manual_seed(1000 + ci)  # replay the same stream to capture the draws (rand_like, then exponential_)
        u = th.rand_like(q[:, :, 0])
        e = th.empty(bs * N, A).exponential_()
        out["c%d.q" % ci] = q.numpy()
        out["c%d.avail" % ci] = avail.numpy()
        out["c%d.u" % ci] = u.numpy()
        out["c%d.e" % ci] = e.numpy()
        out["c%d.picked" % ci] = picked.numpy()
        out["c%d.greedy" % ci] = greedy.numpy()
        out["c%d.eps" % ci] = np.float64(sel.epsilon)
        out["c%d.t_env" % ci] = np.int64(t_env)
        out["c%d.test_mode" % ci] = np.int64(test_mode)
        cases.append(ci)
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def mac_step_case(R, name, *, bs, N, A, OBS, S, seed=0):
    """BasicMAC.select_actions over 3 consecutive steps incl. t=0 (basic_controller.py:29-50)."""
    th.manual_seed(seed)
    gen = th.Generator().manual_seed(seed + 7)
    args = make_args(N, A, S, "vdn", True)
    scheme, groups, pre = make_scheme(R, N, A, OBS, S)
    TT = 4
    eb = synth_episodes(R, scheme, groups, pre, bs, TT, N, A, OBS, S, gen, var_len=False)
    full_scheme = R.ReplayBuffer(scheme, groups, 1, TT, preprocess=pre).scheme
    mac = R.mac["basic"](full_scheme, groups, args)
    out = {"meta": np.array([bs, TT, N, A, OBS, S], dtype=np.int64)}
    out.update({"batch." + k: v for k, v in batch_to_np(eb).items()})
    out.update(sd_to_np(mac.agent.state_dict(), "agent."))
    mac.init_hidden(bs)
    for t in range(3):
        with th.no_grad():
            th.manual_seed(500 + t)
            h_before = mac.hidden_states.reshape(bs * N, -1).clone()
            acts, greedy = mac.select_actions(eb, t_ep=t, t_env=20000 * t, test_mode=False)
            th.manual_seed(500 + t)
            u = th.rand(bs, N)
            e = th.empty(bs * N, A).exponential_()
            # q of this step recomputed from the stored hidden state
            q, _ = mac.agent(mac._build_inputs(eb, t), h_before)
        out["t%d.actions" % t] = acts.numpy().copy()
        out["t%d.greedy" % t] = greedy.numpy().copy()
        out["t%d.u" % t] = u.numpy()
        out["t%d.e" % t] = e.numpy()
        out["t%d.q" % t] = q.view(bs, N, A).numpy().copy()
        out["t%d.hidden" % t] = mac.hidden_states.detach().numpy().copy()
        out["t%d.eps" % t] = np.float64(mac.action_selector.epsilon)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def replay_case(R, name, seed=3):
    """ReplayBuffer insert with wrap-around + uniform sampling indices (replay_buffer.py:22-53)."""
    N, A, OBS, S, TT = 2, 4, 3, 5, 5
    scheme, groups, pre = make_scheme(R, N, A, OBS, S)
    gen = th.Generator().manual_seed(seed)
    buf = R.ReplayBuffer(scheme, groups, 7, TT, preprocess=pre, device="cpu")
    out = {"meta": np.array([7, TT, N, A, OBS, S], dtype=np.int64)}
    log = []
    np.random.seed(seed)
    for i, n in enumerate([3, 3, 3, 5, 1]):
        eb = synth_episodes(R, scheme, groups, pre, n, TT, N, A, OBS, S, gen)
        for k, v in batch_to_np(eb).items():
            out["ins%d.%s" % (i, k)] = v
        buf.insert_episode_batch(eb)
        log.append([buf.buffer_index, buf.episodes_in_buffer])
        for k, v in buf.data.transition_data.items():
            out["buf%d.%s" % (i, k)] = v.numpy().copy()
        if buf.can_sample(4):
            state = np.random.get_state()
            ids = np.random.choice(buf.episodes_in_buffer, 4, replace=False)
            np.random.set_state(state)
            smp = buf.sample(4)
            out["smp%d.ids" % i] = ids
            for k, v in smp.data.transition_data.items():
                out["smp%d.%s" % (i, k)] = v.numpy().copy()
            out["smp%d.max_t" % i] = np.int64(int(smp.max_t_filled()))
    out["counters"] = np.array(log, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", log)


def dqn_agent_case(R, name, *, rows, d_in, A, seed=0):
    """DQNAgentNetwork.forward(inputs, hidden_states) (marl/modules/agents/dqn_agent.py:34-37) and its state_dict."""
    from marl.modules.agents import REGISTRY as agent_REGISTRY
    th.manual_seed(seed)
    args = SN(rnn_hidden_dim=64, n_actions=A, device=th.device("cpu"), batch_size=rows)
    net = agent_REGISTRY["dqn"](d_in, args)
    x = th.randn(rows, d_in)
    hidden = net.init_hidden()
    with th.no_grad():
        q, h_out = net(x, hidden)
    assert h_out is hidden
    out = {"meta": np.array([rows, d_in, A], dtype=np.int64), "x": x.numpy(), "q": q.numpy(),
           "hidden_shape": np.array(hidden.shape, dtype=np.int64)}
    out.update(sd_to_np(net.state_dict(), "agent."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def main():
    R = ref_imports()
    if "--only-new" in sys.argv:      # fixtures added in round 2 (the earlier ones are reproduced bit-identically by a full run)
        dqn_agent_case(R, "dqn_agent", rows=12, d_in=25, A=9)
        learner_case(R, "learner_qmix_dqn", B=3, TT=7, N=3, A=9, OBS=12, S=14, mixer="qmix", double_q=True, seed=4,
                     agent="dqn")
        return
    learner_case(R, "learner_qmix_3v3", B=4, TT=9, N=3, A=9, OBS=12, S=14, mixer="qmix", double_q=True)
    learner_case(R, "learner_vdn_2v2", B=3, TT=6, N=2, A=8, OBS=10, S=9, mixer="vdn", double_q=True, seed=1)
    learner_case(R, "learner_qmix_nodouble", B=2, TT=5, N=4, A=10, OBS=7, S=11, mixer="qmix", double_q=False,
                 seed=2, clip=0.5)
    select_case(R, "select_eps_greedy", bs=6, N=5, A=11)
    mac_step_case(R, "mac_select_actions", bs=3, N=4, A=10, OBS=9, S=6)
    replay_case(R, "replay_ring")
    dqn_agent_case(R, "dqn_agent", rows=12, d_in=25, A=9)
    learner_case(R, "learner_qmix_dqn", B=3, TT=7, N=3, A=9, OBS=12, S=14, mixer="qmix", double_q=True, seed=4, agent="dqn")


if __name__ == "__main__":
    main()
