"""Synthetic workloads of SURVEY.md section 8(d): the env (`maenv`) is an un-vendored dependency, so dims are defined
here: A = 6 + N, OBS = 8 + 8N, S = 16N, H = 64, E = 32, HE = 64, T = 200 transitions (201 stored steps)."""
from types import SimpleNamespace as SN

import torch as th

from .components.transforms import OneHot

CONFIGS = {   # BASELINE.json `configs`
    "qmix_3v3_b32": dict(N=3, B=32, mixer="qmix"),
    "vdn_5v5_b32": dict(N=5, B=32, mixer="vdn"),
    "qmix_5v5_b32": dict(N=5, B=32, mixer="qmix"),       # the league matchup shape; metric is quoted on this
    "qmix_10v10_b128": dict(N=10, B=128, mixer="qmix"),
    "qmix_20v20_b1024": dict(N=20, B=1024, mixer="qmix"),
}


def dims(N):
    return dict(N=N, A=6 + N, OBS=8 + 8 * N, S=16 * N)


def make_args(N, A, S, mixer="qmix", double_q=True, device="cuda", hypernet_layers=2, **over):
    a = SN(n_agents=N, n_actions=A, state_shape=S, agent_output_type="q", action_selector="epsilon_greedy",
           freeze_native=False, agent="rnn", obs_last_action=True, obs_agent_id=True, rnn_hidden_dim=64,
           device=th.device(device), epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000,
           mixer=mixer, double_q=double_q, gamma=0.99, grad_norm_clip=10, target_update_interval=200,
           learner_log_interval=10 ** 9, lr=5e-4, optim_alpha=0.99, optim_eps=1e-5, mixing_embed_dim=32,
           hypernet_layers=hypernet_layers, hypernet_embed=64, mac="basic", learner="q", batch_size=32,
           buffer_size=5000)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def make_scheme(N, A, OBS, S):
    """runs/train/ma_experiment.py:99-118"""
    scheme = {"state": {"vshape": S}, "obs": {"vshape": OBS, "group": "agents"},
              "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
              "avail_actions": {"vshape": (A,), "group": "agents", "dtype": th.int},
              "reward": {"vshape": (1,)}, "terminated": {"vshape": (1,), "dtype": th.uint8}}
    groups = {"agents": N}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=A)])}
    return scheme, groups, preprocess


def synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=True, device="cpu"):
    """N(0,1) obs/state/reward, Bernoulli(0.7) avail with action 0 forced, uniform available actions,
    episode length L_b ~ U{T/2..T} with terminated[b, L_b-1] = 1.  Returns (dict of tensors, lengths)."""
    T = TT - 1
    avail = (th.rand(B, TT, N, A, generator=gen) < 0.7).int()
    avail[..., 0] = 1
    e = th.empty(B, TT, N, A).exponential_(generator=gen)
    actions = (avail.float() / e).argmax(-1, keepdim=True)
    lens = th.randint(max(T // 2, 1), T + 1, (B,), generator=gen) if var_len else th.full((B,), T)
    term = th.zeros(B, TT, 1, dtype=th.uint8)
    term[th.arange(B), lens - 1] = 1
    data = {"state": th.randn(B, TT, S, generator=gen), "obs": th.randn(B, TT, N, OBS, generator=gen),
            "actions": actions, "avail_actions": avail, "reward": th.randn(B, TT, 1, generator=gen),
            "terminated": term}
    return {k: v.to(device) for k, v in data.items()}, lens


def fill_episode_batch(eb, data, lens):
    """Write full tensors with one update() and clear everything past each episode's end (as a rollout leaves it)."""
    eb.update(data)
    TT = eb.max_seq_length
    steps = th.arange(TT, device=eb["filled"].device).view(1, TT)
    keep = steps <= lens.to(steps.device).view(-1, 1)      # steps 0..L_b are filled
    for k, v in eb.data.transition_data.items():
        m = keep.view(keep.shape + (1,) * (v.dim() - 2))
        v.mul_(m.to(v.dtype))
    return eb
