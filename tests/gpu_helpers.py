"""Build learners / batches on the GPU from golden fixtures or seeds (test infrastructure)."""
import numpy as np
import torch as th

import ma_league_b200 as M
from ma_league_b200.synthetic import make_args, make_scheme, synth_episode_data, fill_episode_batch
from tests.helpers import sub


class ListLogger:
    def __init__(self):
        self.stats = {}
        self.infos = []

    def log_stat(self, k, v, t):
        self.stats[k] = (float(np.asarray(v)), t)

    def info(self, s):
        self.infos.append(s)


def to_sd(d, device):
    return {k: th.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in d.items()}


def build_system(N, A, OBS, S, B, TT, mixer, double_q, device="cuda", hypernet_layers=2, clip=10.0, buffer_size=None,
                 **over):
    args = make_args(N, A, S, mixer=mixer, double_q=double_q, device=device, hypernet_layers=hypernet_layers,
                     grad_norm_clip=clip, **over)
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, buffer_size or B, TT, preprocess=pre, device=device)
    mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    logger = ListLogger()
    learner = M.learner_REGISTRY["q"](mac, buf.scheme, logger, args, name="home")
    learner.build_optimizer()
    return SN(args=args, scheme=scheme, groups=groups, pre=pre, buf=buf, mac=mac, learner=learner, logger=logger)


class SN(dict):
    __getattr__ = dict.__getitem__


def system_from_golden(g, device="cuda"):
    B, TT, N, A, OBS, S, is_qmix, double_q, layers, steps = [int(x) for x in g["meta"]]
    gamma, lr, alpha, eps, clip = [float(x) for x in g["hyper"]]
    over = dict(agent="dqn") if int(g.get("agent_kind", 0)) == 1 else {}
    sysm = build_system(N, A, OBS, S, B, TT, "qmix" if is_qmix else "vdn", bool(double_q), device, layers, clip, **over)
    sysm.mac.agent.load_state_dict(to_sd(sub(g, "agent0."), device))
    sysm.learner.target_mac.agent.load_state_dict(to_sd(sub(g, "tagent0."), device))
    if is_qmix:
        sysm.learner.mixer.load_state_dict(to_sd(sub(g, "mixer0."), device))
        sysm.learner.target_mixer.load_state_dict(to_sd(sub(g, "tmixer0."), device))
    eb = M.EpisodeBatch(sysm.scheme, sysm.groups, B, TT, preprocess=sysm.pre, device=device)
    for k, v in sub(g, "batch.").items():
        eb.data.transition_data[k].copy_(th.from_numpy(v).to(device))
    sysm["batch"] = eb
    sysm["steps"] = steps
    return sysm


def seeded_system(N, B, TT, mixer, double_q=True, seed=0, device="cuda", var_len=True, perturb_target=True, **kw):
    A, OBS, S = 6 + N, 8 + 8 * N, 16 * N
    th.manual_seed(seed)
    sysm = build_system(N, A, OBS, S, B, TT, mixer, double_q, device, **kw)
    gen = th.Generator().manual_seed(seed + 1)
    if perturb_target:
        with th.no_grad():
            for p in list(sysm.learner.target_mac.parameters()) + list(sysm.learner.target_mixer.parameters()):
                p.add_(0.05 * th.randn(p.shape, generator=gen).to(device))
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=var_len, device=device)
    lens[0] = TT - 1
    eb = M.EpisodeBatch(sysm.scheme, sysm.groups, B, TT, preprocess=sysm.pre, device=device)
    fill_episode_batch(eb, data, lens)
    sysm["batch"] = eb
    return sysm


def np_params(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def np_batch(eb):
    return {k: v.detach().cpu().numpy().copy() for k, v in eb.data.transition_data.items()}


def split_grad(flat, learner):
    out, off = {}, 0
    names = ["agent." + k for k, _ in learner.mac.agent.named_parameters()] + \
            ["mixer." + k for k, _ in learner.mixer.named_parameters()]
    for n, p in zip(names, learner.parameters()):
        out[n] = flat[off:off + p.numel()].view(p.shape).cpu().numpy().copy()
        off += p.numel()
    return out
