"""CPU port of the reference's hot path that issues the SAME ATen op sequence the reference does.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/np_oracle.py for the rules): used by bench.py's `cpu_baseline` leg
and its `--impl reference` arm, and by tests as a second checker.  `/root/reference` is pure Python and cannot
travel to the GPU box, and its arithmetic lives in PyTorch (which IS on the box), so the faithful CPU baseline is
this functional restatement: per-timestep Python loop, th.cat input assembly, F.linear / gru_cell / bmm / elu,
autograd backward, clip_grad_norm_, torch.optim.RMSprop -- op for op what the reference executes on CPU:

    QLearner.train            marl/learners/q_learner.py:34-125
    BasicMAC.forward          marl/controllers/basic_controller.py:38-54,80-92
    DRQNAgentNetwork.forward  marl/modules/agents/drqn_agent.py:29-35
    QMixer.forward            marl/modules/mixers/qmix.py:41-59
    EpsilonGreedy...select    marl/components/action_selectors.py:44-62
    ReplayBuffer.sample       marl/components/replay_buffers/replay_buffer.py:46-53

Pinned by tests/test_oracle_golden.py::test_torch_port_* against fixtures produced by the unmodified reference
(and timed against the real reference in the build container: see DESIGN.md "CPU baseline").
"""
from collections import OrderedDict

import numpy as np
import torch as th
import torch.nn.functional as F
from torch.distributions import Categorical


def to_params(np_dict, requires_grad=True, device="cpu"):
    return OrderedDict((k, th.tensor(np.asarray(v), dtype=th.float32, device=device, requires_grad=requires_grad))
                       for k, v in np_dict.items())


def agent_step(p, batch, t, h):
    """_build_inputs + DRQN forward for step t; h is [B*N, 64]."""
    obs, onehot = batch["obs"], batch["actions_onehot"]
    B, _, N, _ = obs.shape
    parts = [obs[:, t]]
    parts.append(th.zeros_like(onehot[:, t]) if t == 0 else onehot[:, t - 1])
    parts.append(th.eye(N, device=obs.device).unsqueeze(0).expand(B, -1, -1))
    inp = th.cat([x.reshape(B * N, -1) for x in parts], dim=1)
    x = F.relu(F.linear(inp, p["fc1.weight"], p["fc1.bias"]))
    if "gru.weight_ih" not in p:                    # DQNAgentNetwork (dqn_agent.py:34-37): no recurrence
        return F.linear(x, p["fc2.weight"], p["fc2.bias"]).view(B, N, -1), h
    h = th.gru_cell(x, h.reshape(-1, x.shape[1]), p["gru.weight_ih"], p["gru.weight_hh"], p["gru.bias_ih"],
                    p["gru.bias_hh"])
    q = F.linear(h, p["fc2.weight"], p["fc2.bias"])
    return q.view(B, N, -1), h


def unroll(p, batch):
    B, TT, N, _ = batch["obs"].shape
    h = th.zeros(1, p["fc1.weight"].shape[0], device=batch["obs"].device).unsqueeze(0).expand(B, N, -1)
    outs = []
    for t in range(TT):
        q, h = agent_step(p, batch, t, h)
        outs.append(q)
    return th.stack(outs, dim=1)


def _hyper(mp, name, s):
    if name + ".0.weight" in mp:
        return F.linear(F.relu(F.linear(s, mp[name + ".0.weight"], mp[name + ".0.bias"])), mp[name + ".2.weight"],
                        mp[name + ".2.bias"])
    return F.linear(s, mp[name + ".weight"], mp[name + ".bias"])


def qmix(mp, agent_qs, states):
    bs, _, N = agent_qs.shape
    E = mp["hyper_b_1.weight"].shape[0]
    s = states.reshape(-1, states.shape[-1])
    qs = agent_qs.view(-1, 1, N)
    w1 = th.abs(_hyper(mp, "hyper_w_1", s)).view(-1, N, E)
    b1 = F.linear(s, mp["hyper_b_1.weight"], mp["hyper_b_1.bias"]).view(-1, 1, E)
    hidden = F.elu(th.bmm(qs, w1) + b1)
    wf = th.abs(_hyper(mp, "hyper_w_final", s)).view(-1, E, 1)
    v = F.linear(F.relu(F.linear(s, mp["V.0.weight"], mp["V.0.bias"])), mp["V.2.weight"], mp["V.2.bias"]).view(-1, 1, 1)
    return (th.bmm(hidden, wf) + v).view(bs, -1, 1)


class TorchPortLearner:
    """Stateful CPU learner: parameters, target copies and a stock torch RMSprop."""

    def __init__(self, agent_p, target_agent_p, mixer_p, target_mixer_p, *, mixer, double_q, gamma, lr, alpha, eps,
                 clip, device="cpu"):
        # device="cuda": the same ATen op sequence in PyTorch eager on the GPU ("just run pymarl on the GPU",
        # BASELINE.md section 3 optional bar) -- a reported baseline, never part of the product path
        self.ap = to_params(agent_p, device=device)
        self.tp = to_params(target_agent_p, device=device)
        self.mp = to_params(mixer_p, device=device) if mixer == "qmix" else OrderedDict()
        self.tmp = to_params(target_mixer_p, device=device) if mixer == "qmix" else OrderedDict()
        self.mixer, self.double_q, self.gamma, self.clip = mixer, double_q, gamma, clip
        self.params = list(self.ap.values()) + list(self.mp.values())
        self.opt = th.optim.RMSprop(self.params, lr=lr, alpha=alpha, eps=eps)
        self.last = {}

    def train(self, batch):
        rewards = batch["reward"][:, :-1]
        actions = batch["actions"][:, :-1]
        terminated = batch["terminated"][:, :-1].float()
        mask = batch["filled"][:, :-1].float()
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        avail = batch["avail_actions"]
        mac_out = unroll(self.ap, batch)
        chosen = th.gather(mac_out[:, :-1], dim=3, index=actions).squeeze(3)
        target_out = unroll(self.tp, batch)[:, 1:]          # built with autograd on, like the reference (:59-62)
        target_out[avail[:, 1:] == 0] = -9999999
        if self.double_q:
            det = mac_out.clone().detach()
            det[avail == 0] = -9999999
            cur_max = det[:, 1:].max(dim=3, keepdim=True)[1]
            tmax = th.gather(target_out, 3, cur_max).squeeze(3)
        else:
            tmax = target_out.max(dim=3)[0]
        if self.mixer == "qmix":
            q_tot = qmix(self.mp, chosen, batch["state"][:, :-1])
            tq_tot = qmix(self.tmp, tmax, batch["state"][:, 1:])
        else:
            q_tot, tq_tot = th.sum(chosen, dim=2, keepdim=True), th.sum(tmax, dim=2, keepdim=True)
        targets = rewards + self.gamma * (1 - terminated) * tq_tot
        td = q_tot - targets.detach()
        mask = mask.expand_as(td)
        masked = td * mask
        loss = (masked ** 2).sum() / mask.sum()
        self.opt.zero_grad()
        loss.backward()
        grad_norm = th.nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        trained = th.count_nonzero(mask).item()            # per-step host sync of the reference (:112-113)
        self.last = dict(loss=float(loss.item()), grad_norm=float(grad_norm), trained_steps=trained,
                         q_tot=q_tot.detach(), mac_out=mac_out.detach())
        return self.last


def select_actions(p, batch, t, h, epsilon):
    """BasicMAC.select_actions on CPU with torch's CPU generator; returns (actions, greedy, new hidden)."""
    with th.no_grad():
        q, h = agent_step(p, batch, t, h)
        avail = batch["avail_actions"][:, t]
        masked = q.clone()
        masked[avail == 0.0] = -float("inf")
        rnd = th.rand_like(q[:, :, 0])
        pick_random = (rnd < epsilon).long()
        random_actions = Categorical(avail.float()).sample().long()
        picked = pick_random * random_actions + (1 - pick_random) * masked.max(dim=2)[1]
    return picked, 1 - pick_random, h


def replay_sample(buffer, episodes_in_buffer, batch_size):
    """ReplayBuffer.sample on host tensors: per-key advanced-index gather (replay_buffer.py:46-53)."""
    ids = np.random.choice(episodes_in_buffer, batch_size, replace=False)
    return {k: v[ids] for k, v in buffer.items()}
