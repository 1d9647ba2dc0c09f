"""CPU oracle for the value-decomposition hot path of PMatthaei/ma-league.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ma_league_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker /
baseline, never as the product path.

This is a plain-numpy restatement (explicit forward AND hand-derived backward,
no autograd) of the algorithm the reference executes through PyTorch ATen ops.
The arithmetic itself lives in a third-party dependency that is not vendored in
``/root/reference``: PyTorch (pinned ``torch==1.9.0+cu111`` in the reference's
``run.sh:25``); the formulas below restate the documented semantics of
``nn.Linear``, ``nn.GRUCell`` (gate order r,z,n), ``F.elu``, ``th.abs``,
``clip_grad_norm_`` and ``optim.RMSprop`` at the reference's call sites, which
are cited per function as ``file:line`` relative to ``/root/reference/src``.

Parity pinning: the reference ships NO test or golden vector for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the
reference itself, imported and run in the build container by
``tests/golden/make_golden.py``; the resulting fixtures are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks every function
here against them.

All functions are dtype-generic: run them on float32 inputs to mimic the
reference's precision or on float64 casts to obtain a high-precision truth.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

H_DEFAULT = 64
NEG_MASK = -9999999.0  # marl/learners/q_learner.py:68,73


# --------------------------------------------------------------------------
# parameter layouts (state_dict order of the reference modules)
# --------------------------------------------------------------------------
def agent_param_shapes(d_in: int, n_actions: int, hidden: int = H_DEFAULT) -> "OrderedDict[str, tuple]":
    """Keys/shapes of DRQNAgentNetwork.state_dict(), marl/modules/agents/drqn_agent.py:21-23."""
    return OrderedDict([
        ("fc1.weight", (hidden, d_in)), ("fc1.bias", (hidden,)),
        ("gru.weight_ih", (3 * hidden, hidden)), ("gru.weight_hh", (3 * hidden, hidden)),
        ("gru.bias_ih", (3 * hidden,)), ("gru.bias_hh", (3 * hidden,)),
        ("fc2.weight", (n_actions, hidden)), ("fc2.bias", (n_actions,)),
    ])


def qmix_param_shapes(state_dim: int, n_agents: int, embed: int = 32, hypernet_embed: int = 64,
                      hypernet_layers: int = 2) -> "OrderedDict[str, tuple]":
    """Keys/shapes of QMixer.state_dict(), marl/modules/mixers/qmix.py:16-39."""
    if hypernet_layers == 2:
        d = [
            ("hyper_w_1.0.weight", (hypernet_embed, state_dim)), ("hyper_w_1.0.bias", (hypernet_embed,)),
            ("hyper_w_1.2.weight", (embed * n_agents, hypernet_embed)), ("hyper_w_1.2.bias", (embed * n_agents,)),
            ("hyper_w_final.0.weight", (hypernet_embed, state_dim)), ("hyper_w_final.0.bias", (hypernet_embed,)),
            ("hyper_w_final.2.weight", (embed, hypernet_embed)), ("hyper_w_final.2.bias", (embed,)),
        ]
    elif hypernet_layers == 1:
        d = [
            ("hyper_w_1.weight", (embed * n_agents, state_dim)), ("hyper_w_1.bias", (embed * n_agents,)),
            ("hyper_w_final.weight", (embed, state_dim)), ("hyper_w_final.bias", (embed,)),
        ]
    else:
        raise ValueError("hypernet_layers must be 1 or 2 (qmix.py:29-32)")
    d += [
        ("hyper_b_1.weight", (embed, state_dim)), ("hyper_b_1.bias", (embed,)),
        ("V.0.weight", (embed, state_dim)), ("V.0.bias", (embed,)),
        ("V.2.weight", (1, embed)), ("V.2.bias", (1,)),
    ]
    return OrderedDict(d)


def init_params(shapes, rng: np.random.Generator, dtype=np.float32):
    """U(+-1/sqrt(fan_in)) like torch's default Linear/GRUCell init (distribution only, not the stream)."""
    out = OrderedDict()
    fan = {}
    for k, s in shapes.items():
        if len(s) == 2:
            fan[k.rsplit(".", 1)[0] if not k.startswith("gru") else "gru"] = s[1]
    for k, s in shapes.items():
        base = "gru" if k.startswith("gru") else k.rsplit(".", 1)[0]
        bound = 1.0 / np.sqrt(fan[base])
        out[k] = rng.uniform(-bound, bound, size=s).astype(dtype)
    return out


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------
# a1: BasicMAC._build_inputs   marl/controllers/basic_controller.py:80-92
# --------------------------------------------------------------------------
def build_inputs(obs, actions_onehot, t: int):
    """obs [B,TT,N,OBS], actions_onehot [B,TT,N,A] -> [B*N, OBS+A+N]; columns obs | last-action | agent id."""
    B, _, N, _ = obs.shape
    last = np.zeros_like(actions_onehot[:, 0]) if t == 0 else actions_onehot[:, t - 1]
    eye = np.broadcast_to(np.eye(N, dtype=obs.dtype)[None], (B, N, N))
    parts = [obs[:, t].reshape(B * N, -1), last.reshape(B * N, -1), eye.reshape(B * N, -1)]
    return np.concatenate(parts, axis=1)


# --------------------------------------------------------------------------
# a2: DRQNAgentNetwork.forward   marl/modules/agents/drqn_agent.py:29-35
# --------------------------------------------------------------------------
def drqn_step(p, inp, h):
    """One fc1->relu->GRUCell->fc2 step.  Returns q [R,A], h' [R,H] and the cache the backward needs."""
    Hh = p["gru.weight_hh"].shape[1]
    x_pre = inp @ p["fc1.weight"].T + p["fc1.bias"]
    x = np.maximum(x_pre, 0)
    gi = x @ p["gru.weight_ih"].T + p["gru.bias_ih"]
    gh = h @ p["gru.weight_hh"].T + p["gru.bias_hh"]
    r = _sigmoid(gi[:, :Hh] + gh[:, :Hh])
    z = _sigmoid(gi[:, Hh:2 * Hh] + gh[:, Hh:2 * Hh])
    n = np.tanh(gi[:, 2 * Hh:] + r * gh[:, 2 * Hh:])
    h_new = n + z * (h - n)  # == (1-z)*n + z*h, the form torch's CPU gru_cell evaluates
    q = h_new @ p["fc2.weight"].T + p["fc2.bias"]
    cache = dict(inp=inp, x=x, x_pre=x_pre, h_prev=h, r=r, z=z, n=n, ghn=gh[:, 2 * Hh:], h=h_new)
    return q, h_new, cache


def dqn_agent_param_shapes(d_in: int, n_actions: int, hidden: int = H_DEFAULT) -> "OrderedDict[str, tuple]":
    """Keys/shapes of DQNAgentNetwork.state_dict(), marl/modules/agents/dqn_agent.py:24-25."""
    return OrderedDict([("fc1.weight", (hidden, d_in)), ("fc1.bias", (hidden,)),
                        ("fc2.weight", (n_actions, hidden)), ("fc2.bias", (n_actions,))])


def dqn_step(p, inp, h=None):
    """DQNAgentNetwork.forward, marl/modules/agents/dqn_agent.py:34-37: q = fc2(relu(fc1(inputs))); the hidden state is
    passed through untouched.  The cache mirrors drqn_step's (h := x, the activation fc2 reads)."""
    x_pre = inp @ p["fc1.weight"].T + p["fc1.bias"]
    x = np.maximum(x_pre, 0)
    q = x @ p["fc2.weight"].T + p["fc2.bias"]
    return q, h, dict(inp=inp, x=x, x_pre=x_pre, h=x)


def is_dqn(p):
    return "gru.weight_hh" not in p


def unroll(p, obs, actions_onehot):
    """QLearner.train's time loop, marl/learners/q_learner.py:46-52 (and :58-62 for the target net).

    Returns mac_out [B,TT,N,A] and the per-step caches."""
    B, TT, N, _ = obs.shape
    if is_dqn(p):
        outs, caches = [], []
        for t in range(TT):
            q, _, c = dqn_step(p, build_inputs(obs, actions_onehot, t))
            outs.append(q.reshape(B, N, -1))
            caches.append(c)
        return np.stack(outs, axis=1), caches
    Hh = p["gru.weight_hh"].shape[1]
    h = np.zeros((B * N, Hh), dtype=obs.dtype)  # init_hidden: basic_controller.py:59-60, drqn_agent.py:25-27
    outs, caches = [], []
    for t in range(TT):
        q, h, c = drqn_step(p, build_inputs(obs, actions_onehot, t), h)
        outs.append(q.reshape(B, N, -1))
        caches.append(c)
    return np.stack(outs, axis=1), caches


# --------------------------------------------------------------------------
# a5: masked (double-Q) target max   marl/learners/q_learner.py:65-78
# --------------------------------------------------------------------------
def masked_target_max(mac_out, target_mac_out_full, avail, double_q: bool, argmax_override=None):
    """mac_out/target_mac_out_full [B,TT,N,A], avail [B,TT,N,A] int -> (target_max [B,T,N], argmax [B,T,N] int64).

    Ties resolve to the lowest action index (np.argmax == torch.max first occurrence); NaN wins the max, as in torch.
    ``argmax_override`` (tests only, double-Q): use these indices instead of the oracle's own arg-max -- a checker that
    has verified that an fp32 implementation flipped a genuine near-tie re-evaluates everything downstream of the
    discrete choice on the implementation's indices ("discrete argmax inside the loss")."""
    tq = target_mac_out_full[:, 1:].copy()
    tq[avail[:, 1:] == 0] = NEG_MASK
    if double_q:
        oq = mac_out.copy()
        oq[avail == 0] = NEG_MASK
        amax = np.argmax(oq[:, 1:], axis=3) if argmax_override is None else np.asarray(argmax_override, dtype=np.int64)
        tmax = np.take_along_axis(tq, amax[..., None], axis=3)[..., 0]
    else:
        amax = np.argmax(tq, axis=3)
        tmax = tq.max(axis=3)
    return tmax, amax.astype(np.int64)


# --------------------------------------------------------------------------
# a6: mixers   marl/modules/mixers/qmix.py:41-59, marl/modules/mixers/vdn.py:9-10
# --------------------------------------------------------------------------
def _hyper(mp, name, s):
    """Evaluate a 1- or 2-layer hypernet branch; returns (out, hidden_or_None, hidden_pre_activation_or_None)."""
    if name + ".0.weight" in mp:
        pre = s @ mp[name + ".0.weight"].T + mp[name + ".0.bias"]
        hid = np.maximum(pre, 0)
        return hid @ mp[name + ".2.weight"].T + mp[name + ".2.bias"], hid, pre
    return s @ mp[name + ".weight"].T + mp[name + ".bias"], None, None


def qmix_forward(mp, agent_qs, states):
    """agent_qs [B,T,N], states [B,T,S] -> q_tot [B,T,1] and cache."""
    B, T, N = agent_qs.shape
    E = mp["hyper_b_1.weight"].shape[0]
    s = states.reshape(B * T, -1)
    q = agent_qs.reshape(B * T, N)
    a1, h1, h1_pre = _hyper(mp, "hyper_w_1", s)
    w1 = np.abs(a1).reshape(B * T, N, E)
    b1 = s @ mp["hyper_b_1.weight"].T + mp["hyper_b_1.bias"]
    pre = np.einsum("mn,mne->me", q, w1) + b1
    hidden = np.where(pre > 0, pre, np.expm1(np.minimum(pre, 0)))  # F.elu, alpha=1
    af, hf, hf_pre = _hyper(mp, "hyper_w_final", s)
    wf = np.abs(af)
    v1_pre = s @ mp["V.0.weight"].T + mp["V.0.bias"]
    v1 = np.maximum(v1_pre, 0)
    v = v1 @ mp["V.2.weight"].T + mp["V.2.bias"]
    y = (hidden * wf).sum(axis=1, keepdims=True) + v
    cache = dict(s=s, q=q, a1=a1, h1=h1, w1=w1, b1=b1, pre=pre, hidden=hidden, af=af, hf=hf, wf=wf, v1=v1,
                 h1_pre=h1_pre, hf_pre=hf_pre, v1_pre=v1_pre)
    return y.reshape(B, T, 1), cache


def qmix_backward(mp, cache, g, discrete=None):
    """g = dL/dq_tot [B,T,1] -> (grads dict keyed like the state_dict, dq [B,T,N]).

    ``discrete`` (tests only): dict with any of ``a1_sign`` [M,N*E], ``af_sign`` [M,E], ``h1_mask`` / ``hf_mask`` [M,HE],
    ``v1_mask`` [M,E] -- the derivative of abs / ReLU at these elements is taken from the caller instead of from the
    oracle's own pre-activations (see learner_forward_backward)."""
    c = cache
    M, N, E = c["w1"].shape
    g = g.reshape(M, 1)
    grads = OrderedDict()
    discrete = discrete or {}
    sign_a1 = discrete.get("a1_sign", np.sign(c["a1"])).astype(c["a1"].dtype)
    sign_af = discrete.get("af_sign", np.sign(c["af"])).astype(c["af"].dtype)
    hid_mask = {"hyper_w_1": discrete.get("h1_mask"), "hyper_w_final": discrete.get("hf_mask")}
    v1_mask = discrete.get("v1_mask", c["v1"] > 0)
    # y = sum_e hidden*wf + v
    dhidden = g * c["wf"]
    dwf = g * c["hidden"]
    # V = Linear(S,E)-ReLU-Linear(E,1)
    grads["V.2.weight"] = (g * c["v1"]).sum(axis=0, keepdims=True)
    grads["V.2.bias"] = g.sum(axis=0)
    dv1 = (g @ mp["V.2.weight"]) * v1_mask
    grads["V.0.weight"] = dv1.T @ c["s"]
    grads["V.0.bias"] = dv1.sum(axis=0)
    # elu
    dpre = dhidden * np.where(c["pre"] > 0, 1.0, np.exp(np.minimum(c["pre"], 0)))
    dq = np.einsum("me,mne->mn", dpre, c["w1"])
    dw1 = c["q"][:, :, None] * dpre[:, None, :]
    grads["hyper_b_1.weight"] = dpre.T @ c["s"]
    grads["hyper_b_1.bias"] = dpre.sum(axis=0)
    da1 = (dw1 * sign_a1.reshape(M, N, E)).reshape(M, N * E)
    daf = dwf * sign_af

    def hyper_back(name, dout, hid):
        if hid is not None:
            grads[name + ".2.weight"] = dout.T @ hid
            grads[name + ".2.bias"] = dout.sum(axis=0)
            dh = (dout @ mp[name + ".2.weight"]) * (hid > 0 if hid_mask[name] is None else hid_mask[name])
            grads[name + ".0.weight"] = dh.T @ c["s"]
            grads[name + ".0.bias"] = dh.sum(axis=0)
        else:
            grads[name + ".weight"] = dout.T @ c["s"]
            grads[name + ".bias"] = dout.sum(axis=0)

    hyper_back("hyper_w_1", da1, c["h1"])
    hyper_back("hyper_w_final", daf, c["hf"])
    ordered = OrderedDict((k, grads[k].reshape(mp[k].shape)) for k in mp)
    return ordered, dq


def vdn_forward(agent_qs):
    return agent_qs.sum(axis=2, keepdims=True)


# --------------------------------------------------------------------------
# a7: mask, TD target, masked L2 loss   marl/learners/q_learner.py:36-43, 85-98
# --------------------------------------------------------------------------
def make_mask(filled, terminated):
    """filled [B,TT,1] int, terminated [B,TT,1] uint8 -> (mask [B,T,1], terminated_f [B,T,1]) as float."""
    term = terminated[:, :-1].astype(np.float32)
    mask = filled[:, :-1].astype(np.float32).copy()
    mask[:, 1:] = mask[:, 1:] * (1 - term[:, :-1])
    return mask, term


def td_loss(q_tot, target_q_tot, rewards, term, mask, gamma):
    """Returns loss, td_error, masked_td_error, targets, dL/dq_tot."""
    dt = q_tot.dtype
    targets = rewards.astype(dt) + dt.type(gamma) * (1 - term.astype(dt)) * target_q_tot
    td = q_tot - targets
    m = np.broadcast_to(mask.astype(dt), td.shape)
    mtd = td * m
    msum = m.sum(dtype=dt)
    loss = (mtd ** 2).sum(dtype=dt) / msum
    g = 2.0 * mtd * m / msum
    return loss, td, mtd, targets, g


# --------------------------------------------------------------------------
# a4+a8: full learner step forward + backward   marl/learners/q_learner.py:34-105
# --------------------------------------------------------------------------
def agent_backward(p, caches, dq_all, x_mask=None):
    """BPTT through the unroll.  dq_all [B,TT,N,A] = dL/d mac_out.  Returns grads keyed like the state_dict.
    ``x_mask`` (tests only) [TT,R,H] bool: ReLU derivative of fc1's output taken from the caller."""
    g = OrderedDict((k, np.zeros_like(v)) for k, v in p.items())
    TT = len(caches)
    R = caches[0]["h"].shape[0]
    if is_dqn(p):                                   # feed-forward agent: no recurrence, every step independent
        for t in range(TT):
            c = caches[t]
            dq = dq_all[:, t].reshape(R, -1)
            g["fc2.weight"] += dq.T @ c["x"]
            g["fc2.bias"] += dq.sum(axis=0)
            dx = (dq @ p["fc2.weight"]) * (c["x"] > 0 if x_mask is None else x_mask[t])
            g["fc1.weight"] += dx.T @ c["inp"]
            g["fc1.bias"] += dx.sum(axis=0)
        return g
    Hh = p["gru.weight_hh"].shape[1]
    dh_next = np.zeros((R, Hh), dtype=caches[0]["h"].dtype)
    Whh, Wih = p["gru.weight_hh"], p["gru.weight_ih"]
    for t in range(TT - 1, -1, -1):
        c = caches[t]
        dq = dq_all[:, t].reshape(R, -1)
        g["fc2.weight"] += dq.T @ c["h"]
        g["fc2.bias"] += dq.sum(axis=0)
        dh = dh_next + dq @ p["fc2.weight"]
        dn = dh * (1 - c["z"])
        dz = dh * (c["h_prev"] - c["n"])
        dn_pre = dn * (1 - c["n"] ** 2)
        dr = dn_pre * c["ghn"]
        dz_pre = dz * c["z"] * (1 - c["z"])
        dr_pre = dr * c["r"] * (1 - c["r"])
        dgi = np.concatenate([dr_pre, dz_pre, dn_pre], axis=1)
        dgh = np.concatenate([dr_pre, dz_pre, dn_pre * c["r"]], axis=1)
        g["gru.weight_ih"] += dgi.T @ c["x"]
        g["gru.bias_ih"] += dgi.sum(axis=0)
        g["gru.weight_hh"] += dgh.T @ c["h_prev"]
        g["gru.bias_hh"] += dgh.sum(axis=0)
        dh_next = dh * c["z"] + dgh @ Whh
        dx = (dgi @ Wih) * (c["x"] > 0 if x_mask is None else x_mask[t])
        g["fc1.weight"] += dx.T @ c["inp"]
        g["fc1.bias"] += dx.sum(axis=0)
    return g


def learner_forward_backward(agent_p, target_agent_p, mixer_p, target_mixer_p, batch, *, mixer: str,
                             double_q: bool, gamma: float, dtype=np.float32, argmax_override=None, discrete=None):
    """Everything QLearner.train computes up to and including loss.backward() (q_learner.py:34-103).

    ``batch`` is a dict of numpy arrays with the reference's keys/shapes
    (state, obs, actions, avail_actions, reward, terminated, actions_onehot, filled), already truncated in time.
    ``mixer`` in {"qmix", "vdn"}.  Returns a dict of intermediates and gradients.

    ``argmax_override`` / ``discrete`` (tests only): the loss contains DISCRETE choices -- the double-Q arg-max, the ReLU
    derivatives of fc1 / the hypernet hidden layers / V, and sign() from abs() of the hypernet outputs.  An fp32
    implementation may legitimately take the other branch where the fp64 pre-activation lies within rounding distance of
    the discontinuity (a handful of the 10^7 elements of a full-size batch; one flipped element moves a weight-gradient
    tensor by ~1e-4 norm-wise).  A checker that has VERIFIED that every differing choice sits on such a near-tie passes the
    implementation's choices here and compares everything downstream at the 1e-5 tolerance.
    ``discrete`` keys: ``x_mask`` [TT,R,H] and the keys of qmix_backward."""
    cast = lambda d: OrderedDict((k, v.astype(dtype)) for k, v in d.items()) if d is not None else None
    ap, tp, mp, tmp = cast(agent_p), cast(target_agent_p), cast(mixer_p), cast(target_mixer_p)
    obs = batch["obs"].astype(dtype)
    onehot = batch["actions_onehot"].astype(dtype)
    state = batch["state"].astype(dtype)
    rewards = batch["reward"][:, :-1].astype(dtype)
    actions = batch["actions"][:, :-1]
    avail = batch["avail_actions"]
    mask, term = make_mask(batch["filled"], batch["terminated"])
    mask, term = mask.astype(dtype), term.astype(dtype)

    mac_out, caches = unroll(ap, obs, onehot)
    chosen = np.take_along_axis(mac_out[:, :-1], actions, axis=3)[..., 0]          # :55
    target_full, _ = unroll(tp, obs, onehot)
    tmax, amax = masked_target_max(mac_out, target_full, avail, double_q, argmax_override)   # :65-78

    if mixer == "qmix":
        q_tot, mc = qmix_forward(mp, chosen, state[:, :-1])                         # :82
        tq_tot, _ = qmix_forward(tmp, tmax, state[:, 1:])                           # :83
    elif mixer == "vdn":
        q_tot, tq_tot, mc = vdn_forward(chosen), vdn_forward(tmax), None
    else:
        raise ValueError("Mixer {} not recognised.".format(mixer))                 # :24

    loss, td, mtd, targets, g = td_loss(q_tot, tq_tot, rewards, term, mask, gamma)  # :86-98

    if mixer == "qmix":
        mixer_grads, dchosen = qmix_backward(mp, mc, g, discrete)
        dchosen = dchosen.reshape(chosen.shape)
    else:
        mixer_grads, dchosen = OrderedDict(), np.broadcast_to(g, chosen.shape).copy()
    dmac = np.zeros_like(mac_out)
    np.put_along_axis(dmac[:, :-1], actions, dchosen[..., None], axis=3)
    agent_grads = agent_backward(ap, caches, dmac, (discrete or {}).get("x_mask"))
    hout = np.stack([c["h"] for c in caches], axis=0)  # [TT, R, H]
    msum = mask.sum(dtype=dtype)
    n_agents = obs.shape[2]
    stats = dict(  # q_learner.py:117-124
        loss=loss, td_error_abs=np.abs(mtd).sum(dtype=dtype) / msum,
        q_taken_mean=(q_tot * mask).sum(dtype=dtype) / (msum * n_agents),
        target_mean=(targets * mask).sum(dtype=dtype) / (msum * n_agents),
        mask_sum=msum, trained_steps=int(np.count_nonzero(np.broadcast_to(mask, td.shape))),
    )
    # pre-activations at the discontinuities (ReLU inputs, abs inputs): see `discrete`
    pre = dict(x=np.stack([c["x_pre"] for c in caches], axis=0))      # [TT, R, H]
    if mixer == "qmix":
        pre.update(a1=mc["a1"], af=mc["af"], h1=mc["h1_pre"], hf=mc["hf_pre"], v1=mc["v1_pre"])
    return dict(pre=pre, mac_out=mac_out, target_mac_out=target_full, hout=hout, chosen=chosen, target_max=tmax,
                argmax=amax, q_tot=q_tot, target_q_tot=tq_tot, targets=targets, td=td, mask=mask, loss=loss,
                dq_tot=g, dchosen=dchosen, agent_grads=agent_grads, mixer_grads=mixer_grads, stats=stats)


def clip_grad_norm(grads, max_norm: float):
    """th.nn.utils.clip_grad_norm_ at q_learner.py:104: global L2 norm, coef = max_norm/(norm+1e-6) clamped to 1."""
    dt = grads[0].dtype
    total = np.sqrt(sum((g.astype(dt) ** 2).sum(dtype=dt) for g in grads))
    coef = min(dt.type(max_norm) / (total + dt.type(1e-6)), dt.type(1.0))
    return total, [g * coef for g in grads]


def rmsprop_update(p, g, sq, lr, alpha, eps):
    """torch.optim.RMSprop as configured in marl/learners/learner.py:25-31 (no momentum, not centered).

    v = alpha*v + (1-alpha)*g^2 ; p -= lr * g / (sqrt(v) + eps)."""
    dt = p.dtype
    sq_new = dt.type(alpha) * sq + dt.type(1 - alpha) * g * g
    return p - dt.type(lr) * g / (np.sqrt(sq_new) + dt.type(eps)), sq_new


# --------------------------------------------------------------------------
# a12: epsilon-greedy selection   marl/components/action_selectors.py:44-62, epsilon_schedules.py:21-25
# --------------------------------------------------------------------------
def epsilon_linear(start, finish, anneal_time, t_env):
    delta = (start - finish) / anneal_time
    return max(finish, start - delta * t_env)


def eps_greedy_select(q, avail, epsilon: float, u, e):
    """q [bs,N,A] f32, avail [bs,N,A] int, u [bs,N] f32 uniforms, e [bs*N,A] f32 Exp(1) draws.

    Categorical(avail.float()).sample() == argmax_a (p_a / e_a) with p = avail / sum(avail)
    (torch.multinomial's single-sample path); ties -> lowest index.  Returns (picked int64, pick_greedy int64)."""
    bs, N, A = q.shape
    masked = q.astype(np.float32).copy()
    masked[avail == 0] = -np.inf
    # torch compares the fp32 tensor with the Python double after casting the scalar to fp32
    pick_random = (u.astype(np.float32) < np.float32(epsilon)).astype(np.int64)
    av = avail.astype(np.float32)
    if np.any(av.sum(axis=-1) <= 0):
        raise ValueError("Categorical needs at least one available action per agent")
    p = av / av.sum(axis=-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = p.reshape(bs * N, A) / e.astype(np.float32).reshape(bs * N, A)
    rand_act = np.argmax(ratio, axis=1).reshape(bs, N).astype(np.int64)
    greedy = np.argmax(masked, axis=2).astype(np.int64)
    pick_greedy = 1 - pick_random
    return pick_random * rand_act + pick_greedy * greedy, pick_greedy


# --------------------------------------------------------------------------
# a14: ReplayBuffer ring semantics   marl/components/replay_buffers/replay_buffer.py:22-53
# --------------------------------------------------------------------------
def ring_insert_slots(buffer_index: int, episodes_in_buffer: int, buffer_size: int, n: int):
    """Slots written by insert_episode_batch for n episodes (wrap-around split :36-41) and the new counters."""
    slots = []
    idx, filled = buffer_index, episodes_in_buffer
    left = n
    while left > 0:
        take = min(left, buffer_size - idx)
        slots.extend(range(idx, idx + take))
        idx += take
        filled = max(filled, idx)
        idx %= buffer_size
        left -= take
    return np.asarray(slots, dtype=np.int64), idx, filled


def max_t_filled(filled):
    """EpisodeBatch.max_t_filled, marl/components/episode_batch.py:240-242."""
    return int(filled.sum(axis=1).max())


def onehot(actions, n_actions: int):
    """OneHot.transform, marl/components/transforms.py:16-19."""
    out = np.zeros(actions.shape[:-1] + (n_actions,), dtype=np.float32)
    np.put_along_axis(out, actions.astype(np.int64), 1.0, axis=-1)
    return out
