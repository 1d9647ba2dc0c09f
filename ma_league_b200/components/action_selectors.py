"""Epsilon-greedy action selection (reference: marl/components/action_selectors.py:35-68) as one kernel launch.

RNG contract: the selector draws with Philox4x32-10 using the SAME counter layout torch's
CUDA kernels use for ``th.rand_like(q[:, :, 0])`` followed by ``Categorical(avail).sample()`` (= ``exponential_`` on
``[bs*N, A]``), taken from and advancing torch's default CUDA generator -- so under ``torch.manual_seed`` the chosen
actions are bit-identical to the reference running on the same GPU.  For CPU-oracle tests the draws can be injected
(``select(..., u=..., e=...)``).
"""
import ctypes as C

import torch as th

from .. import _native as nat
from .epsilon_schedules import DecayThenFlatSchedule


class Selector:
    def select(self, agent_inputs, avail_actions, t_env, test_mode=False):
        raise NotImplementedError()


_ADVANCE_CACHE = {}


def philox_advance(rows, n_actions, device_index):
    key = (rows, n_actions, device_index)
    adv = _ADVANCE_CACHE.get(key)
    if adv is None:
        out = C.c_uint64(0)
        nat.check(nat.lib().mal_select_philox_advance(rows, n_actions, C.byref(out)), "mal_select_philox_advance")
        adv = _ADVANCE_CACHE[key] = out.value
    return adv


def make_select_struct(avail, epsilon, actions, greedy, status, u=None, e=None, generator=None):
    """Fill the C `mal_select_t` for avail [bs, N, A] (inner dims contiguous)."""
    bs, N, A = avail.shape
    if avail.dtype != th.int32:
        avail = avail.to(th.int32)
    if avail.stride(2) != 1 or avail.stride(1) != A:
        avail = avail.contiguous()
    s = nat.Select()
    s.avail = avail.data_ptr()
    s.avail_sb = avail.stride(0)
    s.epsilon = float(epsilon)
    s.actions = actions.data_ptr()
    s.greedy = greedy.data_ptr()
    s.status = status.data_ptr() if status is not None else None
    keep = [avail]
    if u is not None or e is not None:
        if u is None or e is None:
            raise ValueError("inject both u and e or neither")
        u = u.to(device=avail.device, dtype=th.float32).contiguous()
        e = e.to(device=avail.device, dtype=th.float32).contiguous()
        assert u.numel() == bs * N and e.numel() == bs * N * A
        s.rng_mode, s.u, s.e = 0, u.data_ptr(), e.data_ptr()
        keep += [u, e]
    else:
        gen = generator if generator is not None else th.cuda.default_generators[avail.device.index]
        off = gen.get_offset()
        s.rng_mode, s.seed, s.offset = 1, gen.initial_seed(), off
        gen.set_offset(off + philox_advance(bs * N, A, avail.device.index))
    return s, keep


class EpsilonGreedyActionSelector(Selector):
    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time,
                                              decay="linear")
        self.epsilon = self.schedule.eval(0)
        # reference behaviour: Categorical raises ValueError on an all-zero avail row (costs a device sync)
        self.validate = getattr(args, "validate_avail", True)

    def _epsilon(self, t_env, test_mode):
        self.epsilon = self.schedule.eval(t_env)
        if test_mode:
            self.epsilon = 0.0   # the random draws below are still consumed, as in action_selectors.py:48-58
        return self.epsilon

    def select(self, agent_outputs, avail_actions, t_env, test_mode=False, u=None, e=None):
        """agent_outputs [bs,N,A] f32, avail_actions [bs,N,A] -> (picked_actions [bs,N] i64, pick_greedy [bs,N] i64)."""
        eps = self._epsilon(t_env, test_mode)
        q = nat.require_cuda(agent_outputs, "agent_outputs")
        bs, N, A = q.shape
        if q.dtype != th.float32 or not q.is_contiguous():
            q = q.float().contiguous()
        actions = th.empty(bs, N, dtype=th.long, device=q.device)
        greedy = th.empty(bs, N, dtype=th.long, device=q.device)
        status = th.zeros(1, dtype=th.int32, device=q.device) if self.validate else None
        s, keep = make_select_struct(avail_actions, eps, actions, greedy, status, u, e)
        with nat.on_device(q.device):
            nat.check(nat.lib().mal_eps_greedy_select(nat.ptr(q), A, bs * N, N, A, C.byref(s),
                                                      nat.current_stream(q.device)), "mal_eps_greedy_select")
        if status is not None and int(status.item()) != 0:
            raise ValueError("Expected at least one available action per agent (Categorical probs are all zero)")
        return actions, greedy


REGISTRY = {"epsilon_greedy": EpsilonGreedyActionSelector}
