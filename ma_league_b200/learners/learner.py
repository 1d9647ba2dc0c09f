"""Learner base (reference: marl/learners/learner.py:6-78): name prefix + RMSprop construction."""
import torch as th
from torch.optim import RMSprop

from .. import _native as nat
from ..flat import flat_views


class FusedRMSprop(RMSprop):
    """torch.optim.RMSprop (lr/alpha/eps as in learner.py:25-31; no momentum, not centered) whose state lives in ONE
    flat square_avg buffer and whose step is the fused clip+RMSprop kernel.  state_dict()/load_state_dict() keep
    torch's format, so the reference's `..opt.th` checkpoints (q_learner.py:133-147) interoperate."""

    def __init__(self, params, lr, alpha, eps):
        params = list(params)
        super().__init__(params, lr=lr, alpha=alpha, eps=eps)
        self._plist = params
        self._steps = 0
        self.flat_sq = None
        self._bind(None)

    def _bind(self, loaded):
        total = sum(p.numel() for p in self._plist)
        dev = self._plist[0].device
        self.flat_sq = th.zeros(total, dtype=th.float32, device=dev)
        views = flat_views(self.flat_sq, self._plist)
        for p, v in zip(self._plist, views):
            st = self.state[p]
            if loaded and "square_avg" in st:
                v.copy_(st["square_avg"].to(dev))
                self._steps = max(self._steps, int(st.get("step", th.tensor(0.)).item()))
            st["square_avg"] = v
            st["step"] = th.tensor(float(self._steps))

    def state_dict(self):
        for p in self._plist:
            self.state[p]["step"] = th.tensor(float(self._steps))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._bind(True)

    def step(self, closure=None):
        raise nat.MalError("FusedRMSprop is stepped by QLearner.train (clip + RMSprop run inside mal_learner_step)")


class Learner:
    def __init__(self, mac, scheme, logger, args, name=None):
        self.mac = mac
        self.scheme = scheme
        self.logger = logger
        self.args = args
        self.name = f'{"" if name is None else name}_{self.__class__.__name__.lower()}_'
        self.log_stats_t = -self.args.learner_log_interval - 1
        self.optimiser = None

    def build_optimizer(self):
        self.optimiser = FusedRMSprop(params=self.parameters(), lr=self.args.lr, alpha=self.args.optim_alpha,
                                      eps=self.args.optim_eps)

    def parameters(self):
        raise NotImplementedError()

    def train(self, batch, t_env: int, episode_num: int) -> None:
        raise NotImplementedError()

    def cuda(self) -> None:
        raise NotImplementedError()

    def save_models(self, path, name):
        raise NotImplementedError()

    def load_models(self, path):
        raise NotImplementedError()

    def update_targets(self):
        pass
