"""Batched, device-resident rollout loop: the caller either side of `select_actions`.

Restates `EpisodeStepper.run` (steppers/episode_stepper.py:86-186) and `SelfPlayStepper.run`
(steppers/self_play_stepper.py:44-147) for `batch_size_run = B >= 1` matches that advance in lock-step, which is what the
reference's `ParallelStepper` was meant to do (it does not run).  Per timestep and team the
episode batch receives one pre-transition update (state / avail_actions / obs), the controller's fused act-select
launch picks the actions of all B x N agents, and one post-transition update stores actions / reward / terminated --
all on device tensors: no Python lists, no host round trip inside the step (the reference converts lists to tensors
twice per step, episode_stepper.py:177-199).

A match that has ended leaves nothing behind: with `running[t]` = "match still alive at step t",
    state / avail_actions / obs / filled at index t   are written iff running[t-1]   (index 0 always)
    actions                                 at index t   iff running[t-1]   (the last stored state gets actions too, :156-165)
    reward / terminated                     at index t   iff running[t]
and everything else stays zero, so every row of the returned batch is exactly the batch a single-match run of the
reference loop would have produced for that match.  The loop leaves as soon as every match has ended; that check is
the only host synchronisation (every `sync_every` steps).
"""
from functools import partial

import torch as th

from ..components.episode_batch import EpisodeBatch


class BatchedEpisodeStepper:
    def __init__(self, args, logger, env, log_start_t=0, sync_every=1, fuse=True):
        self.args = args
        self.logger = logger
        self.env = env
        self.batch_size = env.n_envs
        self.episode_limit = env.episode_limit
        self.n_teams = env.n_teams
        self.t = 0
        self.t_env = 0
        self.log_start_t = log_start_t
        self.sync_every = max(1, int(sync_every))
        self.fuse = bool(fuse)
        self.home_mac = self.away_mac = None
        self.home_batch = self.away_batch = None
        self.new_batch_fn = None
        self.is_initalized = False          # (sic) attribute name of the reference's EnvStepper

    def initialize(self, scheme, groups, preprocess, home_mac, away_mac=None):
        self.new_batch_fn = partial(EpisodeBatch, scheme, groups, self.batch_size, self.episode_limit + 1,
                                    preprocess=preprocess, device=self.args.device)
        self.home_mac, self.away_mac = home_mac, away_mac
        if self.n_teams == 2 and away_mac is None:
            raise ValueError("a two-team environment needs an away controller (self-play)")
        self.is_initalized = True

    def get_env_info(self):
        return self.env.get_env_info()

    def close_env(self):
        self.env.close()

    @property
    def epsilon(self):
        return getattr(self.home_mac.action_selector, "epsilon", None)

    @property
    def epsilons(self):
        return [getattr(m.action_selector, "epsilon", None) for m in self._macs()]

    @property
    def log_t(self):
        return self.log_start_t + self.t_env

    def _macs(self):
        return [self.home_mac] if self.n_teams == 1 else [self.home_mac, self.away_mac]

    def reset(self):
        self.home_batch = self.new_batch_fn()
        self.away_batch = self.new_batch_fn() if self.n_teams == 2 else None
        self.env.reset()
        self.t = 0

    def run(self, test_mode=False, draws=None):
        """Run B matches to their ends.  Returns (home_batch, env_info) or (home_batch, away_batch, env_info).
        `draws` (tests): callable (team, t) -> (u [B, N], e [B*N, A]) of injected random draws for the selector."""
        if self.home_mac is None:
            raise RuntimeError("MultiAgentControllerNotInitialized")     # exceptions/runner_exceptions.py
        self.reset()
        macs = self._macs()
        batches = [self.home_batch] if self.n_teams == 1 else [self.home_batch, self.away_batch]
        B, dev = self.batch_size, self.home_batch["filled"].device
        for m in macs:
            m.init_hidden(batch_size=B)
        running_prev = th.ones(B, dtype=th.bool, device=dev)       # running[t-1]; every match stores index 0
        running = th.ones(B, dtype=th.bool, device=dev)            # running[t]
        returns = [th.zeros(B, device=dev) for _ in macs]
        steps = th.zeros(B, dtype=th.long, device=dev)
        env_info = {}
        views = [bt.data.transition_data for bt in batches]        # strided views into the packed episode records
        # one launch per team and timestep when the controller offers the fused rollout step (pre-transition update +
        # previous step's reward / terminated + act-select + actions / one-hot), else the separate calls
        fused = [self.fuse and hasattr(m, "rollout_step") and m._rollout_fusable() and bt._layout is not None
                 for m, bt in zip(macs, batches)]
        prev_rewards, prev_done = None, None
        t = 0
        while True:
            acts = []
            for k, (mac, batch, tv) in enumerate(zip(macs, batches, views)):
                pre = self.env.observe(k)
                kw = {}
                if draws is not None:
                    kw["u"], kw["e"] = draws(k, t)
                if fused[k]:
                    # every match writes index t unconditionally; a match that has ended selects on a dummy avail row and
                    # what lies past its end is cleared once, after the loop
                    a, _ = mac.rollout_step(batch, t, self.t_env, pre, prev=(prev_rewards[k], prev_done) if t > 0 else None,
                                            alive=running_prev, test_mode=test_mode, **kw)
                else:
                    tv["state"][:, t].copy_(pre["state"])
                    avail = pre["avail_actions"]
                    if t > 0:                                        # ended matches: a valid dummy row (never all-zero)
                        avail = avail.clone()
                        avail[:, :, 0] |= (~running_prev).to(avail.dtype).view(B, 1)
                    tv["avail_actions"][:, t].copy_(avail)
                    tv["obs"][:, t].copy_(pre["obs"])
                    a, _ = mac.select_actions(batch, t_ep=t, t_env=self.t_env, test_mode=test_mode, **kw)
                    tv["actions"][:, t].copy_(a.view_as(tv["actions"][:, t]))
                    for key, (new_key, transforms) in batch.preprocess.items():       # actions -> actions_onehot
                        v = tv[key][:, t]
                        for tr in transforms:
                            v = tr.transform(v)
                        tv[new_key][:, t].copy_(v)
                    if t > 0:
                        tv["reward"][:, t - 1, 0].copy_(prev_rewards[k])
                        tv["terminated"][:, t - 1, 0].copy_(prev_done)
                acts.append(a)
            if t == self.episode_limit:
                break
            # NB the environment keeps being stepped for matches that have already ended (lock-step batch); their
            # outcome is ignored: `running` masks it out of the returns and the data is cleared below
            rewards, done, env_info = self.env.step(acts)
            for k in range(len(macs)):
                returns[k] += rewards[k] * running
            steps += running
            prev_rewards, prev_done = rewards, done
            running_prev = running
            running = running & ~done
            t += 1
            # the iteration after the last match ended still writes that match's last stored state and its actions;
            # once no match was alive at the previous step there is nothing left to write
            if t % self.sync_every == 0 and not bool(running_prev.any()):
                break
        # a match of L transitions keeps indices 0..L of the pre-transition keys / actions / filled and 0..L-1 of
        # reward / terminated; the rest goes back to the zeros of an untouched batch
        TT = self.episode_limit + 1
        idx = th.arange(TT, device=dev).view(1, TT)
        keep_state = idx <= steps.view(B, 1)
        keep_trans = idx < steps.view(B, 1)
        for tv in views:
            for key, v in tv.items():
                m = keep_trans if key in ("reward", "terminated") else keep_state
                v.masked_fill_(~m.view((B, TT) + (1,) * (v.dim() - 2)), 0)      # (0 * NaN would stay NaN)
            tv["filled"][:, :, 0].copy_(keep_state)
        n_steps = int(steps.sum())
        self.t = int(steps.max())                 # transitions of the longest match (the reference's self.t for one match)
        if not test_mode:
            self.t_env += n_steps
        if self.logger is not None and hasattr(self.logger, "log_stat"):
            self.logger.log_stat("home_epsilon", self.epsilon, self.log_t)
            self.logger.log_stat("home_return_mean", float(returns[0].mean()), self.log_t)
            self.logger.log_stat("ep_length_mean", n_steps / B, self.log_t)
        env_info = dict(env_info)
        env_info.update(episode_returns=returns, episode_steps=steps)
        if self.n_teams == 1:
            return self.home_batch, env_info
        return self.home_batch, self.away_batch, env_info
