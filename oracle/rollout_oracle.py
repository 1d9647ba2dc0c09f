"""CPU restatement of the reference's rollout loops (the callers either side of ``select_actions``).

TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py for the rules): only ``tests/`` import this module.

    EpisodeStepper.run       steppers/episode_stepper.py:86-186   (one policy team against the environment)
    SelfPlayStepper.run      steppers/self_play_stepper.py:44-147 (home and away controllers in one match)
    build_pre_transition_data  steppers/utils/stepper_utils.py:4-25

The reference's steppers cannot be imported here (they need ``maenv`` / ``sacred`` / ``bin.controls``, SURVEY.md 8c iii),
so the loop is restated line by line on plain numpy episode dictionaries: per step the pre-transition data
(state / avail_actions / obs) is written at index t, the controller selects actions on that step
(np_oracle.drqn_step + np_oracle.eps_greedy_select with injected draws), the environment steps, and the
post-transition data (actions / reward / terminated) is written at index t; after the terminal step the last state is
stored and actions are selected in it once more (episode_stepper.py:156-165).  ``filled`` is set wherever ``update``
writes (episode_batch.py:163-166) and ``actions_onehot`` is the OneHot transform of ``actions`` (transforms.py:16-19).

``PlaybackEnv`` offers the reference's environment calls (get_state / get_avail_actions / get_obs / step) for ONE
match over pre-recorded arrays, with the agents of all teams concatenated as the reference's environments do.
"""
from __future__ import annotations

import numpy as np

from . import np_oracle as O


class PlaybackEnv:
    """One match of a seeded playback environment: `obs [teams, TT, N, OBS]`, `state [teams, TT, S]`,
    `avail [teams, TT, N, A]`, `noise [teams, TT]`, `length` = number of transitions until `done`.
    reward_k(t) = noise[k, t] + 0.01 * sum(actions of team k)   (float32, like the device environment)."""

    def __init__(self, obs, state, avail, noise, length, episode_limit):
        self.obs, self.state, self.avail, self.noise = obs, state, avail, noise
        self.length, self.episode_limit = int(length), int(episode_limit)
        self.n_teams, _, self.n_agents, _ = obs.shape
        self.t = 0

    def reset(self):
        self.t = 0

    def _t(self):
        return min(self.t, self.episode_limit)

    # the reference's environments return per-agent lists over ALL teams (home agents first)
    def get_state(self, team=0):
        return self.state[team, self._t()]

    def get_avail_actions(self):
        return np.concatenate([self.avail[k, self._t()] for k in range(self.n_teams)], axis=0)

    def get_obs(self):
        return np.concatenate([self.obs[k, self._t()] for k in range(self.n_teams)], axis=0)

    def step(self, all_actions):
        """all_actions: [teams * N] ints.  Returns (obs, reward per team, done_n per team, env_info)."""
        t = self.t
        acts = np.asarray(all_actions).reshape(self.n_teams, self.n_agents)
        reward = [np.float32(self.noise[k, t]) + np.float32(0.01) * np.float32(acts[k].sum()) for k in range(self.n_teams)]
        done = (t + 1) >= self.length
        self.t = t + 1
        return self.get_obs(), reward, [done] * self.n_teams, {}


def _new_batch(TT, N, A, OBS, S):
    """EpisodeBatch(batch_size=1) as zero-initialised arrays (episode_batch.py:89-143), batch dim dropped."""
    return {"state": np.zeros((TT, S), np.float32), "obs": np.zeros((TT, N, OBS), np.float32),
            "avail_actions": np.zeros((TT, N, A), np.int32), "actions": np.zeros((TT, N, 1), np.int64),
            "actions_onehot": np.zeros((TT, N, A), np.float32), "reward": np.zeros((TT, 1), np.float32),
            "terminated": np.zeros((TT, 1), np.uint8), "filled": np.zeros((TT, 1), np.int64)}


def _update(batch, data, t):
    """EpisodeBatch.update(data, ts=t): write, mark filled, re-derive actions_onehot (episode_batch.py:157-195)."""
    for k, v in data.items():
        batch[k][t] = np.asarray(v).reshape(batch[k][t].shape)
        if k == "actions":
            batch["actions_onehot"][t] = O.onehot(batch["actions"][t], batch["actions_onehot"].shape[-1])
    batch["filled"][t] = 1


class OracleMAC:
    """BasicMAC on numpy (basic_controller.py:29-60): shared DRQN agent + epsilon-greedy selector with injected draws.
    `draws(t) -> (u [N], e [N, A])`; epsilon follows the linear schedule (epsilon_schedules.py:21-25), 0 in test mode."""

    def __init__(self, params, n_agents, n_actions, draws, eps_start=1.0, eps_finish=0.05, eps_anneal=50000, dtype=np.float32):
        self.p = {k: np.asarray(v, dtype) for k, v in params.items()}
        self.N, self.A, self.draws, self.dtype = n_agents, n_actions, draws, dtype
        self.sched = (eps_start, eps_finish, eps_anneal)
        self.h = None
        self.q_log = []

    def init_hidden(self):
        self.h = np.zeros((self.N, O.H_DEFAULT), self.dtype)

    def select_actions(self, batch, t_ep, t_env, test_mode):
        inp = O.build_inputs(batch["obs"][None].astype(self.dtype), batch["actions_onehot"][None].astype(self.dtype), t_ep)
        q, self.h, _ = O.drqn_step(self.p, inp, self.h)
        eps = 0.0 if test_mode else O.epsilon_linear(*self.sched, t_env)          # action_selectors.py:48-52
        u, e = self.draws(t_ep)
        self.q_log.append(q.copy())
        a, g = O.eps_greedy_select(q.reshape(1, self.N, self.A), batch["avail_actions"][t_ep][None], eps,
                                   np.asarray(u, np.float32).reshape(1, self.N), np.asarray(e, np.float32).reshape(self.N, self.A))
        return a[0], g[0]


def run_episode(env: PlaybackEnv, macs, dims, t_env=0, test_mode=False):
    """EpisodeStepper.run (len(macs) == 1) / SelfPlayStepper.run (len(macs) == 2) for ONE match.
    `dims` = (N, A, OBS, S).  Returns (batches per team, returns per team, steps)."""
    N, A, OBS, S = dims
    TT = env.episode_limit + 1
    n_teams = len(macs)
    batches = [_new_batch(TT, N, A, OBS, S) for _ in range(n_teams)]           # reset(): new_batch_fn(), env.reset()
    env.reset()
    t = 0
    for m in macs:
        m.init_hidden()                                                        # episode_stepper.py:103
    returns = [np.float32(0.0)] * n_teams
    terminated = False

    def pre_transition():
        avail, obs = env.get_avail_actions(), env.get_obs()                    # stepper_utils.py:4-25: split by team
        for k in range(n_teams):
            _update(batches[k], {"state": env.get_state(k), "avail_actions": avail[k * N:(k + 1) * N],
                                 "obs": obs[k * N:(k + 1) * N]}, t)

    while not terminated:                                                      # episode_stepper.py:110-142
        pre_transition()
        acts = [m.select_actions(batches[k], t, t_env, test_mode)[0] for k, m in enumerate(macs)]
        _, reward, done_n, _ = env.step(np.concatenate(acts))
        terminated = any(done_n)
        for k in range(n_teams):
            returns[k] = np.float32(returns[k] + reward[k])
            _update(batches[k], {"actions": acts[k], "reward": reward[k], "terminated": terminated}, t)
        t += 1
    pre_transition()                                                           # :144-156: last stored state + its actions
    for k, m in enumerate(macs):
        a, _ = m.select_actions(batches[k], t, t_env, test_mode)
        _update(batches[k], {"actions": a}, t)
    return batches, returns, t
