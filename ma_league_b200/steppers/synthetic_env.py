"""A device-resident, vectorised stand-in for the reference's `maenv` environments (an un-vendored dependency that is
not installable here): `n_envs` independent matches advance in lock-step, every tensor lives on
the device.  Observations / states / available actions are a seeded playback (so a match is reproducible whatever
batch it is part of), each match ends after its own pre-drawn number of steps, and the reward depends on the actions
taken so that tests can see the action routing.

Protocol expected by BatchedEpisodeStepper (what an adapter around a real vectorised env has to provide):
    n_envs, n_teams, episode_limit
    reset()
    observe(team) -> {"state": [B,S] f32, "avail_actions": [B,N,A] i32, "obs": [B,N,OBS] f32}     (episode_stepper.py:189-199)
    step(actions: list of [B,N] i64 per team) -> (rewards: list of [B] f32 per team, done: [B] bool, info: dict)
"""
import torch as th


class SyntheticVecEnv:
    def __init__(self, n_envs, n_agents, n_actions, obs_dim, state_dim, episode_limit, n_teams=1, seed=0, device="cuda",
                 min_len=None, env_ids=None):
        self.n_envs, self.n_agents, self.n_actions = n_envs, n_agents, n_actions
        self.obs_dim, self.state_dim, self.episode_limit, self.n_teams = obs_dim, state_dim, episode_limit, n_teams
        self.device = th.device(device)
        ids = list(range(n_envs)) if env_ids is None else list(env_ids)   # match identity: fixes its data and length
        assert len(ids) == n_envs
        TT = episode_limit + 1
        lo = max(1, episode_limit // 2) if min_len is None else min_len
        obs, state, avail, noise, lens = [], [], [], [], []
        for i in ids:
            g = th.Generator().manual_seed(1_000_003 * (seed + 1) + i)
            obs.append(th.randn(n_teams, TT, n_agents, obs_dim, generator=g))
            state.append(th.randn(n_teams, TT, state_dim, generator=g))
            a = (th.rand(n_teams, TT, n_agents, n_actions, generator=g) < 0.7).int()
            a[..., 0] = 1
            avail.append(a)
            noise.append(th.randn(n_teams, TT, generator=g))
            lens.append(int(th.randint(lo, episode_limit + 1, (1,), generator=g)))
        dev = self.device
        self._obs = th.stack(obs, 1).to(dev)        # [teams, B, TT, N, OBS]
        self._state = th.stack(state, 1).to(dev)
        self._avail = th.stack(avail, 1).to(dev)
        self._noise = th.stack(noise, 1).to(dev)    # [teams, B, TT]
        self.lengths = th.tensor(lens, device=dev)  # match b ends with its step lengths[b]-1
        self.t = 0

    def reset(self):
        self.t = 0

    def observe(self, team=0):
        t = min(self.t, self.episode_limit)
        return {"state": self._state[team, :, t], "avail_actions": self._avail[team, :, t], "obs": self._obs[team, :, t]}

    def step(self, actions):
        t = self.t
        rewards = [self._noise[k, :, t] + 0.01 * actions[k].sum(dim=1).float() for k in range(self.n_teams)]
        done = (t + 1) >= self.lengths
        self.t = t + 1
        return rewards, done, {}

    def get_env_info(self):
        return {"n_agents": self.n_agents, "n_actions": self.n_actions, "obs_shape": self.obs_dim,
                "state_shape": self.state_dim, "episode_limit": self.episode_limit}

    def close(self):
        pass
