"""Batched device-resident rollout loop against the reference's single-match loop semantics.

The reference's EpisodeStepper / SelfPlayStepper only run one match at a time (`assert self.batch_size == 1`,
steppers/episode_stepper.py:26).  A B-match lock-step run must therefore leave, row by row, exactly the batch that B
separate single-match runs of the same loop leave -- including matches of different lengths -- and a single-match run
is checked line by line against the reference loop's bookkeeping (filled / terminated / final-state actions / t_env)."""
import numpy as np
import pytest
import torch as th

import ma_league_b200 as M
from ma_league_b200.steppers import BatchedEpisodeStepper, SyntheticVecEnv
from ma_league_b200.synthetic import make_args, make_scheme

pytestmark = pytest.mark.gpu
DEV = "cuda"
N, A, OBS, S, LIMIT = 3, 9, 32, 48, 12


def _system(seed=0):
    th.manual_seed(seed)
    args = make_args(N, A, S, mixer="vdn", device=DEV, batch_size_run=1)
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, 8, LIMIT + 1, preprocess=pre, device=DEV)
    home = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    away = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    return args, scheme, groups, pre, home, away


def _run(env_ids, n_teams, args, scheme, groups, pre, home, away, test_mode=True, sync_every=1):
    env = SyntheticVecEnv(len(env_ids), N, A, OBS, S, LIMIT, n_teams=n_teams, seed=7, device=DEV, env_ids=env_ids, min_len=3)
    st = BatchedEpisodeStepper(args, None, env, sync_every=sync_every)
    st.initialize(scheme, groups, pre, home, away if n_teams == 2 else None)
    out = st.run(test_mode=test_mode)
    return st, env, out


@pytest.mark.parametrize("n_teams", [1, 2])
def test_lockstep_batch_equals_single_match_runs(n_teams):
    sysm = _system()
    ids = [0, 1, 2, 3, 4]
    st, env, out = _run(ids, n_teams, *sysm)
    batches = out[:-1]
    lens = env.lengths.cpu().numpy()
    assert len(set(lens.tolist())) > 1                       # matches of different lengths
    assert st.t_env == 0                                     # test_mode: t_env does not advance
    for i in ids:
        st1, env1, out1 = _run([i], n_teams, *sysm)
        for bb, b1 in zip(batches, out1[:-1]):
            for k in bb.data.transition_data:
                assert th.equal(bb[k][i], b1[k][0]), (k, i)
    # bookkeeping of the reference loop, per match
    for bb in batches:
        filled = bb["filled"][..., 0].cpu().numpy()
        term = bb["terminated"][..., 0].cpu().numpy()
        for b, L in enumerate(lens):
            assert filled[b, :L + 1].all() and not filled[b, L + 1:].any()         # L transitions + the last stored state
            assert term[b, L - 1] == 1 and term[b].sum() == 1
            assert int(bb.max_t_filled()) == lens.max() + 1
            oh = bb["actions_onehot"][b].cpu().numpy()
            acts = bb["actions"][b, :, :, 0].cpu().numpy()
            assert (oh[:L + 1].argmax(-1) == acts[:L + 1]).all() and oh[:L + 1].sum(-1).min() == 1   # incl. final-state actions
            assert not oh[L + 1:].any() and not bb["obs"][b, L + 1:].any()        # nothing written after the end
            avail = bb["avail_actions"][b].cpu().numpy()
            assert np.take_along_axis(avail[:L + 1], acts[:L + 1, :, None], -1).all()   # only available actions chosen
    assert th.equal(out[-1]["episode_steps"].cpu(), env.lengths.cpu())


def test_training_mode_advances_t_env_and_feeds_the_buffer():
    args, scheme, groups, pre, home, away = _system(1)
    st, env, (batch, info) = _run([0, 1, 2, 3], 1, args, scheme, groups, pre, home, away, test_mode=False, sync_every=4)
    assert st.t_env == int(env.lengths.sum())
    assert st.epsilon == home.action_selector.epsilon
    buf = M.ReplayBuffer(scheme, groups, 8, LIMIT + 1, preprocess=pre, device=DEV)
    buf.insert_episode_batch(batch)                            # same record layout: one bulk copy
    assert buf.episodes_in_buffer == 4 and th.equal(buf["obs"][:4], batch["obs"])
    # rewards follow the routed actions: reward = noise + 0.01 * sum(actions)
    L = int(env.lengths[0])
    a = batch["actions"][0, :L, :, 0].sum(-1).float()
    assert th.allclose(batch["reward"][0, :L, 0], env._noise[0, 0, :L] + 0.01 * a)


@pytest.mark.parametrize("n_teams", [1, 2])
@pytest.mark.parametrize("fuse", [True, False])
def test_rollout_against_the_reference_loop_oracle(n_teams, fuse):
    """SURVEY.md 8(f1) against an ORACLE: oracle/rollout_oracle.py restates EpisodeStepper.run (steppers/episode_stepper.py:
    86-186) and SelfPlayStepper.run (self_play_stepper.py:44-147) for one match on numpy.  B lock-step matches on the
    device -- through the fused one-launch rollout step and through the separate update / select_actions calls -- must
    leave, match by match, exactly the episode batch the reference loop leaves, with exploration on (injected draws)."""
    from oracle import rollout_oracle as RO
    from tests.gpu_helpers import np_params
    args, scheme, groups, pre, home, away = _system(3)
    macs = [home, away][:n_teams]
    B, T_ENV = 5, 25000                                           # epsilon = 0.525: random and greedy picks both occur
    env = SyntheticVecEnv(B, N, A, OBS, S, LIMIT, n_teams=n_teams, seed=11, device=DEV, min_len=3)
    gen = th.Generator().manual_seed(5)
    U = [[th.rand(B, N, generator=gen) for _ in range(LIMIT + 1)] for _ in range(n_teams)]
    E = [[th.empty(B * N, A).exponential_(generator=gen) for _ in range(LIMIT + 1)] for _ in range(n_teams)]
    st = BatchedEpisodeStepper(args, None, env, sync_every=1, fuse=fuse)
    st.initialize(scheme, groups, pre, home, away if n_teams == 2 else None)
    st.t_env = T_ENV
    out = st.run(test_mode=False, draws=lambda k, t: (U[k][t], E[k][t]))
    batches, info = out[:-1], out[-1]
    lens = env.lengths.cpu().numpy()
    assert st.t_env == T_ENV + int(lens.sum()) and len(set(lens.tolist())) > 1
    n_random = 0
    for b in range(B):
        penv = RO.PlaybackEnv(env._obs[:, b].cpu().numpy(), env._state[:, b].cpu().numpy(), env._avail[:, b].cpu().numpy(),
                              env._noise[:, b].cpu().numpy(), lens[b], LIMIT)
        omacs = [RO.OracleMAC(np_params(m.agent), N, A, dtype=np.float64,
                              draws=(lambda t, k=k: (U[k][t][b].numpy(), E[k][t].view(B, N, A)[b].numpy())))
                 for k, m in enumerate(macs)]
        ref_batches, ref_returns, ref_steps = RO.run_episode(penv, omacs, (N, A, OBS, S), t_env=T_ENV, test_mode=False)
        assert ref_steps == lens[b] == int(info["episode_steps"][b])
        for k in range(n_teams):
            for key, ref in ref_batches[k].items():
                got = batches[k][key][b].cpu().numpy()
                if key == "reward":
                    assert np.allclose(got, ref, rtol=1e-6, atol=1e-6), (key, b, k)
                else:
                    assert np.array_equal(got, ref), (key, b, k)              # bit-exact: indices, masks, copied floats
            assert abs(float(info["episode_returns"][k][b]) - float(ref_returns[k])) <= 1e-5 * max(1.0, abs(float(ref_returns[k])))
            u = np.stack([U[k][t][b].numpy() for t in range(lens[b] + 1)])
            n_random += int((u < 0.525).sum())
    assert n_random > 0


def test_fused_rollout_step_is_one_launch_per_team_and_timestep():
    from ma_league_b200 import _native as nat
    args, scheme, groups, pre, home, away = _system(4)
    env = SyntheticVecEnv(4, N, A, OBS, S, LIMIT, n_teams=2, seed=2, device=DEV, min_len=LIMIT)
    st = BatchedEpisodeStepper(args, None, env, sync_every=4)
    st.initialize(scheme, groups, pre, home, away)
    home.action_selector.validate = away.action_selector.validate = False
    n0 = nat.lib().mal_launch_count()
    st.run(test_mode=True)
    assert nat.lib().mal_launch_count() - n0 == 2 * (LIMIT + 1)
