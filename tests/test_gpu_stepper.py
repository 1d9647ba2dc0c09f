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
