"""Host-side cost of BasicMAC.select_actions at the rollout shape (bs = 1): cProfile over 2000 calls (not product code)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import ma_league_b200 as M
from ma_league_b200.synthetic import make_args, make_scheme, synth_episode_data, fill_episode_batch

N, bs = 5, 1
A, OBS, S, TT = 11, 48, 80, 201
args = make_args(N, A, S, device="cuda")
scheme, groups, pre = make_scheme(N, A, OBS, S)
buf = M.ReplayBuffer(scheme, groups, 1, TT, preprocess=pre, device="cuda")
mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
gen = th.Generator().manual_seed(1)
data, lens = synth_episode_data(bs, TT, N, A, OBS, S, gen, var_len=False, device="cuda")
eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, bs, TT, preprocess=pre, device="cuda"), data, lens)
for validate in (False, True):
    mac.action_selector.validate = validate
    mac.init_hidden(bs)
    for i in range(50):
        mac.select_actions(eb, t_ep=1 + i % 100, t_env=i)
    th.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(2000):
        mac.select_actions(eb, t_ep=1 + i % 100, t_env=i)
    th.cuda.synchronize()
    print("validate=%s: %.1f us per call" % (validate, (time.perf_counter() - t0) / 2000 * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(2000):
        mac.select_actions(eb, t_ep=1 + i % 100, t_env=i)
    pr.disable()
    th.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
