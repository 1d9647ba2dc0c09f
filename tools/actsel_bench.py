"""Act-select latency probe (not product code): kernel / API time of BasicMAC.select_actions at rollout shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import ma_league_b200 as M
from ma_league_b200 import _native as nat
from ma_league_b200.synthetic import make_args, make_scheme, synth_episode_data, fill_episode_batch

def run(N, bs, lat, validate):
    A, OBS, S, TT = 6 + N, 8 + 8 * N, 16 * N, 20
    nat.lib().mal_set_option(b"actsel_lat", lat)
    th.manual_seed(0)
    args = make_args(N, A, S, device="cuda")
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, 1, TT, preprocess=pre, device="cuda")
    mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    mac.action_selector.validate = validate
    gen = th.Generator().manual_seed(1)
    data, lens = synth_episode_data(bs, TT, N, A, OBS, S, gen, var_len=False, device="cuda")
    eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, bs, TT, preprocess=pre, device="cuda"), data, lens)
    mac.init_hidden(bs)
    for i in range(20):
        mac.select_actions(eb, t_ep=1 + i % 15, t_env=i)
    th.cuda.synchronize()
    reps = 300
    w0 = time.perf_counter()
    for i in range(reps):
        mac.select_actions(eb, t_ep=1 + i % 15, t_env=i)
    th.cuda.synchronize()
    api = (time.perf_counter() - w0) / reps * 1e6
    nat.profile_begin()
    for i in range(50):
        mac.select_actions(eb, t_ep=1 + i % 15, t_env=i)
    pk = nat.profile_end()["k_agent_step"]
    print("N=%2d bs=%4d lat=%d validate=%d: kernel %6.2f us  api %6.2f us/call" % (N, bs, lat, validate, pk[1] / pk[0] * 1e3, api), flush=True)

def run_modes(N, bs, lat):
    A, OBS, S, TT = 6 + N, 8 + 8 * N, 16 * N, 20
    nat.lib().mal_set_option(b"actsel_lat", lat)
    th.manual_seed(0)
    args = make_args(N, A, S, device="cuda")
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, 1, TT, preprocess=pre, device="cuda")
    mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    mac.action_selector.validate = False
    gen = th.Generator().manual_seed(1)
    data, lens = synth_episode_data(bs, TT, N, A, OBS, S, gen, var_len=False, device="cuda")
    eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, bs, TT, preprocess=pre, device="cuda"), data, lens)
    u, e = th.rand(bs, N, device="cuda"), th.empty(bs * N, A, device="cuda").exponential_()
    for what, fn in (("forward only", lambda i: mac.forward(eb, 1 + i % 15)),
                     ("select, injected draws", lambda i: mac.select_actions(eb, 1 + i % 15, i, u=u, e=e)),
                     ("select, philox", lambda i: mac.select_actions(eb, 1 + i % 15, i))):
        mac.init_hidden(bs)
        for i in range(10):
            fn(i)
        nat.profile_begin()
        for i in range(50):
            fn(i)
        pk = nat.profile_end()["k_agent_step"]
        print("N=%2d bs=%4d lat=%d %-24s kernel %6.2f us" % (N, bs, lat, what, pk[1] / pk[0] * 1e3), flush=True)

for lat in (0, 1):
    run_modes(5, 1, lat)
run(5, 1, 1, False)
run(5, 32, 1, False)
run(20, 1, 1, False)
