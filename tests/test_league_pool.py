"""Device agent pool: world-size-2 gloo run of the parameter exchange on CPU modules, and the
state_dict contract the reference's league commands rely on."""
import os
import socket

import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

KEYS = ["fc1.weight", "fc1.bias", "gru.weight_ih", "gru.weight_hh", "gru.bias_ih", "gru.bias_hh", "fc2.weight", "fc2.bias"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mac(seed):
    import ma_league_b200 as M
    from ma_league_b200.synthetic import make_args, make_scheme
    th.manual_seed(seed)
    N, A, OBS, S = 3, 9, 32, 48
    args = make_args(N, A, S, mixer="vdn", device="cpu")
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    scheme = dict(scheme, actions_onehot={"vshape": (A,), "group": "agents"})
    return M.mac_REGISTRY["basic"](scheme, groups, args)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ma_league_b200.league import DeviceAgentPool
        home, away = _mac(100 + rank), _mac(999)
        pool = DeviceAgentPool(home)
        pool.sync()
        other = 1 - rank
        pool.load_into(away, other)                        # play against the other instance's agent
        sd = pool.state_dict(other)
        out.put((rank, {k: v.numpy().copy() for k, v in home.agent.state_dict().items()},
                 {k: v.numpy().copy() for k, v in away.agent.state_dict().items()}, list(sd.keys()),
                 all(th.equal(sd[k], away.agent.state_dict()[k]) for k in sd)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_pool_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(), out.get()], key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, home, away, keys, same in res:
        assert keys == KEYS and same
        peer_home = res[1 - rank][1]
        for k in KEYS:
            assert (away[k] == peer_home[k]).all(), k        # the away controller now holds the peer's home agent
        assert not (home["fc1.weight"] == peer_home["fc1.weight"]).all()


def test_single_process_pool_roundtrip():
    from ma_league_b200.league import DeviceAgentPool
    home, away = _mac(1), _mac(2)
    pool = DeviceAgentPool(home)
    pool.sync()
    away.load_state_dict(agent=pool.state_dict(0))         # the reference's path: load_state_dict(agent=OrderedDict)
    for k, v in home.agent.state_dict().items():
        assert th.equal(v, away.agent.state_dict()[k])
    with th.no_grad():
        next(home.agent.parameters()).add_(1.0)
    pool.sync()
    pool.load_into(away, 0)
    assert th.equal(next(home.agent.parameters()), next(away.agent.parameters()))


# ---------------------------------------------------------------------------------------------- GPU
import numpy as np   # noqa: E402
import pytest        # noqa: E402


def _gpu_mac_and_batch(seed, dev):
    import ma_league_b200 as M
    from ma_league_b200.synthetic import make_args, make_scheme, synth_episode_data, fill_episode_batch
    th.manual_seed(seed)
    N, A, OBS, S, B, TT = 3, 9, 32, 48, 4, 5
    args = make_args(N, A, S, mixer="vdn", device=dev)
    scheme, groups, pre = make_scheme(N, A, OBS, S)
    buf = M.ReplayBuffer(scheme, groups, B, TT, preprocess=pre, device=dev)
    mac = M.mac_REGISTRY["basic"](buf.scheme, groups, args)
    gen = th.Generator().manual_seed(42)                     # the same evaluation batch everywhere
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=False, device=dev)
    eb = fill_episode_batch(M.EpisodeBatch(scheme, groups, B, TT, preprocess=pre, device=dev), data, lens)
    return mac, eb


def _play(mac, eb):
    """Greedy actions + Q-values of two steps through the fused act-select kernel."""
    mac.init_hidden(eb.batch_size)
    out = []
    for t in range(2):
        a, _ = mac.select_actions(eb, t_ep=t, t_env=0, test_mode=True)
        out.append(a.cpu().numpy())
    mac.init_hidden(eb.batch_size)
    q = mac.forward(eb, 0).cpu().numpy()
    return out, q


@pytest.mark.gpu
def test_device_pool_on_cuda_single_process():
    """SURVEY.md 8(f2) on the device: the pool tensor, the state_dict views and the loaded controller all live on cuda:0;
    a controller loaded from the pool selects exactly the actions of the pooled agent (the pool feeds the act-select kernel)."""
    from ma_league_b200.league import DeviceAgentPool
    home, eb = _gpu_mac_and_batch(1, "cuda:0")
    away, _ = _gpu_mac_and_batch(2, "cuda:0")
    pool = DeviceAgentPool(home)
    pool.sync()
    assert pool.pool.is_cuda and pool.pool.shape == (1, sum(p.numel() for p in home.parameters()))
    sd = pool.state_dict(0)
    assert list(sd.keys()) == KEYS and all(v.is_cuda for v in sd.values())
    assert sd["fc1.weight"].data_ptr() == pool.pool.data_ptr()            # views of the pool tensor, not copies
    a_before, q_before = _play(away, eb)
    pool.load_into(away, 0)
    a_home, q_home = _play(home, eb)
    a_away, q_away = _play(away, eb)
    assert np.array_equal(q_home, q_away) and all(np.array_equal(x, y) for x, y in zip(a_home, a_away))
    assert not np.array_equal(q_before, q_home)
    away2, _ = _gpu_mac_and_batch(3, "cuda:0")
    away2.load_state_dict(agent=pool.state_dict(0))                        # the reference's path (sp_ma_experiment.py:27-29)
    assert np.array_equal(_play(away2, eb)[1], q_home)
    with th.no_grad():
        next(home.agent.parameters()).mul_(1.5)
    pool.sync()
    pool.load_into(away, 0)
    assert np.array_equal(_play(away, eb)[1], _play(home, eb)[1]) and pool.n_syncs == 2


def _gpu_worker(rank, world, port, out, nccl):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = "cuda:%d" % (rank if nccl else 0)
    th.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=th.device(dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ma_league_b200.league import DeviceAgentPool
        home, eb = _gpu_mac_and_batch(100 + rank, dev)
        away, _ = _gpu_mac_and_batch(999, dev)
        pool = DeviceAgentPool(home)
        pool.sync()
        pool.load_into(away, 1 - rank)
        out.put((rank, _play(home, eb), _play(away, eb), bool(pool.pool.is_cuda)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("nccl", [False, True])
def test_device_pool_two_ranks_on_gpu(nccl):
    """Two league instances exchange their home agents on the device (gloo with both ranks on cuda:0; NCCL over NVLink
    when two GPUs exist): each rank's away controller then plays exactly like the peer's home controller."""
    if nccl and th.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (all-gather over NCCL / NVLink)")
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, out, nccl)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(), out.get()], key=lambda r: r[0])
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for rank, home_play, away_play, on_cuda in res:
        assert on_cuda
        peer_home = res[1 - rank][1]
        assert np.array_equal(away_play[1], peer_home[1])                      # Q-values of the peer's agent, bit for bit
        assert all(np.array_equal(x, y) for x, y in zip(away_play[0], peer_home[0]))
        assert not np.array_equal(home_play[1], peer_home[1])
