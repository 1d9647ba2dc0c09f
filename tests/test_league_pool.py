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
