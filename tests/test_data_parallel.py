"""Data-parallel learner mode: world_size-2 tests.

CPU (gloo): the reduction contract -- ranks hold UN-normalised gradient sums of their batch shard plus raw statistic
sums, one all-reduce(sum), divide by the global mask sum -- reproduces the full-batch gradient / statistics of
q_learner.py:98-124 (checked with the numpy oracle, which is test infrastructure).
GPU: two processes on cuda:0 (gloo moves the CUDA tensors) run QLearner.train(data_parallel=True) on half batches and
must land on the parameters of one full-batch step.
"""
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_close


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_case(seed=3, B=4, TT=7, N=3):
    from oracle import np_oracle as O
    from ma_league_b200.synthetic import synth_episode_data
    A, OBS, S = 6 + N, 8 + 8 * N, 16 * N
    rng = np.random.default_rng(seed)
    ap = O.init_params(O.agent_param_shapes(OBS + A + N, A), rng)
    tp = O.init_params(O.agent_param_shapes(OBS + A + N, A), rng)
    mp_ = O.init_params(O.qmix_param_shapes(S, N), rng)
    tmp = O.init_params(O.qmix_param_shapes(S, N), rng)
    gen = th.Generator().manual_seed(seed)
    data, lens = synth_episode_data(B, TT, N, A, OBS, S, gen, var_len=True)
    batch = {k: v.numpy() for k, v in data.items()}
    batch["actions_onehot"] = O.onehot(batch["actions"], A)
    filled = np.zeros((B, TT, 1), np.int64)
    for b, l in enumerate(lens):
        filled[b, :int(l) + 1] = 1
    batch["filled"] = filled
    return O, (ap, tp, mp_, tmp), batch, N


def _flat(res):
    return np.concatenate([v.ravel() for v in list(res["agent_grads"].values()) + list(res["mixer_grads"].values())])


def _raw(res, n_agents):
    st = res["stats"]
    ms = float(st["mask_sum"])
    return np.array([st["loss"] * ms, st["td_error_abs"] * ms, st["q_taken_mean"] * ms * n_agents,
                     st["target_mean"] * ms * n_agents, ms, st["trained_steps"], 0, 0], np.float64)


def _cpu_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O, params, batch, N = _oracle_case()
        B = batch["obs"].shape[0]
        lo, hi = rank * B // world, (rank + 1) * B // world
        shard = {k: v[lo:hi] for k, v in batch.items()}
        res = O.learner_forward_backward(*params, shard, mixer="qmix", double_q=True, gamma=0.99, dtype=np.float64)
        raw = _raw(res, N)
        buf = th.from_numpy(np.concatenate([_flat(res) * raw[4], raw]))       # un-normalised sums | raw stats
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if rank == 0:
            from ma_league_b200.learners.q_learner import global_stats, DP_TAIL
            n = buf.numel() - DP_TAIL
            out.put((buf[:n].numpy() / float(buf[n + 4]), global_stats(buf[n:].float(), N).numpy(), float(buf[n + 5])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_dp_reduction_contract_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    grad, stats, count = out.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    O, params, batch, N = _oracle_case()
    full = O.learner_forward_backward(*params, batch, mixer="qmix", double_q=True, gamma=0.99, dtype=np.float64)
    assert_close(grad, _flat(full), 1e-12, "all-reduced gradient")
    st = full["stats"]
    from ma_league_b200 import _native as nat
    assert abs(stats[nat.SC_LOSS] - st["loss"]) <= 1e-6 * abs(st["loss"])
    assert abs(stats[nat.SC_TD_ABS] - st["td_error_abs"]) <= 1e-6 * abs(st["td_error_abs"])
    assert abs(stats[nat.SC_Q_TAKEN] - st["q_taken_mean"]) <= 1e-6 * max(1, abs(st["q_taken_mean"]))
    assert abs(stats[nat.SC_TARGET] - st["target_mean"]) <= 1e-6 * max(1, abs(st["target_mean"]))
    assert int(count) == st["trained_steps"]


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_worker(rank, world, port, out, nccl, fused=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = "cuda:%d" % (rank if nccl else 0)      # gloo: both ranks share cuda:0; nccl: one GPU per rank
    th.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=th.device(dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.gpu_helpers import seeded_system, np_params
        s = seeded_system(3, 8, 12, "qmix", True, seed=5, device=dev, data_parallel=True, learner_log_interval=0,
                          dp_fused=fused)
        B = s.batch.batch_size
        lo, hi = rank * B // world, (rank + 1) * B // world
        for i in range(3):
            s.learner.train(s.batch[lo:hi], t_env=i, episode_num=i)
        th.cuda.synchronize()
        out.put((rank, np_params(s.mac.agent), np_params(s.learner.mixer),
                 {k: v[0] for k, v in s.logger.stats.items()}, s.mac.agent.trained_steps,
                 bool(getattr(s.learner, "_dp_sym", None)), getattr(s.learner, "_dp_sym_error", "")))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("nccl,fused", [(False, False), (True, False), (True, True)])
def test_dp_train_two_ranks_equals_full_batch(nccl, fused):
    """gloo on one GPU and NCCL on two GPUs use the all-reduce path; (True, True) is the fused exchange: the peer-memory
    all-reduce kernel reads both ranks' symmetric gradient buffers over NVLink inside the optimiser prologue."""
    from tests.gpu_helpers import seeded_system, np_params
    if nccl and th.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gradient all-reduce over NCCL / NVLink)")
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, out, nccl, fused)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(), out.get()], key=lambda r: r[0])
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    s = seeded_system(3, 8, 12, "qmix", True, seed=5, device="cuda:0", learner_log_interval=0)
    for i in range(3):
        s.learner.train(s.batch, t_env=i, episode_num=i)
    ref_a, ref_m = np_params(s.mac.agent), np_params(s.learner.mixer)
    for _, pa, pm, stats, steps, was_fused, err in res:
        assert was_fused == fused, err
        for k in ref_a:
            assert_close(pa[k], ref_a[k], 1e-5, "dp agent " + k)
        for k in ref_m:
            assert_close(pm[k], ref_m[k], 1e-5, "dp mixer " + k)
        for k, v in s.logger.stats.items():
            assert abs(stats[k] - v[0]) <= 1e-5 * max(1.0, abs(v[0])), (k, stats[k], v[0])
        assert steps == s.mac.agent.trained_steps
    # replicas stay bit-identical: same reduced gradient, same update on every rank
    for k in ref_a:
        assert np.array_equal(res[0][1][k], res[1][1][k]), k
