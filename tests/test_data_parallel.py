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


# ---------------------------------------------------------------------------------------------- fused exchange kernel
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("clip", [10.0, 0.05])
def test_peer_allreduce_kernel_one_device_fake_peers(world, clip):
    """`mal_peer_allreduce_clip_rmsprop` (k_peer_allreduce_grad + k_clip_rmsprop) takes raw peer pointers, so W "ranks"
    can live on ONE device: every fake rank's buffer holds the oracle's UN-normalised gradient sum of its batch shard
    followed by its raw statistic sums.  Checked against the numpy oracle on the full batch: rank-order sums, the global
    mask-sum normaliser, grad norm, clipped gradient, post-step parameters and square_avg (q_learner.py:98-105)."""
    import ctypes as C
    from ma_league_b200 import _native as nat
    from ma_league_b200.learners.q_learner import DP_TAIL
    O, params, batch, N = _oracle_case(seed=11, B=8, TT=9, N=3)
    B = batch["obs"].shape[0]
    full = O.learner_forward_backward(*params, batch, mixer="qmix", double_q=True, gamma=0.99, dtype=np.float64)
    n_agent = sum(v.size for v in params[0].values())
    n_mixer = sum(v.size for v in params[2].values())
    P = n_agent + n_mixer
    dev = "cuda:0"
    bufs, raw_sum = [], np.zeros(DP_TAIL)
    for r in range(world):
        lo, hi = r * B // world, (r + 1) * B // world
        res = O.learner_forward_backward(*params, {k: v[lo:hi] for k, v in batch.items()}, mixer="qmix",
                                         double_q=True, gamma=0.99, dtype=np.float64)
        raw = _raw(res, N)
        raw_sum += raw
        bufs.append(th.from_numpy(np.concatenate([_flat(res) * raw[4], raw]).astype(np.float32)).to(dev))
    flat0 = np.concatenate([v.ravel() for v in list(params[0].values()) + list(params[2].values())]).astype(np.float32)
    agent = th.from_numpy(flat0[:n_agent].copy()).to(dev)
    mixer = th.from_numpy(flat0[n_agent:].copy()).to(dev)
    rng = np.random.default_rng(5)
    sq0 = rng.uniform(0.0, 1e-3, P).astype(np.float32)
    sq = th.from_numpy(sq0.copy()).to(dev)
    grad = th.full((P,), 7.0, device=dev)
    tail = th.zeros(DP_TAIL, device=dev)
    scalars = th.zeros(64, device=dev)
    scratch = th.zeros((P + 255) // 256, device=dev)
    ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    lr, alpha, eps = 5e-4, 0.99, 1e-5
    nat.check(nat.lib().mal_peer_allreduce_clip_rmsprop(ptrs, world, nat.ptr(agent), n_agent, nat.ptr(mixer), n_mixer,
                                                        nat.ptr(grad), nat.ptr(tail), DP_TAIL, nat.ptr(sq), lr, alpha,
                                                        eps, clip, nat.ptr(scalars), nat.ptr(scratch), 0,
                                                        nat.current_stream(dev)), "mal_peer_allreduce_clip_rmsprop")
    th.cuda.synchronize()
    g_ref = _flat(full)                                                  # full-batch gradient (normalised)
    norm_ref, clipped = O.clip_grad_norm([g_ref], clip)
    assert (norm_ref > clip) == (clip < 1.0)                             # both branches of the clip are exercised
    assert_close(tail.cpu().numpy()[:6], raw_sum[:6], 1e-6, "reduced statistic sums")
    assert abs(float(scalars[nat.SC_GRAD_NORM]) - norm_ref) <= 1e-5 * norm_ref
    assert_close(grad.cpu().numpy(), clipped[0], 1e-5, "normalised + clipped gradient")
    p_ref, sq_ref = O.rmsprop_update(flat0.astype(np.float64), clipped[0], sq0.astype(np.float64), lr, alpha, eps)
    assert_close(np.concatenate([agent.cpu().numpy(), mixer.cpu().numpy()]), p_ref, 1e-6, "post-step parameters")
    assert_close(np.concatenate([agent.cpu().numpy(), mixer.cpu().numpy()]) - flat0, p_ref - flat0, 2e-4, "update")
    assert_close(sq.cpu().numpy(), sq_ref, 1e-5, "square_avg")
    # rank-order summation in fp32: bit-identical to the same loop on the host
    acc = np.zeros(P, np.float32)
    den = np.float32(0)
    for b in bufs:
        hb = b.cpu().numpy()
        acc = acc + hb[:P]
        den = den + hb[P + 4]
    coef = np.float32(min(np.float32(clip) / (np.float32(scalars[nat.SC_GRAD_NORM].item()) + np.float32(1e-6)), 1.0))
    assert np.array_equal(grad.cpu().numpy(), (acc / den) * coef)
    # frozen prefix (freeze_agent_weights): agent slice zeroed, left out of the norm, parameters and square_avg untouched
    agent2, mixer2 = th.from_numpy(flat0[:n_agent].copy()).to(dev), th.from_numpy(flat0[n_agent:].copy()).to(dev)
    sq2 = th.from_numpy(sq0.copy()).to(dev)
    nat.check(nat.lib().mal_peer_allreduce_clip_rmsprop(ptrs, world, nat.ptr(agent2), n_agent, nat.ptr(mixer2), n_mixer,
                                                        nat.ptr(grad), nat.ptr(tail), DP_TAIL, nat.ptr(sq2), lr, alpha,
                                                        eps, clip, nat.ptr(scalars), nat.ptr(scratch), n_agent,
                                                        nat.current_stream(dev)), "mal_peer_allreduce_clip_rmsprop")
    th.cuda.synchronize()
    norm_m, clipped_m = O.clip_grad_norm([g_ref[n_agent:]], clip)
    assert abs(float(scalars[nat.SC_GRAD_NORM]) - norm_m) <= 1e-5 * norm_m
    assert np.array_equal(agent2.cpu().numpy(), flat0[:n_agent]) and np.array_equal(sq2.cpu().numpy()[:n_agent], sq0[:n_agent])
    assert not grad[:n_agent].any()
    assert_close(grad.cpu().numpy()[n_agent:], clipped_m[0], 1e-5, "mixer gradient with a frozen agent")
    pm_ref, _ = O.rmsprop_update(flat0[n_agent:].astype(np.float64), clipped_m[0], sq0[n_agent:].astype(np.float64), lr, alpha, eps)
    assert_close(mixer2.cpu().numpy(), pm_ref, 1e-6, "mixer parameters with a frozen agent")
