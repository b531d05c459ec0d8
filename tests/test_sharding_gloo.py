"""CPU (-m "not gpu"): the N>1 path with world_size 2 over gloo.  Each rank steps its shard of envs (with the CPU
oracle standing in for the device, since the sharding logic is what is under test), results are all-gathered and
must equal a single-process run over all envs; the timing reduction must return the slowest rank's value."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(N):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
    import scenarios as S
    from gym_fx_b200.config import lower_config
    from gym_fx_b200.synth import synth_candles

    cfgd = {**S.DEFAULTS, "window_size": 16, "feature_columns": list(S.OHLCV), "feature_scaling_window": 32}
    pl = S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, "strategy": "direct_atr_sltp", "reward": "sharpe_reward",
                                       "preprocessor": "feature_window_preprocessor"})
    cfg = lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"],
                       preprocessor_plugin=pl["preprocessor"], reward_plugin=pl["reward"], columns=S.OHLCV,
                       num_envs=N, num_pairs=2)
    return cfg, [synth_candles(2048, p) for p in range(2)]


def _run(cfg, candles, starts, acts):
    from oracle.c_oracle import OracleVec

    env = OracleVec(cfg, candles)
    env.reset(starts)
    out = []
    for a in acts:
        obs, rew, rew64, term = env.step(a)
        out.append(np.concatenate([obs.astype(np.float64), rew64[:, None], term[:, None].astype(np.float64)], axis=1))
    inf = env.info()
    return np.stack(out), inf["equity"]


def _worker(rank, world, port, per_rank, steps, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
    from gym_fx_b200.sharding import check_pair_alignment, max_over_ranks, shard_range, shard_starts

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, candles = _case(per_rank)
    check_pair_alignment(per_rank, cfg.num_pairs)
    r = shard_range(per_rank, rank, world)
    starts = shard_starts(per_rank, rank, world, T, steps, 32)
    acts_all = np.random.default_rng(99).integers(0, 3, (steps, per_rank * world)).astype(np.int32)
    traj, eq = _run(cfg, candles, starts, acts_all[:, r.start:r.stop])
    dist.barrier()
    gathered = [torch.zeros_like(torch.from_numpy(traj)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(traj))
    slow = max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        q.put((torch.cat(gathered, dim=1).numpy(), slow))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_env_sharding_matches_single_process():
    world, per_rank, steps, T = 2, 6, 60, 2048
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, steps, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    sharded, slow = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    sys.path[:0] = [ROOT]
    from gym_fx_b200.synth import start_offsets

    cfg, candles = _case(per_rank * world)
    acts_all = np.random.default_rng(99).integers(0, 3, (steps, per_rank * world)).astype(np.int32)
    full, _ = _run(cfg, candles, start_offsets(per_rank * world, T, steps, 32), acts_all)
    assert sharded.shape == full.shape
    assert np.array_equal(sharded, full), "env shards across 2 ranks differ from the single-process run"
    assert slow == 11.0  # max over ranks


def test_shard_helpers():
    from gym_fx_b200.sharding import check_pair_alignment, shard_range, shard_starts
    from gym_fx_b200.synth import start_offsets

    assert list(shard_range(4, 1, 2)) == [4, 5, 6, 7]
    allr = np.concatenate([shard_starts(8, r, 4, 1 << 12, 100, 32) for r in range(4)])
    assert np.array_equal(allr, start_offsets(32, 1 << 12, 100, 32))
    with pytest.raises(ValueError):
        check_pair_alignment(6, 4)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _learner_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT]
    from gym_fx_b200.sharding import allreduce_mean_grads, global_mean_std

    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 6, generator=g)
    adv_all = torch.randn(40, generator=g) * 3 + 1
    x = x_all[rank * 4:(rank + 1) * 4]
    net(x).pow(2).mean().backward()                 # local minibatch = this rank's shard
    allreduce_mean_grads(list(net.parameters()), dist)
    grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    m, s = global_mean_std(adv_all[rank * 20:(rank + 1) * 20], dist)
    q.put((rank, grads.numpy(), float(m), float(s)))
    dist.barrier()
    dist.destroy_process_group()


def test_learner_collectives_world2_match_single_process():
    """allreduce_mean_grads / global_mean_std over two gloo ranks == the single-process result over the whole batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 6, generator=g)
    adv_all = torch.randn(40, generator=g) * 3 + 1
    net(x_all).pow(2).mean().backward()             # equal shard sizes: mean of shard means == global mean
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    for _, grads, m, s in res:
        np.testing.assert_allclose(grads, ref, rtol=1e-5, atol=1e-7)
        assert abs(m - float(adv_all.mean())) < 1e-5 and abs(s - float(adv_all.std(unbiased=False))) < 1e-5
