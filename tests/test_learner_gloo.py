"""CPU (-m "not gpu"): learner-side glue of the closed loop (gym_fx_b200/learner.py) -- GAE with auto-reset masking,
and the two collectives of a sharded PPO update over gloo with world_size 2: the 3-float advantage-statistics all-reduce
and the flat-bucket gradient all-reduce (replicas must stay bit-identical)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gae_masks_reset_steps_and_cuts_at_done():
    from gym_fx_b200.learner import gae

    H, N = 5, 2
    rew = torch.tensor([[1.0, 0.0], [0.0, 0.0], [2.0, 0.0], [0.0, 1.0], [1.0, 0.0]])
    val = torch.zeros(H + 1, N)
    val[:, 1] = 0.5
    done = torch.zeros(H, N, dtype=torch.uint8)
    done[1, 0] = 1                      # env 0 terminates at step 1 -> step 2 is its reset step
    adv, ret, valid = gae(rew, val, done, torch.tensor([0, 1], dtype=torch.uint8), gamma=0.5, lam=1.0)
    assert valid[:, 0].tolist() == [1, 1, 0, 1, 1] and valid[:, 1].tolist() == [0, 1, 1, 1, 1]   # env 1 starts on a reset step
    # env 0: the recursion is cut at the terminal step: adv[1] = r1 - v1 = 0, adv[0] = r0 + 0.5 * adv[1] = 1
    assert adv[1, 0].item() == 0.0 and adv[0, 0].item() == 1.0
    # after the cut the tail is an ordinary discounted sum: adv[2] = 2 + 0.5 * (0 + 0.5 * 1) = 2.25
    assert abs(adv[2, 0].item() - 2.25) < 1e-6
    assert torch.allclose(ret, adv + val[:H])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from gym_fx_b200.learner import ActorCritic, masked_global_mean_std, ppo_update

    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                      # identical replicas
    H, N, D = 6, 8, 20
    net = ActorCritic(D, hidden=16)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(100 + rank)   # every rank has its own shard of the rollout
    buf = {"obs": torch.randn(H + 1, N, D, generator=g), "actions": torch.randint(0, 3, (H, N), generator=g).int(),
           "logp": -torch.rand(H, N, generator=g) - 0.5, "value": torch.randn(H + 1, N, generator=g),
           "reward": torch.randn(H, N, generator=g), "done": (torch.rand(H, N, generator=g) < 0.15).to(torch.uint8)}
    x, mask = buf["reward"], (torch.rand(H, N, generator=g) < 0.8).float()
    m, s = masked_global_mean_std(x, mask, dist)
    stats = ppo_update(net, opt, buf, torch.zeros(N, dtype=torch.uint8), dist, epochs=2, minibatches=3)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    q.put((rank, float(m), float(s), x.numpy(), mask.numpy(), flat.numpy(), stats))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ppo_update_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.concatenate([r[3].reshape(-1) for r in res]); mk = np.concatenate([r[4].reshape(-1) for r in res]).astype(bool)
    for r in res:   # every rank got the statistics of the UNION of the valid elements
        assert abs(r[1] - x[mk].mean()) < 1e-6 and abs(r[2] - x[mk].std()) < 1e-5
    assert np.array_equal(res[0][5], res[1][5]), "replicas diverged: the gradient all-reduce did not average"
    torch.manual_seed(0)
    sys.path.insert(0, ROOT)
    from gym_fx_b200.learner import ActorCritic
    init = torch.cat([p.detach().reshape(-1) for p in ActorCritic(20, hidden=16).parameters()]).numpy()
    assert not np.allclose(res[0][5], init), "the update did not move the parameters"
    assert 0.0 < res[0][6]["valid_frac"] <= 1.0
