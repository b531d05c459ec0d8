#!/usr/bin/env python3
"""
examples/ppo_rollout.py -- PPO on the device-resident env (BASELINE configs[3] shape).

One process per GPU.  Each rank owns a shard of envs (gym_fx_b200.VecFxEnv, no collective in the step path) and a
replica of an actor-critic MLP(256, 256).  A rollout of H steps runs entirely on the device: the fused tcgen05 policy
kernel and the env step kernel alternate (VecFxEnv.rollout -> fxenv_rollout, a cached CUDA graph of 2H + 2 kernels), the
observations never leave the GPU.  The PPO update is plain torch; gradients are averaged by ONE flat NCCL all-reduce per
minibatch and advantages are normalised with global statistics (gym_fx_b200.learner / gym_fx_b200.sharding).

    python examples/ppo_rollout.py --envs 4096 --horizon 32 --updates 5
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/ppo_rollout.py --envs 4096

Prints one JSON line: env-steps/s of the rollout phase (policy in the loop) and of the whole train loop.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--horizon", type=int, default=32, help="steps per rollout")
    ap.add_argument("--updates", type=int, default=5)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--workload", default="cfg4", help="bench.py workload shape (cfg4 = W=128, fixed SL/TP, sharpe)")
    ap.add_argument("--lr", type=float, default=3e-4)
    args = ap.parse_args()

    import bench
    from gym_fx_b200.learner import ActorCritic, ppo_update
    from gym_fx_b200.sharding import check_pair_alignment, shard_starts
    from gym_fx_b200.vec_env import VecFxEnv

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    cfg, candles, minutes, N, D, _, desc = bench.build_workload(args.workload, args.envs)
    env = VecFxEnv(cfg, candles, minutes, device=dev)
    check_pair_alignment(N, cfg.num_pairs)
    H = args.horizon
    env.reset(torch.as_tensor(shard_starts(N, rank, world, bench.T_BARS, (args.updates + 2) * H + 64, 256)))

    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)  # identical replicas on every rank
    net = ActorCritic(D).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr, eps=1e-5)
    policy = env.make_policy(net)
    buf = None
    prev_done = torch.zeros(N, dtype=torch.uint8, device=dev)

    t_roll = t_upd = 0.0
    stats = {}
    for it in range(args.updates):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        buf = env.rollout(policy, H, buffers=buf, seed=1000 * rank + it)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        stats = ppo_update(net, opt, buf, prev_done, dist, epochs=args.epochs, minibatches=args.minibatches)
        prev_done = buf["done"][-1].clone()
        policy.set_weights(net)                           # bf16 repack of the two hidden layers for the next rollout
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        if it > 0:                                        # the first iteration pays one-off set-up costs
            t_roll += t1 - t0
            t_upd += t2 - t1
        stats.update(mean_reward=float(buf["reward"].mean()), terminated_frac=float(buf["done"].float().mean()))

    n_it = max(1, args.updates - 1)
    tt = torch.tensor([t_roll, t_upd], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd = float(tt[0]), float(tt[1])
    if rank == 0:
        steps = N * world * H * n_it
        print(json.dumps({
            "example": "ppo_rollout", "workload": desc, "n_gpus": world, "envs_per_gpu": N, "horizon": H,
            "policy": f"MLP({D},256,256) actor-critic, fused tcgen05 kernel in the rollout, torch (tf32) in the update",
            "rollout_env_steps_per_s": steps / max(t_roll, 1e-9), "rollout_us_per_step": t_roll / (H * n_it) * 1e6,
            "train_env_steps_per_s": steps / max(t_roll + t_upd, 1e-9), "update_s": t_upd / n_it, **stats}), flush=True)
    env.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
