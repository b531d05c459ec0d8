#!/usr/bin/env python3
"""
examples/ppo_rollout.py -- the env with a policy in the loop (BASELINE configs[3] shape).

One process per GPU.  Each rank owns a shard of envs (gym_fx_b200.VecFxEnv, no collective in the step path) and a
replica of an actor-critic MLP(256, 256).  A rollout of H steps -- policy forward -> sample -> fxenv_step, H times --
is captured ONCE as a CUDA graph (observations never leave the GPU: the env writes step t+1's observation straight into
the rollout buffer the policy reads next), then PPO updates run on the collected batch with gradients averaged by ONE
flat NCCL all-reduce per minibatch and advantages normalised with global statistics (gym_fx_b200.sharding).

    python examples/ppo_rollout.py --envs 4096 --horizon 32 --updates 3
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/ppo_rollout.py --envs 4096

Prints one JSON line: env-steps/s of the rollout phase (policy in the loop) and of the whole train loop.
The policy / learner are plain torch (cuBLAS GEMMs): they are the CALLER of the hot path, not part of it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


class ActorCritic(nn.Module):
    def __init__(self, obs_dim: int, hidden: int = 256, n_actions: int = 3):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.pi = nn.Linear(hidden, n_actions)
        self.v = nn.Linear(hidden, 1)

    def forward(self, obs):
        h = self.body(obs)
        return self.pi(h), self.v(h).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--horizon", type=int, default=32, help="steps per rollout")
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--workload", default="cfg4", help="bench.py workload shape (cfg4 = W=128, fixed SL/TP, sharpe)")
    ap.add_argument("--no-graph", action="store_true", help="eager rollout instead of the captured CUDA graph")
    ap.add_argument("--lr", type=float, default=3e-4)
    args = ap.parse_args()

    import bench
    from gym_fx_b200.sharding import allreduce_mean_grads, check_pair_alignment, global_mean_std, shard_starts
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
    cfg.auto_reset = 1
    env = VecFxEnv(cfg, candles, minutes, device=dev)
    check_pair_alignment(N, cfg.num_pairs)
    H = args.horizon
    env.reset(torch.as_tensor(shard_starts(N, rank, world, bench.T_BARS, (args.updates + 2) * H + 64, 256)))

    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)  # identical replicas on every rank
    net = ActorCritic(D).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr, eps=1e-5)
    params = list(net.parameters())

    # rollout buffers; obs[t] is what the policy sees at step t, obs[H] bootstraps the value of the last state
    obs = torch.zeros((H + 1, N, D), device=dev)
    act = torch.zeros((H, N), dtype=torch.int32, device=dev)
    act64 = torch.zeros((H, N), dtype=torch.int64, device=dev)
    logp = torch.zeros((H, N), device=dev)
    val = torch.zeros((H + 1, N), device=dev)
    rew = torch.zeros((H, N), device=dev)
    done = torch.zeros((H, N), device=dev)
    obs[0].copy_(env.obs)

    gumbel = torch.zeros((H, N, 3), device=dev)
    rew_v, done_v = rew.view(H, N), done.view(H, N)

    def rollout():
        with torch.no_grad():
            # Gumbel noise for the whole rollout in three kernels instead of four per step
            gumbel.uniform_(1e-9, 1.0 - 1e-9)
            gumbel.log_().neg_().log_().neg_()
            for t in range(H):
                logits, v = net(obs[t])
                torch.argmax(logits + gumbel[t], dim=-1, out=act64[t])                # Gumbel-max sample
                act[t].copy_(act64[t])
                logp[t].copy_(torch.log_softmax(logits, -1).gather(1, act64[t][:, None]).squeeze(1))
                val[t].copy_(v)
                _, r, term, _, _ = env.step(act[t], out_obs=obs[t + 1])               # fxenv_step on this stream
                rew_v[t].copy_(r)
                done_v[t].copy_(term)
            val[H].copy_(net(obs[H])[1])

    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            rollout()                                     # warm-up outside capture (cuBLAS handles, allocator)
            obs[0].copy_(obs[H])
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            rollout()
        obs[0].copy_(obs[H])

    def run_rollout():
        if graph is not None:
            graph.replay()
        else:
            rollout()

    gamma, lam, clip, vf_c, ent_c = 0.99, 0.95, 0.2, 0.5, 0.01
    t_roll = t_upd = 0.0
    stats = {}
    for it in range(args.updates):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        run_rollout()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        with torch.no_grad():                             # GAE(lambda)
            adv = torch.zeros((H, N), device=dev)
            last = torch.zeros(N, device=dev)
            for t in reversed(range(H)):
                nd = 1.0 - done[t]
                delta = rew[t] + gamma * val[t + 1] * nd - val[t]
                last = delta + gamma * lam * nd * last
                adv[t] = last
            ret = adv + val[:H]
            m, s = global_mean_std(adv, dist)
            adv = (adv - m) / s
        b_obs, b_act = obs[:H].reshape(H * N, D), act64.reshape(-1)
        b_logp, b_adv, b_ret = logp.reshape(-1), adv.reshape(-1), ret.reshape(-1)
        mb = (H * N) // args.minibatches
        for _ in range(args.epochs):
            perm = torch.randperm(H * N, device=dev)
            for k in range(args.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                logits, v = net(b_obs[idx])
                lp = torch.log_softmax(logits, -1)
                new_logp = lp.gather(1, b_act[idx, None]).squeeze(1)
                ratio = torch.exp(new_logp - b_logp[idx])
                pg = -torch.min(ratio * b_adv[idx], torch.clamp(ratio, 1 - clip, 1 + clip) * b_adv[idx]).mean()
                vloss = F.mse_loss(v, b_ret[idx])
                ent = -(lp.exp() * lp).sum(-1).mean()
                loss = pg + vf_c * vloss - ent_c * ent
                opt.zero_grad(set_to_none=True)
                loss.backward()
                allreduce_mean_grads(params, dist)        # one flat NCCL all-reduce
                nn.utils.clip_grad_norm_(params, 0.5)
                opt.step()
        obs[0].copy_(obs[H])                              # the next rollout continues from the last observation
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        if it > 0:                                        # the first iteration pays one-off set-up costs
            t_roll += t1 - t0
            t_upd += t2 - t1
        stats = {"loss": float(loss.detach()), "entropy": float(ent.detach()), "mean_reward": float(rew.mean()),
                 "terminated_frac": float(done.mean())}

    n_it = max(1, args.updates - 1)
    tt = torch.tensor([t_roll, t_upd], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd = float(tt[0]), float(tt[1])
    if rank == 0:
        steps = N * world * H * n_it
        print(json.dumps({
            "example": "ppo_rollout", "workload": desc, "n_gpus": world, "envs_per_gpu": N, "horizon": H,
            "policy": f"MLP({D},256,256) actor-critic, tf32 matmul, torch", "rollout": "CUDA graph" if graph is not None else "eager",
            "rollout_env_steps_per_s": steps / max(t_roll, 1e-9), "rollout_us_per_step": t_roll / (H * n_it) * 1e6,
            "train_env_steps_per_s": steps / max(t_roll + t_upd, 1e-9), "update_s": t_upd / n_it, **stats}), flush=True)
    env.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
