#!/usr/bin/env python3
"""Debug: where does a K-step fxenv_step_many batch (persistent ticket kernel) spend its time?  Timing build +
FXENV_TIMELINE: globaltimer at the start and end of every (step, env) ticket.  usage: timeline.py [workload] [K]"""
import ctypes as C, os, sys
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
os.environ["FXENV_TIMELINE"] = str(K)
os.environ.setdefault("FXENV_LIB", "libfxenv_timing.so")
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, torch
import bench
from gym_fx_b200.synth import start_offsets
from gym_fx_b200.vec_env import VecFxEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, candles, minutes, N, D, algo, desc = bench.build_workload(wl)
env = VecFxEnv(cfg, candles, minutes)
env.reset(torch.as_tensor(start_offsets(N, bench.T_BARS, 4000, 256)))
acts = torch.randint(0, 3, (K, N), dtype=torch.int32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
slots = max(2, -(-int(bench.L2_BYTES * 1.8) // (N * D * 4)))
ring = torch.empty((slots, N, D), dtype=torch.float32, device="cuda")
rews = torch.empty((K, N), dtype=torch.float32, device="cuda"); terms = torch.empty((K, N), dtype=torch.uint8, device="cuda")
plan = env.plan_step_many(acts, ring, rews, terms)
for _ in range(30):
    plan()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); plan(); ev1.record(); torch.cuda.synchronize()
print(f"{desc}\nK={K}: event time {ev0.elapsed_time(ev1)*1e3:.1f} us = {ev0.elapsed_time(ev1)*1e3/K:.2f} us/step")
env.L.fxenv_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
tl = np.zeros((K, N, 2), np.int64)
assert env.L.fxenv_debug_timeline(env._h, tl.ctypes.data) == K
t0 = tl[:, :, 0].min()
st, en = (tl[:, :, 0] - t0) / 1e3, (tl[:, :, 1] - t0) / 1e3   # us
print(f"first ticket start 0, last ticket end {en.max():.1f} us; env-step duration mean {np.mean(en-st):.2f} p50 {np.median(en-st):.2f} p99 {np.percentile(en-st,99):.2f} max {np.max(en-st):.2f} us")
print(" step | first start  median start  last start | median end  last end | mean dur | gap to dep (start - end of same env's previous step): mean  p99")
show = range(K) if K <= 64 else sorted(set(list(range(0, 8)) + list(range(8, K, max(1, K // 48))) + [K - 1]))
med = np.median(st, axis=1)
print("step period (us, median start of step k+10 minus step k, /10):", " ".join(f"{(med[k+10]-med[k])/10:.1f}" for k in range(0, K - 10, max(1, K // 40))))
for k in show:
    gap = st[k] - en[k - 1] if k else np.zeros(N)
    print(f"  {k:3d} | {st[k].min():9.1f} {np.median(st[k]):12.1f} {st[k].max():11.1f} | {np.median(en[k]):9.1f} {en[k].max():9.1f} | {np.mean(en[k]-st[k]):7.2f} | {gap.mean():8.2f} {np.percentile(gap,99):8.2f}")
# busy warps over time
edges = np.linspace(0, en.max(), 41)
busy = [(np.minimum(en, b) - np.maximum(st, a)).clip(min=0).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
print("busy warps (of %d) per time slice: %s" % (148 * 16, " ".join(f"{x:.0f}" for x in busy)))
per_env_chain = en[-1] - st[0]
print(f"per-env chain (first start -> last end): mean {per_env_chain.mean():.1f} max {per_env_chain.max():.1f} us; sum of its env-step durations: mean {np.sum(en-st,axis=0).mean():.1f} us")
