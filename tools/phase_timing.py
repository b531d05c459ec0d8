#!/usr/bin/env python3
"""Debug: per-phase clock64() durations of the step kernel in a steady-state CUDA-graph run (FXENV_TIMING=1)."""
import ctypes as C, os, sys
os.environ["FXENV_TIMING"] = "1"
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, torch
import bench
from gym_fx_b200.synth import start_offsets
from gym_fx_b200.vec_env import VecFxEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg, candles, minutes, N, D, algo, desc = bench.build_workload(wl, envs)
env = VecFxEnv(cfg, candles, minutes)
env.reset(torch.as_tensor(start_offsets(N, bench.T_BARS, 2000, 256)))
K = 100
acts = torch.randint(0, 3, (K, N), dtype=torch.int32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
slots = max(2, -(-int(bench.L2_BYTES * 1.8) // (N * D * 4)))
ring = torch.empty((slots, N, D), dtype=torch.float32, device="cuda")
rews = torch.empty((K, N), dtype=torch.float32, device="cuda"); terms = torch.empty((K, N), dtype=torch.uint8, device="cuda")
for _ in range(5):
    env.step_many(acts, ring, rews, terms)
torch.cuda.synchronize()
env.L.fxenv_debug_timings.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros((N, 10), np.int64)
assert env.L.fxenv_debug_timings(env._h, buf.ctypes.data) == 10
names = ["loads+stats", "obs(odd warps)", "stage+scan", "check_submitted", "fills+mtm", "strategy+publish", "reward", "writeback+compact", "obs(even warps)"]
t = buf.astype(np.float64)
clk = 1.965  # GHz (clocks.max.sm)
print(f"{desc}\nper-warp phase durations of the LAST step (cycles @~{clk} GHz -> us), mean / p50 / p95 / max over {N} warps")
for i, nm in enumerate(names):
    a, b = t[:, i], t[:, i + 1]
    if i == 2: a = t[:, 2]; b = np.where(t[:, 3] > 0, t[:, 3], t[:, 2])
    d = b - a
    d = d[(a > 0) & (b > 0)]
    if d.size == 0: print(f"  {nm:20s} n/a"); continue
    print(f"  {nm:20s} {d.mean():9.0f} cyc {d.mean()/clk/1e3:6.2f} us | p50 {np.median(d):8.0f} p95 {np.percentile(d,95):8.0f} max {d.max():8.0f}  (n={d.size})")
tot = t[:, 9] - t[:, 0]
print(f"  {'TOTAL warp lifetime':20s} {tot.mean():9.0f} cyc {tot.mean()/clk/1e3:6.2f} us | p50 {np.median(tot):8.0f} p95 {np.percentile(tot,95):8.0f} max {tot.max():8.0f}")
