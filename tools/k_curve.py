#!/usr/bin/env python3
"""Debug: device time of one fxenv_step_many launch as a function of its length K (release library, CUDA events only,
back-to-back launches in one process after a long warm-up) -> marginal cost per step along the launch."""
import os, sys
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, torch
import bench
from gym_fx_b200.synth import start_offsets
from gym_fx_b200.vec_env import VecFxEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
Ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [10, 20, 30, 40, 60, 80, 100, 150, 200, 300, 500]
cfg, candles, minutes, N, D, algo, desc = bench.build_workload(wl)
env = VecFxEnv(cfg, candles, minutes)
env.reset(torch.as_tensor(start_offsets(N, bench.T_BARS, 60000, 256)))
KM = max(Ks)
acts = torch.randint(0, 3, (KM, N), dtype=torch.int32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
slots = max(2, -(-int(bench.L2_BYTES * 1.8) // (N * D * 4)))
ring = torch.empty((slots, N, D), dtype=torch.float32, device="cuda")
rews = torch.empty((KM, N), dtype=torch.float32, device="cuda"); terms = torch.empty((KM, N), dtype=torch.uint8, device="cuda")
plans = {K: env.plan_step_many(acts[:K], ring, rews[:K], terms[:K]) for K in Ks}
for _ in range(3):
    plans[KM]()
torch.cuda.synchronize()
print(desc, "engine", env.step_many_engine(KM))
res = {}
for rep in range(3):
    for K in Ks:
        R = max(2, 1500 // K)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(R):
            plans[K]()
        ev1.record(); torch.cuda.synchronize()
        res.setdefault(K, []).append(ev0.elapsed_time(ev1) * 1e3 / R)
prev = None
for K in Ks:
    t = float(np.median(res[K]))
    marg = "" if prev is None else f"  marginal {(t - prev[1]) / (K - prev[0]):6.2f} us/step over steps {prev[0]}..{K}"
    print(f"K={K:4d}: {t:9.1f} us per launch = {t / K:6.2f} us/step{marg}")
    prev = (K, t)
