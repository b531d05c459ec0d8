#!/usr/bin/env python3
"""Debug: the kernel chain of the closed loop (policy -> step -> policy ...).  Timing build + FXENV_TIMELINE: CTA 0 of every
policy / step kernel of fxenv_rollout logs globaltimer at entry, after griddepcontrol.wait and at its exit.
usage: FXENV_LIB=libfxenv_timing.so chain_probe.py [H] [envs]   (one env group: envs <= 2048 or FXENV_ROLLOUT_GROUPS=1)"""
import ctypes as C, os, sys
os.environ.setdefault("FXENV_TIMELINE", "8")
os.environ.setdefault("FXENV_LIB", "libfxenv_timing.so")
os.environ.setdefault("FXENV_ROLLOUT_GROUPS", "1")
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, torch
import bench
from gym_fx_b200.learner import ActorCritic
from gym_fx_b200.synth import start_offsets
from gym_fx_b200.vec_env import VecFxEnv

H = int(sys.argv[1]) if len(sys.argv) > 1 else 16
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg, candles, minutes, N, D, _, desc = bench.build_workload("cfg4", envs)
env = VecFxEnv(cfg, candles, minutes)
env.reset(torch.as_tensor(start_offsets(N, bench.T_BARS, 4000, 256)))
torch.manual_seed(0)
pol = env.make_policy(ActorCritic(D).cuda())
buf = env.rollout(pol, H, seed=0)
for _ in range(5):
    env.rollout(pol, H, buffers=buf, seed=1)
torch.cuda.synchronize()
env.L.fxenv_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
tl = np.zeros((8, N, 2), np.int64)
assert env.L.fxenv_debug_timeline(env._h, tl.ctypes.data) == 8
flat = tl.reshape(-1)
n = int(flat[0]); rec = flat[8:8 + 4 * 1024].reshape(1024, 4)
last = [rec[(n - 1 - i) % 1024] for i in range(min(n, 2 * H))][::-1]   # the most recent kernels, oldest first
t0 = last[0][1]
print(f"{desc}\n{n} kernels logged; the last {len(last)} (us relative to the first entry):")
print(" kind    entry  after-wait   exit(CTA0) | wait-entry  run(CTA0)  entry - previous kernel's after-wait")
prev_wait = None
for k, e, w, x in last:
    name = "policy" if k == 0 else "step  "
    d = "" if prev_wait is None else f"{(e - prev_wait) / 1e3:8.2f}"
    print(f" {name} {(e - t0) / 1e3:8.2f} {(w - t0) / 1e3:10.2f} {(x - t0) / 1e3:10.2f} | {(w - e) / 1e3:8.2f} {(x - w) / 1e3:9.2f} {d}")
    prev_wait = w
