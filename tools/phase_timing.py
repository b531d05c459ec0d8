#!/usr/bin/env python3
"""Debug: per-phase clock64() durations of the step kernel in a steady-state CUDA-graph run (FXENV_TIMING=1)."""
import ctypes as C, os, sys
os.environ["FXENV_TIMING"] = "1"
os.environ.setdefault("FXENV_LIB", "libfxenv_timing.so")  # built by `make -C gym_fx_b200/csrc timing`
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
buf2 = np.zeros((2, N, 12), np.int64)
assert env.L.fxenv_debug_timings(env._h, buf2.ctypes.data) == 12
lastslot = int(np.argmax([buf2[0,:,10].max(), buf2[1,:,10].max()]))
buf = buf2[lastslot]; prev = buf2[1 - lastslot]
t = buf.astype(np.float64)
clk = 1.965  # GHz (clocks.max.sm)
segs = [("start -> state arrived", 0, 2), ("state -> bar arrived", 2, 1),
        ("bar -> check_submitted done", 1, 3), ("order pass + mark-to-market", 3, 5), ("strategy + publish", 5, 6),
        ("reward", 6, 7), ("write-back", 7, 8), ("wait window + emit obs", 8, 9), ("TOTAL warp lifetime", 0, 9)]
print(f"{desc}\nper-warp phase durations of the LAST step (cycles @~{clk} GHz -> us), over {N} warps")
for nm, a, b in segs:
    x, y = t[:, a], t[:, b]
    ok = (x > 0) & (y > 0)
    d = (y - x)[ok]
    if d.size == 0:
        print(f"  {nm:32s} n/a"); continue
    print(f"  {nm:32s} {d.mean():9.0f} cyc {d.mean()/clk/1e3:6.2f} us | p50 {np.median(d):8.0f} p95 {np.percentile(d,95):8.0f} max {d.max():8.0f}  (n={d.size})")
g0, g1 = buf[:, 10].astype(np.float64), buf[:, 11].astype(np.float64)
base = g0.min()
print(f"  globaltimer (ns): first warp start 0, last warp start {g0.max()-base:.0f}, p50 start {np.median(g0)-base:.0f}, p95 start {np.percentile(g0,95)-base:.0f}; "
      f"last warp end {g1.max()-base:.0f}; warp lifetime mean {(g1-g0).mean():.0f} max {(g1-g0).max():.0f}")
order = np.argsort(g0)
print("  start time (ns) of every 128th warp by start order:", [int(g0[order[i]]-base) for i in range(0, N, max(1, N//32))])
print("  env ids of the 8 last-starting warps:", order[-8:].tolist(), " and 8 first:", order[:8].tolist())
pg0, pg1 = prev[:, 10].astype(np.float64), prev[:, 11].astype(np.float64)
print(f"  consecutive steps: prev kernel first start {pg0.min()-base:.0f} ns, prev last end {pg1.max()-base:.0f} ns -> period {g0.min()-pg0.min():.0f} ns, "
      f"gap (prev last warp end -> this first warp start) {g0.min()-pg1.max():.0f} ns, active span {g1.max()-g0.min():.0f} ns")

tot = (t[:, 9] - t[:, 0]) / clk / 1e3
nf = (buf[:, 4] & 0xffffffff).astype(np.int64); ntab = (buf[:, 4] >> 32).astype(np.int64)
print("  warp lifetime us percentiles: p50 %.1f p90 %.1f p99 %.1f p99.9 %.1f max %.1f" % tuple(np.percentile(tot, [50, 90, 99, 99.9, 100])))
print("  fills per env-step: mean %.2f p50 %d p90 %d p99 %d max %d ; table entries: mean %.1f p99 %d max %d" % (nf.mean(), *np.percentile(nf, [50, 90, 99, 100]).astype(int), ntab.mean(), *np.percentile(ntab, [99, 100]).astype(int)))
for lo, hi in ((0, 0), (1, 1), (2, 3), (4, 7), (8, 15), (16, 1000)):
    sel = (nf >= lo) & (nf <= hi)
    if sel.any(): print(f"    fills {lo}-{hi}: {sel.sum():5d} warps, lifetime mean {tot[sel].mean():6.2f} us max {tot[sel].max():6.2f}")

# what would a work-sorted launch order buy?  greedy list scheduling of the measured warp lifetimes on 148 x 16 slots
import heapq
def makespan(durs, nslots=148 * 16):
    h = [0.0] * nslots
    heapq.heapify(h)
    end = 0.0
    for d in durs:
        s = heapq.heappop(h); heapq.heappush(h, s + d); end = max(end, s + d)
    return end
life = (g1 - g0) / 1e3
print("  list-scheduling model (us): index order %.1f | sorted by table entries desc %.1f | by fills desc %.1f | by lifetime desc (ideal) %.1f | sum/slots %.1f | max %.1f"
      % (makespan(life), makespan(life[np.argsort(-ntab, kind='stable')]), makespan(life[np.argsort(-nf, kind='stable')]),
         makespan(np.sort(life)[::-1]), life.sum() / (148 * 16), life.max()))
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/phase_{wl}_{N}.npz", stamps=buf2)
