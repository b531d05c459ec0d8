"""GPU (-m gpu): the randomised differential test of tests/test_core_vs_oracle_random.py, run through the CUDA KERNEL
(VecFxEnv -> C-ABI -> fx_step_kernel) instead of the g++ build of the scalar core: the same 36 seeded configurations
(commission x leverage x slippage x position size x strategy x reward x sizing mode), 64 envs each with their own start
bar and sticky action stream, against the C oracle, step by step.  This is what exercises the warp-parallel part of the
broker pass -- the 32-entry chunk sweep, stable compaction, the parent-in-lane-31 carry into the next chunk, the cash
bound of check_submitted falling through to the exact simulation -- under margin-heavy random configurations.

Bars: integer state and fp64 account state bit-exact at every step, reward 1e-9 (fp64) / 1e-5 (fp32) relative,
observations rtol 1e-5 / atol 2e-6."""
import numpy as np
import pytest
import torch

import scenarios as S
from common_gpu import GpuVec, compare_info, compare_step, compare_summary, gpu_summary
from gym_fx_b200.config import lower_config
from gym_fx_b200.synth import synth_candles, synth_minutes
from oracle.c_oracle import OracleVec
from test_core_vs_oracle_random import _random_case

pytestmark = pytest.mark.gpu

N = 64


def _streams(seed, steps, cont=False):
    acts = np.empty((steps, N), np.int32)
    rng = np.random.default_rng(seed)
    for i in range(N):
        acts[:, i] = S.make_actions(("sticky", 7000 + 97 * seed + i, float(rng.choice([0.2, 0.6, 0.85]))), steps)
    return acts


@pytest.mark.parametrize("seed", range(36))
def test_cuda_kernel_matches_oracle_on_random_configs(seed):
    cfg1, _, _, _ = _random_case(seed)       # the configuration of the CPU test (num_envs = 1) ...
    cfg = type(cfg1).from_buffer_copy(cfg1)  # ... for 64 envs on a longer table
    cfg.num_envs = N
    cfg.order_capacity = 512
    T, steps = 2600, 560
    rng = np.random.default_rng(100 + seed)
    pair = int(rng.integers(0, 4))
    candles, minutes = synth_candles(T, pair, 5000 + seed), synth_minutes(T)
    if pair == 3:
        cfg.pair_pip_size[0] = 0.01
    starts = rng.integers(0, T - steps - 40, N).astype(np.int64)
    starts[:4] = [0, 1, 2, 3]
    acts = _streams(seed, steps)
    gpu, orc = GpuVec(cfg, [candles], [minutes]), OracleVec(cfg, [candles], [minutes])
    np.testing.assert_allclose(gpu.reset(starts), orc.reset(starts), rtol=1e-5, atol=2e-6, err_msg="reset obs")
    deepest = 0
    for k in range(steps):
        want = (k % 8 == 0) or k == steps - 1
        compare_step(f"seed {seed} step {k}", gpu.step(acts[k], want_obs=want), orc.step(acts[k], want_obs=want))
        if k % 4 == 0 or k == steps - 1:
            gi = gpu.info()
            compare_info(f"seed {seed} step {k}", gi, orc.info())
            deepest = max(deepest, int(gi["n_orders"].max()))
            assert not np.any(gi["flags"] & 16), f"seed {seed}: order table overflow at step {k}"
    compare_summary(f"seed {seed}", gpu_summary(gpu.env), orc.summary())
    gpu.close()
    orc.close()


def test_cuda_kernel_deep_order_tables_cross_chunk_carry():
    """Wide brackets + restless actions pile up > 64 live entries per env (stale parents, orphaned pairs): the sweep runs
    over 3+ chunks per env-step, parents land in lane 31 with their pair in the next chunk, compaction moves entries
    across chunk boundaries.  Margin-heavy sizing makes the cash bound fail so the exact check_submitted path runs too."""
    cfgd = {**S.DEFAULTS, "window_size": 8, "sl_pips": 60.0, "tp_pips": 90.0, "position_size": 1500.0, "commission": 2e-5}
    pl = S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, "strategy": "direct_fixed_sltp"})
    cfg = lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"], preprocessor_plugin=pl["preprocessor"],
                       reward_plugin=pl["reward"], columns=S.OHLCV, num_envs=N, order_capacity=512)
    T, steps = 3000, 700
    candles, minutes = synth_candles(T, 0, 4242), synth_minutes(T)
    rng = np.random.default_rng(9)
    starts = rng.integers(0, T - steps - 20, N).astype(np.int64)
    gpu, orc = GpuVec(cfg, [candles], [minutes]), OracleVec(cfg, [candles], [minutes])
    gpu.reset(starts); orc.reset(starts)
    deepest, margin_seen = 0, False
    for k in range(steps):
        a = rng.integers(0, 3, N).astype(np.int32)
        compare_step(f"deep step {k}", gpu.step(a, want_obs=(k % 16 == 0)), orc.step(a, want_obs=(k % 16 == 0)))
        if k % 5 == 0 or k == steps - 1:
            gi = gpu.info()
            compare_info(f"deep step {k}", gi, orc.info())
            deepest = max(deepest, int(gi["n_orders"].max()))
            margin_seen |= bool(np.any(gi["cash"] < 1500.0 * 1.2))
    assert deepest > 64, f"only {deepest} live entries: the cross-chunk paths were not exercised"
    assert margin_seen, "cash never got close to the order size: the margin paths were not exercised"
    compare_summary("deep", gpu_summary(gpu.env), orc.summary())
    assert not np.any(gpu.info()["flags"] & 16)
    gpu.close()
    orc.close()
