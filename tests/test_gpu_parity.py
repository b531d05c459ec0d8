"""GPU (-m gpu) parity tests: the CUDA env (through the C-ABI, libfxenv.so) against
  (1) the committed golden trajectories of the reference (tests/golden/*.npz), and
  (2) the C oracle on the same seeded inputs, many envs in lockstep,
plus full-size (BASELINE cfg2: 4096 envs, W=128, F=5) size-independent properties.

Bars: integer/index state and fp64 account state bit-exact; rewards 1e-5 relative in fp32 (north_star) and 1e-9
in fp64; observations rtol 1e-5 / atol 2e-6 in fp32 (the z-score uses a reciprocal multiply and a warp-ordered
sum, the oracle divides and sums sequentially)."""
import numpy as np
import pytest
import torch

import scenarios as S
from common import assert_summary_matches, assert_traj_matches, config_from_meta, golden_names, load_golden, replay
from common_gpu import GpuVec, compare_info, compare_step, compare_summary, gpu_summary
from gym_fx_b200.config import lower_config
from gym_fx_b200.synth import start_offsets, synth_candles, synth_minutes
from oracle.c_oracle import OracleVec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"


# ---------------------------------------------------------------------------------------------- (1) goldens
@pytest.mark.parametrize("name", golden_names())
def test_gpu_matches_reference_golden(name):
    g = load_golden(name)
    cfg = config_from_meta(g["meta"], order_capacity=512)
    env = GpuVec(cfg, [g["candles"]], [g["minutes"]])
    traj = replay(env, g, env.info)
    assert not np.any(env.info()["flags"] & 16), "order table overflow"
    summ = gpu_summary(env.env, 0)
    env.close()
    assert_traj_matches(traj, g, obs_rtol=1e-5, obs_atol=2e-6, label=name)
    # the reference's GymFxEnv.summary() after the run (DrawDown / TradeAnalyzer / SQN): section 8f #3
    assert_summary_matches(summ, g["meta"]["summary"], label=name)


def test_gpu_reference_known_answer():
    # examples/results/buy_hold_summary.json:3-4 of the reference, produced through REAL backtrader
    g = load_golden("buy_hold_uptrend")
    env = GpuVec(config_from_meta(g["meta"]), [g["candles"]], [g["minutes"]])
    traj = replay(env, g, env.info)
    assert traj["equity"][-1] == 10000.095791583166


# ---------------------------------------------------------------------------------------------- (2) vs oracle
def _mk(cfgd, plugins, N, T=4096, pairs=1, columns=None, **kw):
    cfgd = {**S.DEFAULTS, **cfgd}
    pl = S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, **plugins})
    cfg = lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"],
                       preprocessor_plugin=pl["preprocessor"], reward_plugin=pl["reward"],
                       columns=columns or S.OHLCV, num_envs=N, num_pairs=pairs, **kw)
    candles = [synth_candles(T, p) for p in range(pairs)]
    minutes = [synth_minutes(T) for _ in range(pairs)]
    return cfg, candles, minutes


FW = {"feature_columns": list(S.OHLCV)}
VEC_CASES = {
    # BASELINE configs[1] shape: feature_window W=128 F=5 S=256, direct_fixed_sltp, pnl
    "cfg2_fixed_fw128_pnl": (dict(window_size=128, **FW), dict(strategy="direct_fixed_sltp",
                             preprocessor="feature_window_preprocessor"), {}),
    # configs[2] shape: W=256, direct_atr_sltp, dd_penalized
    "cfg3_atr_fw256_dd": (dict(window_size=256, **FW), dict(strategy="direct_atr_sltp",
                          preprocessor="feature_window_preprocessor", reward="dd_penalized_reward"), {}),
    # configs[3] shape: W=128 + sharpe
    "cfg4_fixed_fw128_sharpe": (dict(window_size=128, **FW), dict(strategy="direct_fixed_sltp",
                                preprocessor="feature_window_preprocessor", reward="sharpe_reward"), {}),
    # configs[4] shape (scaled down): 4 pairs interleaved, W=512, atr, sharpe
    "cfg5_atr_fw512_sharpe_4pairs": (dict(window_size=512, **FW), dict(strategy="direct_atr_sltp",
                                     preprocessor="feature_window_preprocessor", reward="sharpe_reward"),
                                     dict(pairs=4, pair_pip_size=[1e-4, 1e-4, 1e-4, 1e-2])),
    "cfg1_default": (dict(window_size=32), {}, {}),
    "commission_leverage_relvol": (dict(window_size=16, commission=2e-5, leverage=5.0, rel_volume=0.3,
                                        max_order_volume=9000.0), dict(strategy="direct_atr_sltp"), {}),
    "slippage_fixed_brackets": (dict(window_size=16, slippage_perc=2e-4, commission=1e-5, sl_pips=4.0, tp_pips=6.0),
                                dict(strategy="direct_fixed_sltp"), {}),
    "margin_heavy_fixed": (dict(window_size=8, position_size=6000.0, commission=5e-5, sl_pips=4.0, tp_pips=6.0),
                           dict(strategy="direct_fixed_sltp", reward="dd_penalized_reward"), {}),
}


@pytest.mark.parametrize("case", sorted(VEC_CASES))
def test_gpu_matches_oracle_vectorised(case):
    cfgd, plugins, kw = VEC_CASES[case]
    N, T, steps = 192, 6000, 330
    cfg, candles, minutes = _mk(cfgd, plugins, N, T=T, order_capacity=512, **kw)
    starts = start_offsets(N, T, steps + 10, 300)
    starts[:8] = [0, 1, 2, 5, 100, 255, 256, 257]  # warm-up edge offsets
    gpu, orc = GpuVec(cfg, candles, minutes), OracleVec(cfg, candles, minutes)
    np.testing.assert_allclose(gpu.reset(starts), orc.reset(starts), rtol=1e-5, atol=2e-6, err_msg="reset obs")
    rng = np.random.default_rng(2024)
    inexact = 0
    for k in range(steps):
        a = rng.integers(0, 3, N).astype(np.int32)
        inexact += compare_step(f"{case} step {k}", gpu.step(a), orc.step(a))
        if k % 25 == 0 or k == steps - 1:
            compare_info(f"{case} step {k}", gpu.info(), orc.info())
    assert not np.any(gpu.info()["flags"] & 16), "order table overflow"
    compare_summary(case, gpu_summary(gpu.env), orc.summary())
    total = steps * N * gpu.env.obs_dim
    assert inexact <= max(10, total * 1e-4), f"{inexact}/{total} obs floats not bit-identical to the oracle"
    gpu.close()


def test_gpu_episode_end_and_auto_reset():
    # data exhaustion (A.7), stepping after termination, next-step auto reset, masked reset
    for auto in (False, True):
        cfg, candles, minutes = _mk(dict(window_size=8, sl_pips=3.0, tp_pips=4.0), dict(strategy="direct_fixed_sltp",
                                    reward="sharpe_reward"), 64, T=600, episode_bars=40, auto_reset=auto)
        starts = (np.arange(64) * 7) % 500
        gpu, orc = GpuVec(cfg, candles, minutes), OracleVec(cfg, candles, minutes)
        gpu.reset(starts); orc.reset(starts)
        rng = np.random.default_rng(5)
        for k in range(95):
            a = rng.integers(0, 3, 64).astype(np.int32)
            compare_step(f"auto={auto} step {k}", gpu.step(a), orc.step(a))
            compare_info(f"auto={auto} step {k}", gpu.info(), orc.info())
            if k % 10 == 0 or k == 94:
                compare_summary(f"auto={auto} step {k}", gpu_summary(gpu.env), orc.summary())
            if k == 50:
                mask = (np.arange(64) % 3 == 0).astype(np.uint8)
                np.testing.assert_allclose(gpu.reset(None, mask), orc.reset(None, mask), rtol=1e-5, atol=2e-6)
        gpu.close()


def test_gpu_continuous_actions_and_bad_discrete():
    cfg, candles, minutes = _mk(dict(window_size=8, action_space_mode="continuous"), dict(strategy="direct_fixed_sltp"), 32)
    gpu, orc = GpuVec(cfg, candles, minutes), OracleVec(cfg, candles, minutes)
    gpu.reset(np.zeros(32, np.int64)); orc.reset(np.zeros(32, np.int64))
    rng = np.random.default_rng(1)
    for k in range(120):
        a = rng.uniform(-1, 1, 32).astype(np.float32)
        a[k % 32] = 0.33 if k % 2 else -0.33  # exactly on the threshold
        compare_step(f"cont {k}", gpu.step(a), orc.step(a))
    gpu.close()
    cfg, candles, minutes = _mk(dict(window_size=8), {}, 32)
    gpu, orc = GpuVec(cfg, candles, minutes), OracleVec(cfg, candles, minutes)
    gpu.reset(np.zeros(32, np.int64)); orc.reset(np.zeros(32, np.int64))
    for k in range(60):
        a = rng.integers(-3, 6, 32).astype(np.int32)  # out-of-range ints are coerced to hold (app/env.py:200-204)
        compare_step(f"bad {k}", gpu.step(a), orc.step(a))
    gpu.close()


def test_gpu_order_overflow_flag_is_loud():
    cfg, candles, minutes = _mk(dict(window_size=8), dict(strategy="direct_fixed_sltp"), 32, order_capacity=32)
    gpu = GpuVec(cfg, candles, minutes)
    gpu.reset(np.zeros(32, np.int64))
    rng = np.random.default_rng(3)
    for k in range(400):
        gpu.step(rng.integers(0, 3, 32).astype(np.int32), want_obs=False)
    inf = gpu.info()
    assert np.any(inf["flags"] & 16), "a 32-entry table must overflow under random actions (reference is unbounded)"
    assert inf["n_orders"].max() <= 32
    gpu.close()


# ---------------------------------------------------------------------------------------------- (3) full size
@pytest.fixture(scope="module")
def cfg2_full():
    N, T = 4096, 1 << 17
    cfg, candles, minutes = _mk(dict(window_size=128, **FW), dict(strategy="direct_fixed_sltp",
                                preprocessor="feature_window_preprocessor"), N, T=T)
    return N, T, cfg, candles, minutes


def test_full_size_flat_policy_keeps_equity(cfg2_full):
    # tools/smoke_test.py:113-118 of the reference: the flat driver leaves equity unchanged
    N, T, cfg, candles, minutes = cfg2_full
    gpu = GpuVec(cfg, candles, minutes)
    gpu.reset(start_offsets(N, T, 300, 256))
    z = np.zeros(N, np.int32)
    for k in range(40):
        obs, rew, rew64, term = gpu.step(z, want_obs=(k == 39))
        assert np.all(rew64 == 0.0) and not term.any()
    inf = gpu.info()
    assert np.all(inf["equity"] == 10000.0) and np.all(inf["trades"] == 0) and np.all(inf["bar_index"] == 40)
    assert np.all(np.isfinite(obs)) and np.all(np.abs(obs[:, :640]) <= 10.0)
    gpu.close()


def test_full_size_shard_invariance_determinism_and_sampled_oracle(cfg2_full):
    # env i's trajectory must not depend on how many envs share the launch (what multi-GPU sharding relies on),
    # must be reproducible, and a sample of envs must match the oracle at the full 4096-env size.
    N, T, cfg, candles, minutes = cfg2_full
    steps = 120
    starts = start_offsets(N, T, steps, 256)
    acts = np.random.default_rng(1234).integers(0, 3, (steps, N)).astype(np.int32)

    samp = np.arange(0, N, 64)   # the envs whose observation is kept at EVERY step (a multiple of the shard stride)

    def run(sel):
        """-> per step (obs of the kept envs, reward32, reward64, terminated) + final info.  The observation of every
        env is compared every 40th step; of the kept envs (`samp`) at every step."""
        c2, _, _ = _mk(dict(window_size=128, **FW), dict(strategy="direct_fixed_sltp",
                       preprocessor="feature_window_preprocessor"), len(sel), T=T)
        g = GpuVec(c2, candles, minutes)
        g.reset(starts[sel])
        keep = np.nonzero(np.isin(sel, samp))[0]
        outs, fulls = [], {}
        for k in range(steps):
            o = g.step(acts[k, sel], want_obs=True)
            if k % 40 == 39 or k == steps - 1:
                fulls[k] = o[0]
            outs.append((o[0][keep].copy(), o[1], o[2], o[3]))
        inf = g.info()
        g.close()
        return outs, fulls, inf

    full, ffull, finf = run(np.arange(N))
    again, afull, ainf = run(np.arange(N))
    sel = np.arange(0, N, 8)
    shard, sfull, sinf = run(sel)
    for k in range(steps):
        assert np.array_equal(full[k][0], again[k][0]) and np.array_equal(full[k][0], shard[k][0]), f"kept-env obs at step {k}"
        for j in range(1, 4):
            assert np.array_equal(full[k][j], again[k][j]), f"non-deterministic output {j} at step {k}"
            assert np.array_equal(full[k][j][sel], shard[k][j]), f"shard-dependent output {j} at step {k}"
    for k in ffull:
        assert np.array_equal(ffull[k], afull[k]) and np.array_equal(ffull[k][sel], sfull[k]), f"full obs at step {k}"
    for key in ("equity", "cash", "trades", "n_orders"):
        assert np.array_equal(finf[key][sel], sinf[key])
    c3, _, _ = _mk(dict(window_size=128, **FW), dict(strategy="direct_fixed_sltp",
                   preprocessor="feature_window_preprocessor"), len(samp), T=T)
    orc = OracleVec(c3, candles, minutes)
    orc.reset(starts[samp])
    for k in range(steps):   # the sampled envs against the oracle at EVERY step, observation included
        oo = orc.step(acts[k, samp])
        compare_step(f"full-size sample step {k}", (full[k][0], full[k][1][samp], full[k][2][samp], full[k][3][samp]), oo)
    oi = orc.info()
    assert np.array_equal(finf["equity"][samp], oi["equity"]) and np.array_equal(finf["trades"][samp], oi["trades"])


def test_step_many_graph_and_step_host_match_step(cfg2_full):
    N, T, cfg, candles, minutes = cfg2_full
    K = 24
    starts = torch.as_tensor(start_offsets(N, T, 300, 256))
    acts = torch.randint(0, 3, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(7)).cuda()
    from gym_fx_b200.vec_env import VecFxEnv
    a_env, b_env, c_env = (VecFxEnv(cfg, candles, minutes) for _ in range(3))
    for e in (a_env, b_env, c_env):
        e.reset(starts)
    D = a_env.obs_dim
    ring = torch.empty((2, N, D), dtype=torch.float32, device="cuda")
    rews = torch.empty((K, N), dtype=torch.float32, device="cuda")
    terms = torch.empty((K, N), dtype=torch.uint8, device="cuda")
    for rep in range(2):  # second call replays the cached graph
        b_env.reset(starts)
        b_env.step_many(acts, ring, rews, terms)
    torch.cuda.synchronize()
    h_act = torch.empty(N, dtype=torch.int32).pin_memory()
    h_obs = torch.empty((N, D), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
    h_term = torch.empty(N, dtype=torch.uint8).pin_memory()
    for k in range(K):
        obs, rew, term, _, _ = a_env.step(acts[k])
        h_act.copy_(acts[k])
        c_env.step_host(h_act, h_obs, h_rew, h_term)
        assert torch.equal(rew, rews[k]) and torch.equal(term.to(torch.uint8), terms[k])
        assert torch.equal(rew.cpu(), h_rew) and torch.equal(obs.cpu(), h_obs) and torch.equal(term.to(torch.uint8).cpu(), h_term)
    assert torch.equal(obs, ring[(K - 1) % 2])
    # pageable host buffers (not pinned) take the same path through staged copies: same results
    p_act, p_obs = torch.empty(N, dtype=torch.int32), torch.empty((N, D), dtype=torch.float32)
    p_rew, p_term = torch.empty(N, dtype=torch.float32), torch.empty(N, dtype=torch.uint8)
    for k in range(3):
        obs, rew, term, _, _ = a_env.step(acts[k])
        p_act.copy_(acts[k])
        c_env.step_host(p_act, p_obs, p_rew, p_term)
        assert torch.equal(rew.cpu(), p_rew) and torch.equal(obs.cpu(), p_obs) and torch.equal(term.to(torch.uint8).cpu(), p_term)
    b_env.step_many(acts[:3], ring, rews[:3], terms[:3])   # keep the twin in step for the comparisons below
    torch.cuda.synchronize()
    assert torch.equal(a_env.info()["equity"], b_env.info()["equity"])
    summ = a_env.summary()
    assert torch.equal(summ["final_equity"], a_env.info()["equity"]) and summ["order_overflow_envs"] == 0
    assert abs(summ["mean_total_return"] - float((a_env.info()["equity"] / 10000.0 - 1.0).mean())) < 1e-12
    # snapshot / restore round trip
    blob = a_env.get_state()
    ref = [a_env.step(acts[k])[1].clone() for k in range(3)]
    a_env.set_state(blob)
    for k in range(3):
        assert torch.equal(a_env.step(acts[k])[1], ref[k])
    # a blob whose header does not describe THIS env (other config hash / capacity / size) is refused, not reinterpreted
    from gym_fx_b200._native import FxEnvError
    bad = bytearray(blob); bad[24] ^= 0x5A
    with pytest.raises(FxEnvError, match="different configuration"):
        a_env.set_state(bytes(bad))
    with pytest.raises(FxEnvError, match="size mismatch"):
        a_env.set_state(blob[:-8])
    with pytest.raises(FxEnvError, match="library version"):
        a_env.set_state(b"\0" * len(blob))
    assert a_env.launch_count() >= K
    for e in (a_env, b_env, c_env):
        e.close()


def test_step_many_rollout_kernel_matches_graph_of_steps(cfg2_full):
    """fxenv_step_many has two engines (persistent ticket kernel / CUDA graph of single steps): same results, bit for bit,
    over two consecutive batches with plenty of fills, and the same state snapshot afterwards."""
    import os
    from gym_fx_b200.vec_env import VecFxEnv
    N, T, cfg, candles, minutes = cfg2_full
    K = 96
    starts = torch.as_tensor(start_offsets(N, T, 400, 256))
    acts = torch.randint(0, 3, (2, K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(11)).cuda()
    outs = []
    for eng in ("persistent", "graph"):  # the persistent launch vs the graph of grid-serialised single steps
        os.environ["FXENV_ENGINE"] = eng
        try:
            env = VecFxEnv(cfg, candles, minutes)
        finally:
            del os.environ["FXENV_ENGINE"]
        env.reset(starts)
        l0 = env.launch_count()
        ring = torch.zeros((3, N, env.obs_dim), dtype=torch.float32, device="cuda")
        rews = torch.zeros((2, K, N), dtype=torch.float32, device="cuda")
        terms = torch.zeros((2, K, N), dtype=torch.uint8, device="cuda")
        for b in range(2):
            env.step_many(acts[b], ring, rews[b], terms[b])
        torch.cuda.synchronize()
        inf = env.info()
        info = {k: inf[k].clone() for k in ("equity", "position", "price", "bar_index", "trades", "commission_paid")}
        outs.append((ring, rews, terms, info, env.get_state(), env.launch_count() - l0))
        env.close()
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k
    assert bytes(a[4]) == bytes(b[4])
    assert a[5] == 2 and b[5] == 2 * K  # one launch per batch vs one per step
    assert float(a[3]["trades"].float().mean()) > 5.0  # the batches really traded


def test_gpu_nan_inf_feature_values_follow_nan_to_num():
    """Untamed tables (NaN / +-inf in a feature column) take the general observation path: np.clip then np.nan_to_num
    (feature_window_preprocessor.py:119-123) -- NaN -> 0, +-inf -> +-clip -- exactly like the oracle."""
    N, T, steps = 64, 3000, 120
    cols = list(S.OHLCV) + ["FEAT_A"]
    cfgd = dict(window_size=16, feature_columns=["CLOSE", "VOLUME", "FEAT_A"], feature_scaling_window=32)
    cfg, candles, minutes = _mk(cfgd, dict(strategy="direct_fixed_sltp", preprocessor="feature_window_preprocessor"), N, T=T,
                                columns=cols, order_capacity=256)
    rng = np.random.default_rng(5)
    extra = rng.normal(0.0, 1.0, T)
    extra[rng.integers(0, T, 60)] = np.nan
    extra[rng.integers(0, T, 30)] = np.inf
    extra[rng.integers(0, T, 30)] = -np.inf
    candles = [np.ascontiguousarray(np.concatenate([candles[0], extra[:, None]], axis=1))]
    starts = start_offsets(N, T, steps + 10, 40)
    gpu, orc = GpuVec(cfg, candles, minutes), OracleVec(cfg, candles, minutes)
    go, oo = gpu.reset(starts), orc.reset(starts)
    assert np.array_equal(np.isfinite(go), np.isfinite(oo)) and np.all(np.isfinite(go))
    for k in range(steps):
        a = rng.integers(0, 3, N).astype(np.int32)
        g, o = gpu.step(a), orc.step(a)
        assert np.all(np.isfinite(g[0]))
        compare_step(f"nan/inf step {k}", g, o)
    gpu.close()


@pytest.mark.parametrize("case", ["cfg5_atr_fw512_sharpe_4pairs", "cfg3_atr_fw256_dd", "slippage_fixed_brackets"])
def test_step_many_engines_agree_on_other_kernels(case):
    """Same check as above for the other template instantiations of the persistent kernel (ATR sizing, Sharpe ring staged
    per env-step in shared memory that the warp reuses for its next env, drawdown state, several pairs, slippage), with
    auto-reset and short episodes so that terminations and restarts happen inside the batches."""
    import os
    from gym_fx_b200.vec_env import VecFxEnv
    cfgd, plugins, kw = VEC_CASES[case]
    N, T, K = 384, 6000, 80
    cfg, candles, minutes = _mk(cfgd, plugins, N, T=T, order_capacity=512, auto_reset=True, episode_bars=70, **kw)
    starts = torch.as_tensor(start_offsets(N, T, 400, 300))
    acts = torch.randint(0, 3, (2, K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(3)).cuda()
    outs = []
    for eng in ("persistent", "graph"):
        os.environ["FXENV_ENGINE"] = eng
        try:
            env = VecFxEnv(cfg, candles, minutes)
        finally:
            del os.environ["FXENV_ENGINE"]
        env.reset(starts)
        ring = torch.zeros((2, N, env.obs_dim), dtype=torch.float32, device="cuda")
        rews = torch.zeros((2, K, N), dtype=torch.float32, device="cuda")
        terms = torch.zeros((2, K, N), dtype=torch.uint8, device="cuda")
        for b in range(2):
            env.step_many(acts[b], ring, rews[b], terms[b])
        torch.cuda.synchronize()
        outs.append((ring, rews, terms, env.get_state()))
        env.close()
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and bytes(a[3]) == bytes(b[3])
    assert int(a[2].sum()) >= N and float(a[1].abs().sum()) > 0.0  # every env ended (and restarted) at least once


@pytest.mark.parametrize("N,K", [(1, 60), (33, 40), (5000, 12)])
def test_step_many_edge_sizes_match_single_steps(N, K):
    """Persistent launch with one env, a ragged warp count, and more envs than resident warps; a one-slot observation
    ring (every step overwrites the same rows): identical to K calls of fxenv_step."""
    import os
    from gym_fx_b200.vec_env import VecFxEnv
    cfgd, plugins, kw = VEC_CASES["cfg2_fixed_fw128_pnl"]
    T = 3000
    cfg, candles, minutes = _mk(cfgd, plugins, N, T=T, order_capacity=256, **kw)
    starts = torch.as_tensor(start_offsets(N, T, K + 10, 300))
    acts = torch.randint(0, 3, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(N)).cuda()
    os.environ["FXENV_ENGINE"] = "persistent"
    try:
        many = VecFxEnv(cfg, candles, minutes)
    finally:
        del os.environ["FXENV_ENGINE"]
    single = VecFxEnv(cfg, candles, minutes)
    many.reset(starts); single.reset(starts)
    assert many.step_many_engine(K) == "persistent"
    ring = torch.zeros((1, N, many.obs_dim), dtype=torch.float32, device="cuda")
    rews = torch.zeros((K, N), dtype=torch.float32, device="cuda")
    terms = torch.zeros((K, N), dtype=torch.uint8, device="cuda")
    many.step_many(acts, ring, rews, terms)
    for k in range(K):
        obs, rew, term, _, _ = single.step(acts[k])
        assert torch.equal(rew, rews[k]) and torch.equal(term.to(torch.uint8), terms[k]), k
    assert torch.equal(obs, ring[0]) and bytes(many.get_state()) == bytes(single.get_state())
    many.close(); single.close()


def test_misaligned_observation_buffers_take_the_general_emitter():
    """The 16-byte-store row emitter needs 16-byte aligned observation rows; a caller's buffer that starts on an odd float
    (a view into a larger allocation) must still work: such rows go through the general emitter.  Rewards, done flags and
    state are identical; the rows agree within the observation tolerance (the fast emitter evaluates a z-score as one
    fma), for single steps and for the persistent batch launch."""
    from gym_fx_b200.vec_env import VecFxEnv
    cfgd, plugins, kw = VEC_CASES["cfg2_fixed_fw128_pnl"]
    N, T, K = 300, 3000, 9
    cfg, candles, minutes = _mk(cfgd, plugins, N, T=T, order_capacity=256, **kw)
    starts = torch.as_tensor(start_offsets(N, T, 100, 300))
    acts = torch.randint(0, 3, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(5)).cuda()
    a, b = VecFxEnv(cfg, candles, minutes), VecFxEnv(cfg, candles, minutes)
    a.reset(starts); b.reset(starts)
    D = a.obs_dim
    odd = torch.zeros(N * D + 1, dtype=torch.float32, device="cuda")[1:].view(N, D)     # base pointer = 4 (mod 16)
    assert odd.data_ptr() % 16 == 4
    for k in range(3):
        oa, ra, ta, _, _ = a.step(acts[k])
        ob, rb, tb, _, _ = b.step(acts[k], out_obs=odd)
        assert torch.equal(ra, rb) and torch.equal(ta, tb)
        np.testing.assert_allclose(odd.cpu().numpy(), oa.cpu().numpy(), rtol=1e-5, atol=2e-6)
    ring_a = torch.zeros((2, N, D), dtype=torch.float32, device="cuda")
    ring_b = torch.zeros(2 * N * D + 1, dtype=torch.float32, device="cuda")[1:].view(2, N, D)
    rews = torch.zeros((2, K, N), dtype=torch.float32, device="cuda")
    terms = torch.zeros((2, K, N), dtype=torch.uint8, device="cuda")
    a.step_many(acts, ring_a, rews[0], terms[0])
    b.step_many(acts, ring_b, rews[1], terms[1])
    torch.cuda.synchronize()
    assert torch.equal(rews[0], rews[1]) and torch.equal(terms[0], terms[1])
    np.testing.assert_allclose(ring_b.cpu().numpy(), ring_a.cpu().numpy(), rtol=1e-5, atol=2e-6)
    assert bytes(a.get_state()) == bytes(b.get_state())
    a.close(); b.close()


@pytest.mark.parametrize("chunk", [1, 3, 7, 64])
def test_step_many_ticket_chunk_lengths_agree(chunk):
    """fx_rollout_kernel hands a warp `chunk` consecutive steps of an env per ticket (fx_rollout_chunk; FXENV_CHUNK forces
    a length).  More envs than resident warps, a batch length that no chunk divides, a 3-slot observation ring, two
    batches back to back (epoch-based sequence words), auto-reset inside the batch: every chunk length gives the results
    of the graph of single steps, bit for bit."""
    import os
    from gym_fx_b200.vec_env import VecFxEnv
    cfgd, plugins, kw = VEC_CASES["cfg2_fixed_fw128_pnl"]
    N, T, K = 5000, 4000, 23
    cfg, candles, minutes = _mk(cfgd, plugins, N, T=T, order_capacity=256, auto_reset=True, episode_bars=40, **kw)
    starts = torch.as_tensor(start_offsets(N, T, 200, 300))
    acts = torch.randint(0, 3, (2, K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(11)).cuda()
    outs = []
    for eng in ("persistent", "graph"):
        os.environ["FXENV_ENGINE"] = eng
        if eng == "persistent":
            os.environ["FXENV_CHUNK"] = str(chunk)
        try:
            env = VecFxEnv(cfg, candles, minutes)
            env.reset(starts)
            ring = torch.zeros((3, N, env.obs_dim), dtype=torch.float32, device="cuda")
            rews = torch.zeros((2, K, N), dtype=torch.float32, device="cuda")
            terms = torch.zeros((2, K, N), dtype=torch.uint8, device="cuda")
            for b in range(2):
                env.step_many(acts[b], ring, rews[b], terms[b])
            torch.cuda.synchronize()
        finally:
            del os.environ["FXENV_ENGINE"]
            os.environ.pop("FXENV_CHUNK", None)
        outs.append((ring, rews, terms, env.get_state()))
        env.close()
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and bytes(a[3]) == bytes(b[3])
    assert int(a[2].sum()) >= N   # every env ended (and restarted) inside the batches


@pytest.mark.parametrize("case,N,pairs_kw", [
    ("cfg3_atr_fw256_dd", 16384, {}),                       # BASELINE configs[2] at full size
    ("cfg5_atr_fw512_sharpe_4pairs", 8192, None),           # BASELINE configs[4] (per GPU) at full size
])
def test_full_size_other_baseline_shapes(case, N, pairs_kw):
    """At BASELINE's full sizes for the two large shapes: a batch through fxenv_step_many is reproducible, independent of how many envs share the launch (every 16th env run alone gives the same
    trajectory), keeps equity under the flat policy, and a sample of 48 envs matches the CPU oracle step by step."""
    from gym_fx_b200.vec_env import VecFxEnv
    cfgd, plugins, kw = VEC_CASES[case]
    T, K = 1 << 15, 48
    starts = start_offsets(N, T, K + 8, 300)
    acts = np.random.default_rng(77).integers(0, 3, (K, N)).astype(np.int32)

    def run(sel, actions):
        cfg, candles, minutes = _mk(cfgd, plugins, len(sel), T=T, order_capacity=256, **kw)
        env = VecFxEnv(cfg, candles, minutes)
        env.reset(torch.as_tensor(starts[sel]))
        ring = torch.zeros((K, len(sel), env.obs_dim), dtype=torch.float32, device="cuda")   # one slot per step
        rews = torch.zeros((K, len(sel)), dtype=torch.float32, device="cuda")
        terms = torch.zeros((K, len(sel)), dtype=torch.uint8, device="cuda")
        env.step_many(torch.as_tensor(np.ascontiguousarray(actions[:, sel])).cuda(), ring, rews, terms)
        torch.cuda.synchronize()
        out = ((ring[K - 1].cpu().numpy(), ring[:, :48].cpu().numpy()), rews.cpu().numpy(), terms.cpu().numpy(),
               {k: env.info()[k].cpu().numpy() for k in ("equity", "cash", "trades", "position", "bar_index", "flags")},
               env.step_many_engine(K))
        env.close()
        return out, (cfg, candles, minutes)

    allenv = np.arange(N)
    ((obs, obs_all_steps), rews, terms, inf, engine), _ = run(allenv, acts)
    assert engine == "persistent" and not np.any(inf["flags"] & 16)
    ((obs2, obs_all_steps2), rews2, terms2, inf2, _), _ = run(allenv, acts)
    assert np.array_equal(obs, obs2) and np.array_equal(rews, rews2) and np.array_equal(inf["equity"], inf2["equity"])
    assert np.array_equal(obs_all_steps, obs_all_steps2)
    # same pair assignment needs the same (global id % pairs): take every 16th env (16 % 4 == 0)
    sel = allenv[::16]
    ((obs_s, _), rews_s, terms_s, inf_s, _), (cfg_s, candles, minutes) = run(sel, acts)
    if cfg_s.num_pairs > 1:
        assert np.all(sel % cfg_s.num_pairs == 0)   # all of them trade pair 0 in both runs only if ids line up
    if cfg_s.num_pairs == 1:
        assert np.array_equal(obs[sel], obs_s) and np.array_equal(rews[:, sel], rews_s)
        for key in ("equity", "cash", "trades", "position"):
            assert np.array_equal(inf[key][sel], inf_s[key]), key
    # sampled oracle at full size (local ids 0..47 keep their pair: id % pairs)
    samp = np.arange(48)
    cfg_o, _, _ = _mk(cfgd, plugins, len(samp), T=T, order_capacity=256, **kw)
    orc = OracleVec(cfg_o, candles, minutes)
    orc.reset(starts[samp])
    for k in range(K):
        oo = orc.step(acts[k, samp])
        np.testing.assert_allclose(rews[k, samp], oo[1], rtol=1e-5, atol=1e-12, err_msg=f"{case} step {k}: reward")
        assert np.array_equal(terms[k, samp], oo[3])
        np.testing.assert_allclose(obs_all_steps[k], oo[0], rtol=1e-5, atol=2e-6, err_msg=f"{case} step {k}: obs of the sampled envs")
    np.testing.assert_allclose(obs[samp], oo[0], rtol=1e-5, atol=2e-6, err_msg=f"{case}: last obs")
    oi = orc.info()
    assert np.array_equal(inf["equity"][samp], oi["equity"]) and np.array_equal(inf["trades"][samp], oi["trades"])
    # flat policy: equity untouched (tools/smoke_test.py:113-118 of the reference)
    (_, rews_f, _, inf_f, _), _ = run(allenv, np.zeros_like(acts))  # noqa: the obs pair is unused here
    assert np.all(inf_f["equity"] == 10000.0) and np.all(inf_f["trades"] == 0) and np.all(rews_f[1:] == 0.0)
