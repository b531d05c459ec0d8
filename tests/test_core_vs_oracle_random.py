"""CPU (-m "not gpu"): randomised differential test of the PRODUCT's scalar core (fx_core.cuh via tests/hostsim: FIFO
table of entries, brackets folded into PAIR entries, cash-bound check_submitted) against the C oracle (flat list of
individual backtrader-style orders) on random configurations the goldens do not enumerate: commission x leverage x
slippage x position size x strategy x reward x sizing mode, random candles and action streams.  Account state and the
integer state must agree bit for bit at every step, rewards to 1e-12."""
import numpy as np
import pytest

import scenarios as S
from gym_fx_b200.config import lower_config
from gym_fx_b200.synth import synth_candles, synth_minutes
from hostsim.hostsim import HostSimEnv
from oracle.c_oracle import OracleVec


def _random_case(seed):
    rng = np.random.default_rng(seed)
    strat = ["default_strategy", "direct_fixed_sltp", "direct_atr_sltp"][seed % 3]
    reward = ["pnl_reward", "sharpe_reward", "dd_penalized_reward"][(seed // 3) % 3]
    cfgd = dict(S.DEFAULTS)
    cfgd.update(window_size=int(rng.integers(4, 24)),
                commission=float(rng.choice([0.0, 1e-5, 7e-5])),
                leverage=float(rng.choice([1.0, 2.0, 30.0])),
                slippage_perc=float(rng.choice([0.0, 0.0, 5e-5, 3e-4])),
                position_size=float(rng.choice([1.0, 250.0, 4000.0, 9000.0])),   # the large ones hit margin rejections
                sl_pips=float(rng.choice([2.0, 5.0, 20.0])), tp_pips=float(rng.choice([3.0, 8.0, 40.0])),
                atr_period=int(rng.integers(3, 15)), k_sl=float(rng.choice([0.5, 2.0])), k_tp=float(rng.choice([1.0, 3.0])))
    if strat == "direct_atr_sltp" and rng.random() < 0.5:
        cfgd.update(rel_volume=float(rng.choice([0.05, 0.4])), max_order_volume=20000.0,
                    size_mode=str(rng.choice(["fx_units", "notional"])))
    if reward == "sharpe_reward":
        cfgd.update(window=int(rng.integers(4, 40)))
    pair = int(rng.integers(0, 4))
    T = 700
    pl = S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, "strategy": strat, "reward": reward})
    cfg = lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"], preprocessor_plugin=pl["preprocessor"],
                       reward_plugin=pl["reward"], columns=S.OHLCV, num_envs=1, order_capacity=1024,
                       pair_pip_size=[0.01 if pair == 3 else 1e-4])
    candles = synth_candles(T, pair, 5000 + seed)
    hold = float(rng.choice([0.2, 0.6, 0.85]))
    acts = S.make_actions(("sticky", 900 + seed, hold), T - 40)
    return cfg, candles, synth_minutes(T), acts


@pytest.mark.parametrize("seed", range(36))
def test_product_core_matches_oracle_on_random_configs(seed):
    cfg, candles, minutes, acts = _random_case(seed)
    core = HostSimEnv(cfg, candles, minutes)
    orc = OracleVec(cfg, [candles], [minutes])
    core.reset(0)
    orc.reset(np.zeros(1, np.int64))
    for k, a in enumerate(acts):
        r, t = core.step(a)
        _, _, r64, term = orc.step(np.array([a], np.int32))
        assert t == term[0], (seed, k)
        np.testing.assert_allclose(r, r64[0], rtol=1e-12, atol=1e-15, err_msg=f"seed {seed} step {k}: reward")
        ci, oi = core.info(), orc.info()
        assert not (ci["flags"] & 16), f"seed {seed}: order table overflow at step {k}"
        for key in ("position", "bar_index", "trades"):
            assert ci[key] == oi[key][0], (seed, k, key, ci[key], oi[key][0])
        for key in ("equity", "cash", "position_size", "position_price", "commission_paid", "price"):
            assert ci[key] == oi[key][0], (seed, k, key, repr(ci[key]), repr(oi[key][0]))
        if t:
            break
    # end-of-run statistics of the product core (FX_RS_* record) against the oracle's analyzers
    rs, osum = core.stats(), {k: v[0] for k, v in orc.summary().items()}
    closed = core.info()["trades"]
    assert rs[2] == osum["max_drawdown_pct"] and rs[1] == osum["max_drawdown_money"], (seed, rs[:3], osum)
    assert (rs[9], rs[10], rs[11], closed) == (osum["trades_total"], osum["trades_won"], osum["trades_lost"], osum["trades_closed"])
    if closed:
        assert rs[6] / closed == osum["avg_trade_pnl"], (seed, rs[6] / closed, osum["avg_trade_pnl"])
    if closed > 1 and osum["sqn"] == osum["sqn"]:
        mean = rs[6] / closed
        sqn = np.sqrt(closed) * mean / np.sqrt(rs[7] / closed - mean * mean)   # from the sums of pnl and pnl^2
        np.testing.assert_allclose(sqn, osum["sqn"], rtol=1e-9, err_msg=f"seed {seed}: sqn")


def test_trade_statistics_with_dust_positions():
    """rel_volume sizing gives non-round order sizes: closing a position leaves 1e-12 'dust', whose own close() order can
    execute after the position has flipped and be absorbed by rounding -- backtrader's Trade.update then books a pnl
    instead of averaging the price (|size after| > |size before| is false).  192 envs, product core vs oracle analyzers."""
    from gym_fx_b200.synth import start_offsets
    cfgd = {**S.DEFAULTS, "window_size": 16, "commission": 2e-5, "leverage": 5.0, "rel_volume": 0.3, "max_order_volume": 9000.0}
    pl = S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, "strategy": "direct_atr_sltp"})
    N, T, steps = 192, 6000, 330
    mk = lambda n: lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"], preprocessor_plugin=pl["preprocessor"],
                                reward_plugin=pl["reward"], columns=S.OHLCV, num_envs=n, order_capacity=512)
    candles, minutes = synth_candles(T, 0), synth_minutes(T)
    starts = start_offsets(N, T, steps + 10, 300)
    orc = OracleVec(mk(N), [candles], [minutes])
    orc.reset(starts)
    cores = [HostSimEnv(mk(1), candles, minutes) for _ in range(N)]
    for i, h in enumerate(cores):
        h.reset(int(starts[i]))
    rng = np.random.default_rng(2024)
    dust = 0
    for k in range(steps):
        a = rng.integers(0, 3, N).astype(np.int32)
        orc.step(a, want_obs=False)
        for i, h in enumerate(cores):
            h.step(int(a[i]))
        if k % 4 == 0:
            ps = np.abs(orc.info()["position_size"])
            dust += int(np.count_nonzero((ps > 0.0) & (ps < 1e-6)))
    osum = orc.summary()
    for i, h in enumerate(cores):
        rs, closed = h.stats(), h.info()["trades"]
        assert (rs[9], rs[10], rs[11], closed) == (osum["trades_total"][i], osum["trades_won"][i], osum["trades_lost"][i], osum["trades_closed"][i]), i
        assert rs[2] == osum["max_drawdown_pct"][i] and rs[1] == osum["max_drawdown_money"][i], i
        if closed:
            assert rs[6] / closed == osum["avg_trade_pnl"][i], (i, rs[6] / closed, osum["avg_trade_pnl"][i])
    assert dust > 0, "the scenario no longer produces dust positions"
