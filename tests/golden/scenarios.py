"""
Scenario table shared by tests/golden/make_golden.py (which runs the UNMODIFIED reference over oracle/bt_shim
in the build container) and by the parity tests (which replay the same inputs through the C oracle and the CUDA
env and diff against the stored trajectories).

Every scenario = data + config overrides + plugin names + a seeded action stream.  `DEFAULTS` restates
app/config.py:1-45 of the reference (only the keys the hot path reads) so replays do not need the reference tree.
"""
from __future__ import annotations

import numpy as np

DEFAULTS = {
    "window_size": 32, "initial_cash": 10000.0, "position_size": 1.0, "commission": 0.0, "slippage": 0.0,
    "price_column": "CLOSE", "date_column": "DATE_TIME", "headers": True, "max_rows": None,
}

DEFAULT_PLUGINS = dict(data_feed="default_data_feed", broker="default_broker", strategy="default_strategy",
                       preprocessor="default_preprocessor", reward="pnl_reward", metrics="default_metrics")

OHLCV = ["OPEN", "HIGH", "LOW", "CLOSE", "VOLUME"]


def _pl(**kw):
    d = dict(DEFAULT_PLUGINS)
    d.update(kw)
    return d


def _fw(W, S=256, **kw):
    d = {"window_size": W, "feature_columns": list(OHLCV), "feature_scaling_window": S}
    d.update(kw)
    return d


# data: ("fixture", name) -> reference examples/data/<name>.csv ; ("synth", T, pair, seed) -> gym_fx_b200.synth ;
#       ("synth_extra", T, seed) -> synth + two extra feature columns (FEAT_A gaussian, BIN_FLAG 0/1)
# actions: ("const", first, rest) | ("random", seed) | ("random_cont", seed) | ("sticky", seed, p_hold)
SCENARIOS = [
    dict(name="buy_hold_uptrend", data=("fixture", "eurusd_uptrend"), cfg={}, plugins=_pl(),
         actions=("const", 1, 0), steps=480),
    dict(name="flat_sample", data=("fixture", "eurusd_sample"), cfg={}, plugins=_pl(),
         actions=("const", 0, 0), steps=480),
    dict(name="default_random_exhaust", data=("fixture", "eurusd_sample"), cfg={}, plugins=_pl(),
         actions=("random", 11), steps=520, after_done=3),
    dict(name="default_sticky_commission", data=("synth", 700, 0, 5), cfg={"commission": 2e-5, "leverage": 20.0},
         plugins=_pl(), actions=("sticky", 12, 0.8), steps=650),
    dict(name="fixed_fw128_pnl", data=("synth", 640, 0, 1000),
         cfg=_fw(128), plugins=_pl(strategy="direct_fixed_sltp", preprocessor="feature_window_preprocessor"),
         actions=("random", 1234), steps=600, obs_every=7),
    dict(name="fixed_fw16_sample", data=("fixture", "eurusd_sample"),
         cfg=_fw(16, 32, sl_pips=3.0, tp_pips=5.0),
         plugins=_pl(strategy="direct_fixed_sltp", preprocessor="feature_window_preprocessor"),
         actions=("sticky", 77, 0.7), steps=505, after_done=2),
    dict(name="fixed_same_bar_children", data=("synth", 500, 0, 21),
         cfg=_fw(8, 16, sl_pips=2.0, tp_pips=3.0),
         plugins=_pl(strategy="direct_fixed_sltp", preprocessor="feature_window_preprocessor"),
         actions=("sticky", 78, 0.6), steps=450, children_same_bar=True),
    dict(name="fixed_margin_commission", data=("synth", 600, 0, 33),
         cfg={"position_size": 6000.0, "commission": 5e-5, "sl_pips": 4.0, "tp_pips": 6.0},
         plugins=_pl(strategy="direct_fixed_sltp", reward="dd_penalized_reward"),
         actions=("random", 5), steps=560),
    dict(name="atr_fw32_dd", data=("synth", 800, 0, 1001),
         cfg=_fw(32, 64), plugins=_pl(strategy="direct_atr_sltp", preprocessor="feature_window_preprocessor",
                                      reward="dd_penalized_reward"),
         actions=("random", 4321), steps=760),
    dict(name="atr_relvol_lev_sharpe", data=("synth", 700, 0, 44),
         cfg={"rel_volume": 0.5, "leverage": 3.0, "commission": 1e-5, "min_order_volume": 10.0,
              "max_order_volume": 12000.0, "k_sl": 1.0, "k_tp": 1.5, "min_sltp_frac": None, "window": 16},
         plugins=_pl(strategy="direct_atr_sltp", reward="sharpe_reward"),
         actions=("sticky", 9, 0.5), steps=660),
    dict(name="atr_notional_jpy", data=("synth", 600, 3, 1003),
         cfg={"rel_volume": 0.2, "size_mode": "notional", "leverage": 50.0, "atr_period": 7},
         plugins=_pl(strategy="direct_atr_sltp", reward="pnl_reward"),
         actions=("random", 10), steps=560),
    dict(name="atr_session_filter", data=("synth", 500, 0, 55),
         cfg={"session_filter": True, "entry_dow_start": 0, "entry_hour_start": 1, "force_close_dow": 0,
              "force_close_hour": 5, "atr_period": 5},
         plugins=_pl(strategy="direct_atr_sltp"), actions=("random", 6), steps=470),
    dict(name="default_broke", data=("synth", 600, 0, 66),
         cfg={"position_size": 2.0e6, "leverage": 400.0},
         plugins=_pl(), actions=("sticky", 3, 0.9), steps=560, after_done=2),
    dict(name="continuous_actions", data=("synth", 400, 0, 67), cfg={"action_space_mode": "continuous"},
         plugins=_pl(strategy="direct_fixed_sltp"), actions=("random_cont", 8), steps=360),
    dict(name="sharpe_fixed_sample", data=("fixture", "eurusd_sample"), cfg={"sl_pips": 5.0, "tp_pips": 8.0},
         plugins=_pl(strategy="direct_fixed_sltp", reward="sharpe_reward"),
         actions=("sticky", 15, 0.6), steps=480),
    dict(name="fw_expanding_extra_cols", data=("synth_extra", 420, 68),
         cfg={"window_size": 12, "feature_columns": ["CLOSE", "FEAT_A", "BIN_FLAG", "VOLUME"],
              "feature_binary_columns": ["BIN_FLAG"], "feature_scaling": "expanding_zscore", "feature_clip": 2.5},
         plugins=_pl(preprocessor="feature_window_preprocessor"), actions=("random", 2), steps=380),
    dict(name="fw_none_noprice_noagent", data=("synth_extra", 300, 69),
         cfg={"window_size": 9, "feature_columns": ["FEAT_A", "HIGH", "BIN_FLAG"], "feature_scaling": "none",
              "include_price_window": False, "include_agent_state": False, "feature_clip": 0.0},
         plugins=_pl(preprocessor="feature_window_preprocessor"), actions=("random", 3), steps=260),
    dict(name="fw_rolling_noprice", data=("synth", 300, 1, 70),
         cfg=_fw(10, 20, include_price_window=False),
         plugins=_pl(preprocessor="feature_window_preprocessor", strategy="direct_fixed_sltp"),
         actions=("sticky", 4, 0.7), steps=260),
    dict(name="default_price_open", data=("synth", 300, 0, 71), cfg={"price_column": "OPEN", "window_size": 8},
         plugins=_pl(), actions=("sticky", 5, 0.8), steps=260),
    # slippage (default_broker.py:39-51: set_slippage_perc(perc, slip_open=True, slip_limit=True, slip_match=True))
    dict(name="slip_default_sample", data=("fixture", "eurusd_sample"),     # market orders; OPEN often outside [LOW, HIGH]
         cfg={"slippage_perc": 2e-4, "commission": 1e-5}, plugins=_pl(), actions=("sticky", 21, 0.5), steps=480),
    dict(name="slip_fixed_brackets", data=("synth", 600, 0, 72),           # limit parents, stop / limit children
         cfg={"slippage": 1.5e-4, "sl_pips": 4.0, "tp_pips": 6.0}, plugins=_pl(strategy="direct_fixed_sltp"),
         actions=("random", 22), steps=560),
    dict(name="slip_atr_lev_sharpe", data=("synth", 600, 1, 73),
         cfg={"slippage_perc": 5e-5, "leverage": 4.0, "commission": 2e-5, "atr_period": 6},
         plugins=_pl(strategy="direct_atr_sltp", reward="sharpe_reward"), actions=("sticky", 23, 0.6), steps=560),
]


def make_actions(spec, steps):
    kind = spec[0]
    if kind == "const":
        a = np.full(steps, spec[2], np.int32)
        a[0] = spec[1]
        return a
    if kind == "random":
        return np.random.default_rng(spec[1]).integers(0, 3, steps).astype(np.int32)
    if kind == "random_cont":
        return np.random.default_rng(spec[1]).uniform(-1.0, 1.0, steps).astype(np.float32)
    if kind == "sticky":  # repeat the previous action with probability p (long-lived positions)
        rng = np.random.default_rng(spec[1])
        a = np.zeros(steps, np.int32)
        cur = 0
        for i in range(steps):
            if rng.random() > spec[2]:
                cur = int(rng.integers(0, 3))
            a[i] = cur
        return a
    raise ValueError(kind)


def make_data(spec):
    """-> (float64 [T, C] table, columns, int64 [T] minutes)."""
    from gym_fx_b200.synth import synth_candles, synth_minutes

    kind = spec[0]
    if kind == "synth":
        _, T, pair, seed = spec
        return synth_candles(T, pair, seed), list(OHLCV), synth_minutes(T)
    if kind == "synth_extra":
        _, T, seed = spec
        c = synth_candles(T, 0, seed)
        rng = np.random.default_rng(seed + 7)
        extra = np.stack([rng.normal(0.0, 1.0, T).round(6), rng.integers(0, 2, T).astype(np.float64)], axis=1)
        return np.ascontiguousarray(np.concatenate([c, extra], axis=1)), list(OHLCV) + ["FEAT_A", "BIN_FLAG"], \
            synth_minutes(T)
    raise ValueError(kind)


def full_config(sc):
    cfg = dict(DEFAULTS)
    cfg.update(sc["cfg"])
    return cfg


def build_mirror_plugins(cfg, plugins):
    """Instantiate this package's plugin mirrors the way app/main.py:20-24 instantiates the reference's."""
    from gym_fx_b200.plugin_loader import build_plugins

    return build_plugins(cfg, plugins)
