"""The reference-named Python surface, driven the way the reference's own tools drive it:

  * tools/smoke_test.py:108-155 of the reference (flat driver leaves equity unchanged, buy_hold on an uptrend earns,
    a seeded reset reproduces the first observation, total_return == (final - initial) / initial) through
    `app.env.GymFxEnv` / `gym_fx.GymFxEnv` / `gym_fx_b200.GymFxEnv` and the plugin mirrors obtained by
    `app.plugin_loader.load_plugin` (app/main.py:20-24);
  * the goldens `buy_hold_uptrend`, `flat_sample`, `fixed_fw16_sample` (recorded from the UNMODIFIED reference)
    written back to CSV so that data_feed.load_data -> build_table -> fxenv_load_candles runs, compared step by step:
    Dict observation (keys / shapes / dtypes / values), python-float reward, terminated, info keys and values;
  * error behaviour: step before reset -> RuntimeError (app/env.py:132-133), short data -> ValueError (:64-65);
  * VecFxEnv.obs_dict against obs_layout.

CPU part (-m "not gpu"): the import paths, the plugin loader contract and setup.py's entry-point table.
"""
import math
import os

import numpy as np
import pytest

from common import assert_summary_matches, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = {"data_feed": "data_feed.plugins", "broker": "broker.plugins", "strategy": "strategy.plugins",
          "preprocessor": "preprocessor.plugins", "reward": "reward.plugins", "metrics": "metrics.plugins"}


def _instances(config, plugins):
    """app/main.py:20-24: klass(config); set_params(**config) -- through the app.plugin_loader import path."""
    from app.plugin_loader import load_plugin

    out = {}
    for role, name in plugins.items():
        klass, keys = load_plugin(GROUPS[role], name)
        assert list(klass.plugin_params.keys()) == keys
        inst = klass(config)
        inst.set_params(**config)
        out[role] = inst
    return out


def _build_env(config, plugins, env_class=None):
    if env_class is None:
        from app.env import GymFxEnv as env_class
    p = _instances(config, plugins)
    return env_class(config=config, data_feed_plugin=p["data_feed"], broker_plugin=p["broker"],
                     strategy_plugin=p["strategy"], preprocessor_plugin=p["preprocessor"], reward_plugin=p["reward"],
                     metrics_plugin=p["metrics"])


def _golden_to_csv(g, path):
    from gym_fx_b200.synth import write_csv

    write_csv(path, g["candles"], g["minutes"])


def _run_driver(config, plugins, seed=None):
    """tools/smoke_test.py:70-86 of the reference."""
    env = _build_env(config, plugins)
    strategy = env.strategy_plugin
    obs, info = env.reset(seed=seed)
    first = {k: np.array(v, copy=True) for k, v in obs.items()}
    step, done = 0, False
    while not done and step < config["steps"]:
        action = strategy.decide_action(obs=obs, info=info, step=step)
        obs, _, terminated, truncated, info = env.step(action)
        done = terminated or truncated
        step += 1
    summary = env.summary()
    env.close()
    return summary, first, env


# ---------------------------------------------------------------------------------------------------- CPU
def test_import_paths_resolve_to_one_class():
    import app.env
    import gym_fx
    import gym_fx.env
    import gym_fx_b200

    assert app.env.GymFxEnv is gym_fx_b200.GymFxEnv is gym_fx.GymFxEnv is gym_fx.env.GymFxEnv
    from app.config import DEFAULT_VALUES
    assert DEFAULT_VALUES["window_size"] == 32 and DEFAULT_VALUES["initial_cash"] == 10000.0
    assert DEFAULT_VALUES["reward_plugin"] == "pnl_reward" and DEFAULT_VALUES["price_column"] == "CLOSE"


def test_plugin_loader_contract_and_setup_entry_points():
    """Every (group, name) of the reference's setup.py:11-35 resolves, and setup.py declares exactly those."""
    from app.plugin_loader import get_plugin_params, load_plugin

    want = {
        "data_feed.plugins": ["default_data_feed"],
        "broker.plugins": ["default_broker", "oanda_broker"],
        "strategy.plugins": ["default_strategy", "direct_fixed_sltp", "direct_atr_sltp"],
        "preprocessor.plugins": ["default_preprocessor", "feature_window_preprocessor"],
        "reward.plugins": ["pnl_reward", "sharpe_reward", "dd_penalized_reward"],
        "metrics.plugins": ["default_metrics"],
    }
    for group, names in want.items():
        for name in names:
            klass, keys = load_plugin(group, name)
            assert klass.__name__ == "Plugin" and isinstance(klass.plugin_params, dict)
            assert keys == list(klass.plugin_params.keys()) == list(get_plugin_params(group, name).keys())
            for meth in ("set_params",):
                assert callable(getattr(klass, meth))
    with pytest.raises(ImportError):
        load_plugin("reward.plugins", "no_such_reward")
    # setup.py declares the same table (run with setuptools.setup captured)
    import runpy
    import setuptools

    captured = {}
    real = setuptools.setup
    setuptools.setup = lambda **kw: captured.update(kw)
    try:
        runpy.run_path(os.path.join(ROOT, "setup.py"))
    finally:
        setuptools.setup = real
    eps = captured["entry_points"]
    assert set(eps) == set(want)
    for group, names in want.items():
        assert [e.split("=")[0] for e in eps[group]] == names, group
        for e in eps[group]:
            mod, _, attr = e.split("=")[1].partition(":")
            assert attr == "Plugin" and __import__("importlib").import_module(mod).Plugin is load_plugin(group, e.split("=")[0])[0]
    assert {"gym_fx_b200", "gym_fx", "app"} <= set(captured["packages"])


def test_load_data_and_build_table_roundtrip(tmp_path):
    """default_data_feed.load_data (reference :36-56) -> build_table: the CSV of a golden gives back its candle table."""
    from gym_fx_b200.data_feed_plugins.default_data_feed import Plugin as Feed

    g = load_golden("fixed_fw16_sample")
    path = str(tmp_path / "d.csv")
    _golden_to_csv(g, path)
    cfg = {"input_data_file": path, "date_column": "DATE_TIME", "price_column": "CLOSE", "headers": True, "max_rows": None}
    feed = Feed(cfg)
    df = feed.load_data(cfg)
    assert list(df.columns[:5]) == ["OPEN", "HIGH", "LOW", "CLOSE", "VOLUME"] and len(df) == g["candles"].shape[0]
    table, cols, minutes = Feed.build_table(df, extra_columns=["CLOSE"])
    assert cols == ["OPEN", "HIGH", "LOW", "CLOSE", "VOLUME"]
    np.testing.assert_array_equal(table, g["candles"])
    np.testing.assert_array_equal(minutes, g["minutes"])
    # a price-only file: OHLC filled from the price column, VOLUME 0 (reference :48-55)
    p2 = str(tmp_path / "p.csv")
    with open(p2, "w") as fh:
        fh.write("DATE_TIME,CLOSE\n2024-01-01 00:00:00,1.1\n2024-01-01 00:01:00,1.2\nnot a date,1.3\n")
    df2 = feed.load_data({**cfg, "input_data_file": p2})
    assert len(df2) == 2 and (df2["OPEN"] == df2["CLOSE"]).all() and (df2["VOLUME"] == 0).all()
    with pytest.raises(ValueError):
        feed.load_data({**cfg, "input_data_file": p2, "price_column": "MID"})


# ---------------------------------------------------------------------------------------------------- GPU
def _replay_golden(name, tmp_path, env_class=None):
    g = load_golden(name)
    meta = g["meta"]
    path = str(tmp_path / f"{name}.csv")
    _golden_to_csv(g, path)
    config = {**meta["config"], "input_data_file": path}
    env = _build_env(config, meta["plugins"], env_class)
    assert env.total_bars == g["candles"].shape[0] and len(env.dataframe) == env.total_bars
    layout_keys = list(env.observation_space.spaces.keys())
    rows = {int(r): i for i, r in enumerate(g["obs_rows"])}

    def check_obs(row, obs):
        assert list(obs.keys()) == layout_keys
        flat = []
        for k in layout_keys:
            v = obs[k]
            assert isinstance(v, np.ndarray) and v.dtype == np.float32, (k, type(v))
            assert v.shape == env.observation_space.spaces[k].shape, (k, v.shape)
            flat.append(v.reshape(-1))
        if row in rows:
            np.testing.assert_allclose(np.concatenate(flat), g["obs"][rows[row]], rtol=1e-5, atol=2e-6,
                                       err_msg=f"{name}: obs row {row}")

    def check_info(row, info, stepped):
        keys = {"equity", "position", "price", "bar_index", "total_bars", "trades", "commission_paid"}
        assert keys <= set(info), info.keys()
        if stepped:
            assert {"reward", "pnl", "trade_cost"} <= set(info) and info["trade_cost"] == 0.0
        assert isinstance(info["equity"], float) and isinstance(info["position"], int)
        assert info["equity"] == g["equity"][row], (row, info["equity"], g["equity"][row])      # fp64, bit-exact
        assert info["position"] == int(g["position"][row]) and info["bar_index"] == int(g["bar_index"][row])
        assert info["price"] == g["price"][row] and info["trades"] == int(g["trades"][row])
        assert info["commission_paid"] == g["commission_paid"][row] and info["total_bars"] == env.total_bars

    obs, info = env.reset(seed=7)
    check_obs(0, obs)
    check_info(0, info, False)
    n = g["reward"].shape[0]
    terminated_before = False
    for k in range(n - 1):
        a = g["actions"][k]
        obs, reward, terminated, truncated, info = env.step(a if g["actions"].dtype.kind != "f" else np.array([a], np.float32))
        assert isinstance(reward, float) and isinstance(terminated, bool) and truncated is False
        assert math.isclose(reward, float(g["reward"][k + 1]), rel_tol=1e-9, abs_tol=1e-13), (k, reward, g["reward"][k + 1])
        assert terminated == bool(g["terminated"][k + 1]), k
        check_obs(k + 1, obs)
        check_info(k + 1, info, stepped=not terminated_before)
        terminated_before = terminated
    summary = env.summary()
    assert summary["final_equity"] == g["equity"][n - 1]
    if not terminated_before:   # run still going: the reference's analyzers are not visible yet
        assert summary["max_drawdown_pct"] is None and summary["trades_total"] == 0 and summary["avg_trade_pnl"] is None
    env.close()
    after = env.summary()        # after close() cerebro.run has returned: analyzers visible, bridge.equity kept
    assert after["final_equity"] == g["equity"][n - 1]
    assert_summary_matches(after, meta["summary"], label=name)
    if terminated_before:
        assert_summary_matches(summary, meta["summary"], label=name + " (terminated, before close)")
    return summary, g


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["buy_hold_uptrend", "flat_sample", "fixed_fw16_sample"])
def test_gym_env_replays_reference_golden_from_csv(name, tmp_path):
    summary, g = _replay_golden(name, tmp_path)
    if name == "buy_hold_uptrend":   # the reference's own known answer (examples/results/buy_hold_summary.json:3-4)
        assert summary["final_equity"] == 10000.095791583166 and summary["total_return"] == 9.579158316563863e-06
    if name == "flat_sample":        # examples/results/flat_summary.json:3-4
        assert summary["final_equity"] == 10000.0 and summary["total_return"] == 0.0


@pytest.mark.gpu
def test_gym_fx_alias_class_drives_the_same_path(tmp_path):
    from gym_fx import GymFxEnv
    _replay_golden("fixed_fw16_sample", tmp_path, env_class=GymFxEnv)


@pytest.mark.gpu
def test_reference_smoke_test_assertions(tmp_path):
    """tools/smoke_test.py:108-155 of the reference, same drivers, same four assertions."""
    from app.config import DEFAULT_VALUES

    plugins = dict(data_feed="default_data_feed", broker="default_broker", strategy="default_strategy",
                   preprocessor="default_preprocessor", reward="pnl_reward", metrics="default_metrics")
    sample, up = str(tmp_path / "sample.csv"), str(tmp_path / "uptrend.csv")
    _golden_to_csv(load_golden("flat_sample"), sample)
    _golden_to_csv(load_golden("buy_hold_uptrend"), up)

    def base(driver_mode, data):
        return {**DEFAULT_VALUES, "mode": "inference", "driver_mode": driver_mode, "steps": 480, "input_data_file": data,
                "date_column": "DATE_TIME", "price_column": "CLOSE", "headers": True, "window_size": 32,
                "initial_cash": 10000.0, "position_size": 1.0, "commission": 0.0, "slippage": 0.0}

    flat, _, _ = _run_driver(base("flat", sample), plugins)
    assert math.isclose(flat["final_equity"], flat["initial_cash"], rel_tol=1e-9, abs_tol=1e-3)
    assert math.isclose(flat["total_return"], 0.0, abs_tol=1e-6)
    upsum, _, _ = _run_driver(base("buy_hold", up), plugins, seed=42)
    assert upsum["total_return"] > 0.0
    _, a, _ = _run_driver(base("flat", sample), plugins, seed=123)
    _, b, _ = _run_driver(base("flat", sample), plugins, seed=123)
    for key in a:
        assert np.allclose(a[key], b[key]), key
    expected = (upsum["final_equity"] - upsum["initial_cash"]) / upsum["initial_cash"]
    assert math.isclose(upsum["total_return"], expected, rel_tol=1e-9, abs_tol=1e-9)
    # the summary keys of metrics_plugins/default_metrics.py:48-60
    assert {"initial_cash", "final_equity", "total_return", "max_drawdown_pct", "max_drawdown_money", "sharpe_ratio",
            "sqn", "trades_total", "trades_won", "trades_lost", "avg_trade_pnl"} <= set(upsum)


@pytest.mark.gpu
def test_gym_env_errors_and_spaces(tmp_path):
    from app.config import DEFAULT_VALUES
    from gym_fx_b200 import spaces

    plugins = dict(data_feed="default_data_feed", broker="default_broker", strategy="default_strategy",
                   preprocessor="default_preprocessor", reward="pnl_reward", metrics="default_metrics")
    sample = str(tmp_path / "sample.csv")
    _golden_to_csv(load_golden("flat_sample"), sample)
    cfg = {**DEFAULT_VALUES, "input_data_file": sample, "window_size": 32}
    env = _build_env(cfg, plugins)
    with pytest.raises(RuntimeError, match="reset"):
        env.step(0)
    assert isinstance(env.action_space, spaces.Discrete) and env.action_space.n == 3
    obs, info = env.reset()
    for k, box in env.observation_space.spaces.items():
        assert obs[k].shape == box.shape and obs[k].dtype == np.float32
    assert set(obs) == {"prices", "returns", "position", "equity_norm", "unrealized_pnl_norm", "steps_remaining_norm"}
    # malformed actions are coerced to hold (app/env.py:187-204): equity stays put
    for bad in (7, -1, "x", None, 3.9):
        _, r, term, trunc, info = env.step(bad)
        assert r == 0.0 and not term and info["position"] == 0
    assert env.render() is None
    env.close()
    short = str(tmp_path / "short.csv")
    with open(sample) as fh, open(short, "w") as out:
        out.writelines(fh.readlines()[:20])
    with pytest.raises(ValueError, match="too short"):
        _build_env({**cfg, "input_data_file": short}, plugins)
    with pytest.raises(ValueError, match="price_column"):
        _build_env({**cfg, "price_column": "MID"}, plugins)
    cont = _build_env({**cfg, "action_space_mode": "continuous"}, plugins)
    assert isinstance(cont.action_space, spaces.Box) and cont.action_space.shape == (1,)
    cont.reset()
    _, _, _, _, info = cont.step(np.array([0.9], np.float32))
    cont.close()


@pytest.mark.gpu
def test_vec_obs_dict_matches_layout():
    import torch

    from common import config_from_meta
    from gym_fx_b200.config import obs_layout
    from gym_fx_b200.vec_env import VecFxEnv

    g = load_golden("fixed_fw16_sample")
    cfg = config_from_meta(g["meta"], num_envs=3)
    env = VecFxEnv(cfg, [g["candles"]], [g["minutes"]])
    obs, _ = env.reset(torch.zeros(3, dtype=torch.int64))
    d = env.obs_dict()
    lay = obs_layout(cfg)
    assert list(d.keys()) == list(lay.keys())
    total = 0
    for k, (off, shape) in lay.items():
        assert tuple(d[k].shape) == (3,) + tuple(shape)
        n = int(np.prod(shape))
        assert torch.equal(d[k].reshape(3, -1), obs[:, off:off + n])
        total += n
    assert total == env.obs_dim
    np.testing.assert_allclose(obs[0].cpu().numpy(), g["obs"][0], rtol=1e-5, atol=2e-6)
    # argument checks of the raw-pointer calls: wrong dtype / shape / device raise instead of corrupting
    K, N, D = 4, 3, env.obs_dim
    ring = torch.empty((2, N, D), dtype=torch.float32, device="cuda")
    rews = torch.empty((K, N), dtype=torch.float32, device="cuda")
    terms = torch.empty((K, N), dtype=torch.uint8, device="cuda")
    good = torch.zeros((K, N), dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        env.step_many(torch.zeros((K, N), dtype=torch.int64, device="cuda"), ring, rews, terms)   # torch.randint's default
    with pytest.raises(ValueError):
        env.step_many(good.cpu(), ring, rews, terms)
    with pytest.raises(ValueError):
        env.step_many(good, ring[:, :2], rews, terms)
    with pytest.raises(ValueError):
        env.step_many(good, ring, rews[:2], terms)
    with pytest.raises(ValueError):
        env.reset(torch.zeros(2, dtype=torch.int64))
    env.step_many(good, ring, rews, terms)
    torch.cuda.synchronize()
    env.close()
