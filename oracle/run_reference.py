"""
oracle/run_reference.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the UNMODIFIED reference (/root/reference: app/env.py, app/bt_bridge.py and
its plugins) on top of the backtrader/gymnasium shims in oracle/bt_shim, driven by
a replayed action stream, and records the full per-step trajectory.  This only
works in the build container (the GPU box has no /root/reference); its outputs
are committed as fixtures under tests/golden/ by tests/golden/make_golden.py.
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("GYMFX_REFERENCE_ROOT", "/root/reference")

PLUGIN_MODULES = {
    "data_feed": "data_feed_plugins.{}",
    "broker": "broker_plugins.{}",
    "strategy": "strategy_plugins.{}",
    "preprocessor": "preprocessor_plugins.{}",
    "reward": "reward_plugins.{}",
    "metrics": "metrics_plugins.{}",
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "app", "env.py"))


def _install_paths() -> None:
    shim = os.path.join(_HERE, "bt_shim")
    for p in (REFERENCE_ROOT, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    # make sure a product-side `app`/`gym_fx` package of the same name is not shadowing the reference
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in ("app", "gym_fx") or root.endswith("_plugins"):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(REFERENCE_ROOT):
                del sys.modules[name]


def build_reference_env(config: Dict[str, Any], plugins: Dict[str, str], children_same_bar: bool = False):
    """Instantiate the reference GymFxEnv exactly like app/main.py:20-55 does."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_paths()
    import backtrader as bt  # the shim

    assert "shim" in bt.__version__, "real backtrader unexpectedly importable; shim not in effect"
    insts = {}
    for group, name in plugins.items():
        mod = importlib.import_module(PLUGIN_MODULES[group].format(name))
        inst = mod.Plugin(config)
        inst.set_params(**config)
        insts[group] = inst
    if children_same_bar:
        orig = insts["broker"].build_bt_broker

        def build(cfg, _orig=orig):
            b = _orig(cfg)
            b.children_same_bar = True
            return b

        insts["broker"].build_bt_broker = build
    env_mod = importlib.import_module("app.env")
    env = env_mod.GymFxEnv(
        config=config,
        data_feed_plugin=insts["data_feed"],
        broker_plugin=insts["broker"],
        strategy_plugin=insts["strategy"],
        preprocessor_plugin=insts["preprocessor"],
        reward_plugin=insts["reward"],
        metrics_plugin=insts["metrics"],
    )
    return env


OBS_ORDER = ("features", "prices", "returns", "position", "equity_norm", "unrealized_pnl_norm", "steps_remaining_norm")


def flatten_obs(obs: Dict[str, np.ndarray]) -> np.ndarray:
    """Flat VecEnv layout: features row-major | prices | returns | 4 scalars."""
    parts = [np.asarray(obs[k], dtype=np.float32).reshape(-1) for k in OBS_ORDER if k in obs]
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


def run_reference(
    config: Dict[str, Any],
    plugins: Dict[str, str],
    actions: Iterable,
    children_same_bar: bool = False,
    extra_steps_after_done: int = 0,
) -> Dict[str, np.ndarray]:
    """reset() then step() through `actions`; keeps stepping `extra_steps_after_done`
    times after termination (the reference keeps answering with reward 0)."""
    env = build_reference_env(config, plugins, children_same_bar)
    obs, info = env.reset(seed=config.get("seed"))
    rec: Dict[str, List] = {k: [] for k in (
        "obs", "reward", "terminated", "equity", "position", "price", "bar_index", "trades", "commission_paid")}

    def push(o, r, term, inf):
        rec["obs"].append(flatten_obs(o))
        rec["reward"].append(float(r))
        rec["terminated"].append(bool(term))
        rec["equity"].append(float(inf["equity"]))
        rec["position"].append(int(inf["position"]))
        rec["price"].append(float(inf["price"]))
        rec["bar_index"].append(int(inf["bar_index"]))
        rec["trades"].append(int(inf["trades"]))
        rec["commission_paid"].append(float(inf["commission_paid"]))

    push(obs, 0.0, False, info)  # row 0 = reset()
    after = 0
    for a in actions:
        obs, r, term, trunc, info = env.step(a)
        push(obs, r, term, info)
        if term:
            after += 1
            if after > extra_steps_after_done:
                break
    total_bars = env.total_bars
    env.close()
    # After close() the worker thread has returned from cerebro.run(), so summary() now sees the analyzers
    # (app/env.py:256-271; before close() they are empty): the end-of-run metrics oracle.
    summary = env.summary()
    out = {
        "obs": np.stack(rec["obs"]).astype(np.float32),
        "reward": np.asarray(rec["reward"], np.float64),
        "terminated": np.asarray(rec["terminated"], np.uint8),
        "equity": np.asarray(rec["equity"], np.float64),
        "position": np.asarray(rec["position"], np.int32),
        "price": np.asarray(rec["price"], np.float64),
        "bar_index": np.asarray(rec["bar_index"], np.int64),
        "trades": np.asarray(rec["trades"], np.int32),
        "commission_paid": np.asarray(rec["commission_paid"], np.float64),
        "total_bars": np.asarray([total_bars], np.int64),
        "summary": summary,
    }
    return out
