"""Plugin lookup with the reference's contract (app/plugin_loader.py:12-83): `load_plugin(group, name) ->
(class, param_keys)`, `get_plugin_params(group, name) -> dict`.  Installed entry points win (same six groups as
the reference's setup.py:11-35); otherwise the built-in mirrors of this package are used."""
from __future__ import annotations

import importlib
from importlib.metadata import entry_points

BUILTIN = {
    "data_feed.plugins": {"default_data_feed": "data_feed_plugins.default_data_feed"},
    "broker.plugins": {"default_broker": "broker_plugins.default_broker", "oanda_broker": "broker_plugins.oanda_broker"},
    "strategy.plugins": {"default_strategy": "strategy_plugins.default_strategy",
                         "direct_fixed_sltp": "strategy_plugins.direct_fixed_sltp",
                         "direct_atr_sltp": "strategy_plugins.direct_atr_sltp"},
    "preprocessor.plugins": {"default_preprocessor": "preprocessor_plugins.default_preprocessor",
                             "feature_window_preprocessor": "preprocessor_plugins.feature_window_preprocessor"},
    "reward.plugins": {"pnl_reward": "reward_plugins.pnl_reward", "sharpe_reward": "reward_plugins.sharpe_reward",
                       "dd_penalized_reward": "reward_plugins.dd_penalized_reward"},
    "metrics.plugins": {"default_metrics": "metrics_plugins.default_metrics"},
}


def load_plugin(plugin_group: str, plugin_name: str):
    try:
        for ep in entry_points().select(group=plugin_group):
            if ep.name == plugin_name:
                cls = ep.load()
                return cls, list(cls.plugin_params.keys())
    except Exception:
        pass
    mod = BUILTIN.get(plugin_group, {}).get(plugin_name)
    if mod is None:
        raise ImportError(f"Plugin {plugin_name} not found in group {plugin_group}.")
    cls = importlib.import_module(f"{__package__}.{mod}").Plugin
    return cls, list(cls.plugin_params.keys())


def get_plugin_params(plugin_group: str, plugin_name: str):
    try:
        return load_plugin(plugin_group, plugin_name)[0].plugin_params
    except Exception as exc:
        raise ImportError(f"Failed to get plugin params for {plugin_name} from group {plugin_group}, Error: {exc}")


# app/config.py:1-45 of the reference, only the keys the hot path reads (so that callers do not need the reference tree)
DEFAULT_VALUES = {
    "window_size": 32, "initial_cash": 10000.0, "position_size": 1.0, "commission": 0.0, "slippage": 0.0,
    "price_column": "CLOSE", "date_column": "DATE_TIME", "headers": True, "max_rows": None,
}

DEFAULT_PLUGINS = dict(data_feed="default_data_feed", broker="default_broker", strategy="default_strategy",
                       preprocessor="default_preprocessor", reward="pnl_reward", metrics="default_metrics")

_GROUP_OF = {"data_feed": "data_feed.plugins", "broker": "broker.plugins", "strategy": "strategy.plugins",
             "preprocessor": "preprocessor.plugins", "reward": "reward.plugins", "metrics": "metrics.plugins"}


def build_plugins(config: dict, plugins: dict) -> dict:
    """Instantiate one plugin per role the way the reference's driver does (app/main.py:20-24: `klass(config)` then
    `set_params(**config)`).  `plugins` maps role (data_feed, broker, strategy, preprocessor, reward, metrics) -> name."""
    out = {}
    for role, name in plugins.items():
        klass, _ = load_plugin(_GROUP_OF[role], name)
        inst = klass(config)
        inst.set_params(**config)
        out[role] = inst
    return out
