"""direct_atr_sltp mirror (reference: strategy_plugins/direct_atr_sltp.py:48-263): ATR-sized SL/TP brackets
(simple-mean ATR over `atr_period` true ranges), optional cash-relative sizing and a week-session filter.
Parameters are lowered to FxConfig; TR/ATR, sizing, clamps and order placement run in the step kernel."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "direct_atr_sltp"
    strict_keys = True
    plugin_params = {
        "atr_period": 14, "k_sl": 2.0, "k_tp": 3.0, "position_size": 1.0,
        "rel_volume": None, "leverage": 1.0, "min_order_volume": 0.0, "max_order_volume": 1e12,
        "size_mode": "fx_units", "min_sltp_frac": 0.001, "max_sltp_frac": 0.20,
        "session_filter": False, "entry_dow_start": 0, "entry_hour_start": 12,
        "force_close_dow": 4, "force_close_hour": 20,
    }

    def decide_action(self, obs, info, step: int) -> int:
        return 0

    def on_reset(self, bt_strategy, config) -> None:
        return None

    def apply_action(self, bt_strategy, action, config) -> None:
        kernel_resident("direct_atr_sltp.apply_action")

    def hparam_schema(self):
        return [("atr_period", 7, 30, "int"), ("k_sl", 1.0, 4.0, "float"), ("k_tp", 1.5, 6.0, "float")]
