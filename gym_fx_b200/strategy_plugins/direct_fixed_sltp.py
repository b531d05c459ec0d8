"""direct_fixed_sltp mirror (reference: strategy_plugins/direct_fixed_sltp.py:23-84): fixed-pip SL/TP bracket
around every agent-directed entry.  Parameters are lowered to FxConfig.{sl_pips,tp_pips,pip_size,
strat_position_size}; order placement runs in the step kernel."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "direct_fixed_sltp"
    strict_keys = True
    plugin_params = {"sl_pips": 20.0, "tp_pips": 40.0, "pip_size": 0.0001, "position_size": 1.0}

    def decide_action(self, obs, info, step: int) -> int:
        return 0

    def on_reset(self, bt_strategy, config) -> None:
        return None

    def apply_action(self, bt_strategy, action, config) -> None:
        kernel_resident("direct_fixed_sltp.apply_action")
