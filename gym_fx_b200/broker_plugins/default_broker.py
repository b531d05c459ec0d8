"""default_broker mirror (reference: broker_plugins/default_broker.py:19-53).

The reference plugin configures a backtrader BackBroker (cash, PERC commission as an absolute fraction of
notional, leverage, optional % slippage); the accounting itself lives in backtrader.  Here the four numbers are
lowered into FxConfig and the broker arithmetic (submit -> margin check -> match -> brackets -> value) runs in
the step kernel (gym_fx_b200/csrc/fx_core.cuh)."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "default_broker"
    plugin_params = {"initial_cash": 10000.0, "commission": 0.0, "slippage_perc": 0.0, "leverage": 1.0}

    def build_bt_broker(self, config):
        kernel_resident("default_broker.build_bt_broker (a backtrader object)")
