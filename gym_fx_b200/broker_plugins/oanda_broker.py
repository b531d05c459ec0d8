"""oanda_broker: live-trading stub of the reference (broker_plugins/oanda_broker.py:42-63) -- out of scope
for the simulated GPU env; kept only so the entry-point name resolves."""
from ..plugin_base import PluginBase


class Plugin(PluginBase):
    plugin_kind = "oanda_broker"
    plugin_params = {"oanda_token": None, "oanda_account": None, "oanda_practice": True}

    def build_bt_broker(self, config):
        raise NotImplementedError("live OANDA trading is out of scope for the vectorised simulation env")
