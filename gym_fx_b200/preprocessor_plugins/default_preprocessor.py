"""default_preprocessor mirror (reference: preprocessor_plugins/default_preprocessor.py:19-77): price window +
first differences + 4 agent scalars.  Assembled by the step kernel straight into the flat observation row."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "default_preprocessor"
    plugin_params = {"window_size": 32, "price_column": "CLOSE"}

    def make_observation(self, *, data, step, bridge_state, config):
        kernel_resident("default_preprocessor.make_observation")
