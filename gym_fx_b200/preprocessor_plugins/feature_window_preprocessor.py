"""feature_window_preprocessor mirror (reference: preprocessor_plugins/feature_window_preprocessor.py:31-234):
(window, n_features) leakage-safe z-scored feature tensor + price window + agent scalars.  Column validation
happens at lowering time with the reference's error messages; the tensor is assembled by the step kernel."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "feature_window_preprocessor"
    plugin_params = {
        "window_size": 32, "price_column": "CLOSE", "feature_columns": [], "feature_binary_columns": [],
        "feature_scaling": "rolling_zscore", "feature_scaling_window": 256,
        "include_price_window": True, "include_agent_state": True, "feature_clip": 10.0,
    }
    plugin_debug_vars = ["window_size", "price_column", "feature_scaling", "feature_scaling_window",
                         "include_price_window", "include_agent_state"]

    def get_debug_info(self):
        info = {k: self.params.get(k) for k in self.plugin_debug_vars}
        info["n_features"] = len(self.params.get("feature_columns") or [])
        return info

    def add_debug_info(self, debug_info):
        debug_info.update(self.get_debug_info())

    def make_observation(self, *, data, step, bridge_state, config):
        kernel_resident("feature_window_preprocessor.make_observation")
