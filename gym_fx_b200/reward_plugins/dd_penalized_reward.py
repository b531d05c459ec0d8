"""dd_penalized_reward mirror (reference: reward_plugins/dd_penalized_reward.py:30-47):
pnl_norm - penalty_lambda * (peak - equity) / initial_cash, with a per-env running equity peak.
Evaluated in fp64 inside the step kernel."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "dd_penalized_reward"
    plugin_params = {"penalty_lambda": 1.0, "initial_cash": 10000.0}

    def compute_reward(self, *, prev_equity, new_equity, step, config):
        kernel_resident("dd_penalized_reward.compute_reward")
