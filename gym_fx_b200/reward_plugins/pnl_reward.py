"""pnl_reward mirror (reference: reward_plugins/pnl_reward.py:26-36): (new - prev) / initial_cash * reward_scale.
Evaluated in fp64 inside the step kernel (per-env state lives in the device struct-of-arrays)."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "pnl_reward"
    plugin_params = {"reward_scale": 1.0, "initial_cash": 10000.0}

    def compute_reward(self, *, prev_equity, new_equity, step, config):
        kernel_resident("pnl_reward.compute_reward")
