"""sharpe_reward mirror (reference: reward_plugins/sharpe_reward.py:34-58): rolling annualised Sharpe of the
per-step returns (deque of `window`, sample variance, Python >= 3.12 compensated sums).
Evaluated in fp64 inside the step kernel; the per-env ring lives in the device struct-of-arrays."""
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "sharpe_reward"
    plugin_params = {"window": 64, "annualization_factor": 252.0, "initial_cash": 10000.0}

    def compute_reward(self, *, prev_equity, new_equity, step, config):
        kernel_resident("sharpe_reward.compute_reward")
