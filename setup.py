"""Packaging of the B200-native gym-fx env: the reference's import paths (`app`, `gym_fx`) and its six plugin
entry-point groups with the same plugin names (reference setup.py:11-35), resolved to the mirrors in gym_fx_b200.
The CUDA library is built in-tree by `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a)."""
from setuptools import find_packages, setup

P = "gym_fx_b200"

setup(
    name="gym-fx-b200",
    version="0.2.0",
    packages=find_packages(include=["gym_fx_b200*", "gym_fx*", "app*"]),
    package_data={"gym_fx_b200": ["libfxenv.so", "csrc/*"]},
    entry_points={
        "data_feed.plugins": [f"default_data_feed={P}.data_feed_plugins.default_data_feed:Plugin"],
        "broker.plugins": [f"default_broker={P}.broker_plugins.default_broker:Plugin",
                           f"oanda_broker={P}.broker_plugins.oanda_broker:Plugin"],
        "strategy.plugins": [f"default_strategy={P}.strategy_plugins.default_strategy:Plugin",
                             f"direct_fixed_sltp={P}.strategy_plugins.direct_fixed_sltp:Plugin",
                             f"direct_atr_sltp={P}.strategy_plugins.direct_atr_sltp:Plugin"],
        "preprocessor.plugins": [f"default_preprocessor={P}.preprocessor_plugins.default_preprocessor:Plugin",
                                 f"feature_window_preprocessor={P}.preprocessor_plugins.feature_window_preprocessor:Plugin"],
        "reward.plugins": [f"pnl_reward={P}.reward_plugins.pnl_reward:Plugin",
                           f"sharpe_reward={P}.reward_plugins.sharpe_reward:Plugin",
                           f"dd_penalized_reward={P}.reward_plugins.dd_penalized_reward:Plugin"],
        "metrics.plugins": [f"default_metrics={P}.metrics_plugins.default_metrics:Plugin"],
    },
    install_requires=["numpy", "pandas", "torch"],
    description="B200-native vectorised drop-in for the env.step() hot path of harveybc/gym-fx.",
)
