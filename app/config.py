"""`app.config.DEFAULT_VALUES` (reference: app/config.py:1-45): the default "dict of everything" callers start from
(tools/smoke_test.py:40-55, tools/check_gym_compliance.py:30-38).  Same keys and values."""
DEFAULT_VALUES = dict(
    mode="inference", driver_mode="buy_hold", steps=500,
    data_feed_plugin="default_data_feed", broker_plugin="default_broker", strategy_plugin="default_strategy",
    preprocessor_plugin="default_preprocessor", reward_plugin="pnl_reward", metrics_plugin="default_metrics",
    input_data_file="examples/data/eurusd.csv", date_column="DATE_TIME", price_column="CLOSE", instrument="EUR_USD",
    timeframe="M1", headers=True, max_rows=None,
    window_size=32, initial_cash=10000.0, position_size=1.0, commission=0.0, slippage=0.0,
    replay_actions_file=None,
    remote_log=None, remote_load_config=None, remote_save_config=None, username=None, password=None, load_config=None,
    save_config="./config_out.json", save_log="./debug_out.json", results_file="./results.json", quiet_mode=False,
)
