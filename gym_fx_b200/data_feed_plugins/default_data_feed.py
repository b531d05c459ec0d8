"""default_data_feed mirror (reference: data_feed_plugins/default_data_feed.py:18-79).

`load_data` keeps the reference contract (CSV -> DataFrame indexed by the parsed date column; missing
OPEN/HIGH/LOW/CLOSE filled from the price column; VOLUME defaults to 0).  Instead of `build_bt_feed`
(a backtrader object) it offers `build_table`, which produces the dense float64 [T, n_cols] candle table
+ minutes-since-epoch that fxenv_load_candles() uploads."""
from __future__ import annotations

import numpy as np
import pandas as pd

from ..config import BASE_COLUMNS
from ..plugin_base import PluginBase, kernel_resident


class Plugin(PluginBase):
    plugin_kind = "default_data_feed"
    plugin_params = {"input_data_file": "examples/data/eurusd_sample.csv", "date_column": "DATE_TIME",
                     "headers": True, "max_rows": None, "price_column": "CLOSE"}

    def load_data(self, config) -> pd.DataFrame:
        g = lambda k: config.get(k, self.params[k])
        frame = pd.read_csv(g("input_data_file"), header=0 if bool(g("headers")) else None, nrows=g("max_rows"))
        dcol = g("date_column")
        if dcol in frame.columns:
            frame[dcol] = pd.to_datetime(frame[dcol], errors="coerce")
            frame = frame.dropna(subset=[dcol]).set_index(dcol)
        pcol = g("price_column")
        if pcol not in frame.columns:
            raise ValueError(f"price_column '{pcol}' not found in data")
        for name in BASE_COLUMNS[:4]:
            if name not in frame.columns:
                frame[name] = frame[pcol]
        if "VOLUME" not in frame.columns:
            frame["VOLUME"] = 0
        return frame

    @staticmethod
    def build_table(dataframe: pd.DataFrame, extra_columns=()):
        """DataFrame -> (float64 [T, n_cols] table, column names, int64 [T] minutes since epoch or None)."""
        cols = list(BASE_COLUMNS) + [c for c in extra_columns if c not in BASE_COLUMNS]
        table = np.ascontiguousarray(dataframe[cols].to_numpy(dtype=np.float64))
        minutes = None
        if isinstance(dataframe.index, pd.DatetimeIndex):
            minutes = (dataframe.index.values.astype("datetime64[s]").astype(np.int64) // 60).astype(np.int64)
        return table, cols, minutes

    def build_bt_feed(self, dataframe, config):
        kernel_resident("default_data_feed.build_bt_feed (a backtrader object; use build_table)")
