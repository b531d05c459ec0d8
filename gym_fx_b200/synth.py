"""
Synthetic OHLCV candle streams for benchmarks and parity tests.

The reference ships only two 500-row fixtures (examples/data/*.csv) and its default
``input_data_file`` does not exist (app/config.py:16), so benchmark-size streams are
generated: seeded, modelled on examples/data/eurusd_sample.csv statistics.  Layout is the
one the data feed defines (data_feed_plugins/default_data_feed.py:36-56): float64 rows
``[OPEN, HIGH, LOW, CLOSE, VOLUME]`` plus 1-minute timestamps.
"""
from __future__ import annotations

import numpy as np

PAIR_BASES = (1.10, 1.27, 0.66, 150.0)      # cfg5: 4 pairs, env i uses pair i % 4
PAIR_DECIMALS = (5, 5, 5, 3)
PAIR_PIP = (1e-4, 1e-4, 1e-4, 1e-2)
EPOCH_2024_MIN = 28401120                    # minutes from 1970-01-01 to 2024-01-01 00:00 (a Monday)


def synth_candles(T: int = 1 << 19, pair: int = 0, seed: int | None = None) -> np.ndarray:
    """float64 [T, 5] candle table for `pair` (0..3); seed defaults to 1000 + pair."""
    rng = np.random.default_rng(1000 + pair if seed is None else seed)
    base, dec = PAIR_BASES[pair], PAIR_DECIMALS[pair]
    scale = base / 1.10
    close = np.round(base * np.exp(np.cumsum(rng.normal(0.0, 3e-4, T))), dec)
    prev = np.concatenate([[close[0]], close[:-1]])
    open_ = np.round(prev + rng.normal(0.0, 5e-5 * scale, T), dec)
    high = np.round(np.maximum(open_, close) + np.abs(rng.normal(0.0, 1.5e-4 * scale, T)), dec)
    low = np.round(np.minimum(open_, close) - np.abs(rng.normal(0.0, 1.5e-4 * scale, T)), dec)
    vol = rng.integers(110, 2000, T).astype(np.float64)
    return np.ascontiguousarray(np.stack([open_, high, low, close, vol], axis=1))


def synth_minutes(T: int) -> np.ndarray:
    """int64 [T] minutes since the Unix epoch, 1-minute bars from 2024-01-01 00:00."""
    return EPOCH_2024_MIN + np.arange(T, dtype=np.int64)


def write_csv(path: str, candles: np.ndarray, minutes: np.ndarray | None = None, decimals: int = 5) -> None:
    """Write a table in the reference's CSV layout (DATE_TIME,OPEN,HIGH,LOW,CLOSE,VOLUME)."""
    T = candles.shape[0]
    if minutes is None:
        minutes = synth_minutes(T)
    dt = (minutes.astype("int64") * 60).astype("datetime64[s]")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("DATE_TIME,OPEN,HIGH,LOW,CLOSE,VOLUME\n")
        for i in range(T):
            o, h, l, c, v = candles[i]
            ts = str(dt[i]).replace("T", " ")
            fh.write(f"{ts},{float(o)!r},{float(h)!r},{float(l)!r},{float(c)!r},{int(v)}\n")


def start_offsets(num_envs: int, T: int, steps: int, S: int) -> np.ndarray:
    """Deterministic spread of env start bars: o_i = (i * 9973) mod (T - steps - S - 2)."""
    span = max(1, T - steps - S - 2)
    return (np.arange(num_envs, dtype=np.int64) * 9973) % span
