#!/usr/bin/env python3
"""
oracle/time_reference.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Times the UNMODIFIED Python reference (/root/reference app/env.py + app/bt_bridge.py + plugins) over the
backtrader/gymnasium shims of oracle/bt_shim, the way SURVEY section 8(d) asks for the "CPU baseline timed beside
it": loop = tools/smoke_test.py:79-83 pattern (reset, then step with replayed random actions), >= 5000 steps on a
2^15-row synthetic slice, for
    cfg1   default_preprocessor W=32 + default flow (market orders) + pnl_reward
    cfg2   feature_window W=128 F=5 rolling_zscore S=256 + direct_fixed_sltp + pnl_reward   (1 env)
for one process and for P independent processes (P = host cores).  The shim is lighter than real backtrader
(no analyzers, no line buffers), so these rates OVER-state the reference.

Build container only (the GPU box has no /root/reference).  Usage:
    python oracle/time_reference.py [--steps 5000] [--procs P] [--out profiles/r1_reference_python_rate.json]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

PLUGINS = dict(data_feed="default_data_feed", broker="default_broker", strategy="default_strategy",
               preprocessor="default_preprocessor", reward="pnl_reward", metrics="default_metrics")
BASE = {"window_size": 32, "initial_cash": 10000.0, "position_size": 1.0, "commission": 0.0, "slippage": 0.0,
        "price_column": "CLOSE", "date_column": "DATE_TIME", "headers": True, "max_rows": None}
CASES = {
    "cfg1": (dict(BASE), dict(PLUGINS)),
    "cfg2": (dict(BASE, window_size=128, feature_columns=["OPEN", "HIGH", "LOW", "CLOSE", "VOLUME"],
                  feature_scaling_window=256),
             dict(PLUGINS, strategy="direct_fixed_sltp", preprocessor="feature_window_preprocessor")),
}


def _one(args):
    case, csv_path, steps, seed = args
    from oracle.run_reference import build_reference_env

    cfg, plugins = CASES[case]
    cfg = dict(cfg, input_data_file=csv_path)
    env = build_reference_env(cfg, plugins)
    actions = np.random.default_rng(seed).integers(0, 3, steps).tolist()
    env.reset(seed=seed)
    t0 = time.perf_counter()
    done = 0
    for a in actions:
        _, _, term, _, _ = env.step(a)
        done += 1
        if term:
            break
    dt = time.perf_counter() - t0
    env.close()
    return done, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_reference_python_rate.json"))
    a = ap.parse_args()

    from gym_fx_b200.synth import synth_candles, synth_minutes, write_csv

    T = 1 << 15
    tmp = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
    tmp.close()
    write_csv(tmp.name, synth_candles(T, 0, 1000), synth_minutes(T))
    res = {"host": {"cores": os.cpu_count(), "where": "build container (no GPU); /root/reference over oracle/bt_shim"},
           "steps": a.steps, "rows": T, "cases": {}}
    try:
        for case in CASES:
            n, dt = _one((case, tmp.name, a.steps, 1))
            single = n / dt
            with mp.get_context("fork").Pool(a.procs) as pool:
                t0 = time.perf_counter()
                outs = pool.map(_one, [(case, tmp.name, a.steps, 100 + i) for i in range(a.procs)])
                wall = time.perf_counter() - t0
            res["cases"][case] = {
                "one_process_steps_per_s": single,
                "procs": a.procs,
                "all_procs_steps_per_s_sum_of_rates": sum(n / dt for n, dt in outs),
                "all_procs_steps_per_s_wall_incl_startup": sum(n for n, _ in outs) / wall,
            }
            print(case, json.dumps(res["cases"][case]), flush=True)
    finally:
        os.unlink(tmp.name)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
