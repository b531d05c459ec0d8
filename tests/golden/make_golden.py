#!/usr/bin/env python3
"""
tests/golden/make_golden.py -- regenerates tests/golden/*.npz.

Runs the UNMODIFIED reference (/root/reference: app/env.py + app/bt_bridge.py + its plugins) over the
backtrader/gymnasium shims of oracle/bt_shim, on every scenario of tests/golden/scenarios.py, and stores the
trajectory together with all inputs (candle table, timestamps, config, plugin names, actions) so the parity tests
can replay them WITHOUT the reference tree (which does not exist on the GPU box).

Usage (build container only):  python tests/golden/make_golden.py [scenario_name ...]
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.run_reference import REFERENCE_ROOT, run_reference  # noqa: E402
from gym_fx_b200.synth import write_csv  # noqa: E402
import scenarios as S  # noqa: E402


def _load_fixture(name):
    """Reference CSV fixture -> (table, columns, minutes) via pandas, exactly as the data feed parses it."""
    import pandas as pd

    df = pd.read_csv(os.path.join(REFERENCE_ROOT, "examples", "data", name + ".csv"))
    ts = pd.to_datetime(df["DATE_TIME"])
    minutes = (ts.values.astype("datetime64[s]").astype(np.int64) // 60).astype(np.int64)
    return np.ascontiguousarray(df[S.OHLCV].to_numpy(dtype=np.float64)), list(S.OHLCV), minutes


def generate(sc, outdir):
    cfg = S.full_config(sc)
    if sc["data"][0] == "fixture":
        csv_path = os.path.join(REFERENCE_ROOT, "examples", "data", sc["data"][1] + ".csv")
        table, columns, minutes = _load_fixture(sc["data"][1])
        tmp = None
    else:
        table, columns, minutes = S.make_data(sc["data"])
        tmp = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
        tmp.close()
        csv_path = tmp.name
        if len(columns) == 5:
            write_csv(csv_path, table, minutes)
        else:
            import pandas as pd

            dt = (minutes.astype("int64") * 60).astype("datetime64[s]")
            df = pd.DataFrame({"DATE_TIME": [str(x).replace("T", " ") for x in dt]})
            for j, c in enumerate(columns):
                df[c] = table[:, j] if c != "VOLUME" else table[:, j].astype(np.int64)
            df.to_csv(csv_path, index=False, float_format="%.17g")
    run_cfg = dict(cfg)
    run_cfg["input_data_file"] = csv_path
    actions = S.make_actions(sc["actions"], sc["steps"])
    act_list = [float(a) for a in actions] if actions.dtype.kind == "f" else [int(a) for a in actions]
    traj = run_reference(run_cfg, sc["plugins"], act_list,
                         children_same_bar=bool(sc.get("children_same_bar", False)),
                         extra_steps_after_done=int(sc.get("after_done", 0)))
    if tmp is not None:
        os.unlink(csv_path)
    n = traj["reward"].shape[0]
    every = int(sc.get("obs_every", 1))
    if every > 1:
        keep = np.unique(np.concatenate([np.arange(0, min(n, 40)), np.arange(0, n, every), [n - 1]]))
    else:
        keep = np.arange(n)
    summary = traj.pop("summary")   # GymFxEnv.summary() after close(): includes the analyzers' results
    out = {k: v for k, v in traj.items() if k != "obs"}
    out["obs"] = traj["obs"][keep]
    out["obs_rows"] = keep.astype(np.int64)
    out["candles"] = table
    out["minutes"] = minutes
    out["actions"] = actions
    meta = dict(name=sc["name"], columns=columns, config=cfg, plugins=sc["plugins"],
                children_same_bar=bool(sc.get("children_same_bar", False)),
                summary=summary,
                reference="harveybc/gym-fx@ad8bbc41 over oracle/bt_shim", python=sys.version.split()[0],
                numpy=np.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(outdir, sc["name"] + ".npz")
    np.savez_compressed(path, **out)
    last = n - 1
    print(f"{sc['name']:28s} rows={n:4d} obs_dim={traj['obs'].shape[1]:4d} final_eq={traj['equity'][last]:.6f} "
          f"trades={traj['trades'][last]:3d} term={int(traj['terminated'][last])} "
          f"size={os.path.getsize(path) / 1024:.0f}KB")
    return traj


def main():
    names = set(sys.argv[1:])
    for sc in S.SCENARIOS:
        if names and sc["name"] not in names:
            continue
        generate(sc, HERE)


if __name__ == "__main__":
    main()
