"""Shared helpers of the parity tests: golden loading, FxConfig construction, trajectory diffing."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

import scenarios as S
from gym_fx_b200.config import lower_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(bytes(g["meta"]).decode())
    return g


def config_from_meta(meta, num_envs=1, **kw):
    cfg = meta["config"]
    pl = S.build_mirror_plugins(cfg, meta["plugins"])
    return lower_config(cfg, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"],
                        preprocessor_plugin=pl["preprocessor"], reward_plugin=pl["reward"],
                        columns=meta["columns"], num_envs=num_envs,
                        children_same_bar=meta.get("children_same_bar", False), **kw)


def replay(env_like, g, want_info):
    """Drive an OracleVec-like object (reset/step/info) with the golden's actions; returns a trajectory dict
    shaped like the golden (row 0 = reset)."""
    n = g["reward"].shape[0]
    rec = {k: [] for k in ("obs", "reward", "terminated", "equity", "position", "price", "bar_index", "trades",
                          "commission_paid")}

    def push(obs, r, t):
        inf = want_info()
        rec["obs"].append(obs[0].copy())
        rec["reward"].append(float(r))
        rec["terminated"].append(int(t))
        for k in ("equity", "position", "price", "bar_index", "trades", "commission_paid"):
            rec[k].append(inf[k][0])

    obs = env_like.reset(np.zeros(1, np.int64))
    push(obs, 0.0, 0)
    for k in range(n - 1):
        a = g["actions"][k:k + 1]
        obs, rew, rew64, term = env_like.step(a)
        push(obs, rew64[0], term[0])
    return {k: np.asarray(v) for k, v in rec.items()}


def assert_traj_matches(traj, g, *, obs_rtol=1e-6, obs_atol=1e-6, reward_rtol=1e-9, reward_atol=1e-12,
                        exact_state=True, label=""):
    n = g["reward"].shape[0]
    assert traj["reward"].shape[0] == n, f"{label}: length {traj['reward'].shape[0]} != {n}"
    # integer / index state: bit-exact
    for k in ("position", "bar_index", "trades", "terminated"):
        a, b = np.asarray(traj[k]).astype(np.int64), np.asarray(g[k]).astype(np.int64)
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"{label}: {k} differs first at row {bad[0]}: {a[bad[0]]} != {b[bad[0]]}"
    # fp64 state: bit-exact when the implementation is fp64 operation-for-operation
    for k in ("equity", "price", "commission_paid"):
        a, b = np.asarray(traj[k], np.float64), np.asarray(g[k], np.float64)
        if exact_state:
            bad = np.nonzero(a != b)[0]
            assert bad.size == 0, f"{label}: {k} differs first at row {bad[0]}: {a[bad[0]]!r} != {b[bad[0]]!r}"
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-9, err_msg=f"{label}: {k}")
    np.testing.assert_allclose(traj["reward"], g["reward"], rtol=reward_rtol, atol=reward_atol,
                               err_msg=f"{label}: reward")
    rows = g["obs_rows"]
    np.testing.assert_allclose(np.asarray(traj["obs"])[rows], g["obs"], rtol=obs_rtol, atol=obs_atol,
                               err_msg=f"{label}: obs")


SUMMARY_KEYS = ("max_drawdown_pct", "max_drawdown_money", "trades_total", "trades_won", "trades_lost", "avg_trade_pnl", "sqn")


def assert_summary_matches(got, want, label="", sqn_rtol=1e-9):
    """got / want: dicts with the analyzer-derived fields of metrics_plugins/default_metrics.py:48-60 (None or NaN = the
    reference's None).  Counters and the drawdown / average-pnl values (same fp64 operations in the same order) must be
    equal; `sqn` is a running-moment evaluation on the device side vs math.fsum in the reference: 1e-9 relative."""
    def norm(v):
        if v is None:
            return None
        v = float(v)
        return None if v != v else v
    for k in SUMMARY_KEYS:
        a, b = norm(got[k]), norm(want[k])
        if k.startswith("trades_"):
            assert int(a or 0) == int(b or 0), f"{label}: {k} {a} != {b}"
        elif k == "sqn":
            assert (a is None) == (b is None), f"{label}: sqn {a} vs {b}"
            if a is not None:
                assert abs(a - b) <= sqn_rtol * max(1.0, abs(b)), f"{label}: sqn {a!r} vs {b!r}"
        else:
            assert a == b, f"{label}: {k} {a!r} != {b!r}"
