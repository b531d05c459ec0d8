"""CPU (-m "not gpu"): pins the C oracle (oracle/fxenv_oracle.c) against
  * the reference's two known answers produced through REAL backtrader
    (examples/results/buy_hold_summary.json:3-4, flat_summary.json:3-4), and
  * every trajectory in tests/golden/ (the reference's own files run unmodified over oracle/bt_shim)."""
import numpy as np
import pytest

from common import assert_summary_matches, assert_traj_matches, config_from_meta, golden_names, load_golden, replay
from oracle.c_oracle import OracleVec


def _run(name):
    g = load_golden(name)
    cfg = config_from_meta(g["meta"])
    env = OracleVec(cfg, [g["candles"]], [g["minutes"]])
    traj = replay(env, g, env.info)
    traj["summary"] = {k: v[0] for k, v in env.summary().items()}
    return g, traj


def test_reference_known_answer_buy_hold():
    # examples/results/buy_hold_summary.json:3-4 (480 steps of tools/smoke_test.py:121-136)
    g, traj = _run("buy_hold_uptrend")
    assert traj["equity"][-1] == 10000.095791583166
    assert (traj["equity"][-1] / 10000.0 - 1.0) == 9.579158316563863e-06
    assert traj["trades"][-1] == 0


def test_reference_known_answer_flat():
    g, traj = _run("flat_sample")
    assert traj["equity"][-1] == 10000.0 and np.all(traj["reward"] == 0.0)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g, traj = _run(name)
    assert_traj_matches(traj, g, label=name)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_summary_matches_reference_analyzers(name):
    """GymFxEnv.summary() of the reference after the run has ended (app/env.py:256-271): the analyzer-derived fields of
    metrics_plugins/default_metrics.py:48-60 (backtrader DrawDown / TradeAnalyzer / SQN, restated in oracle/bt_shim)."""
    g, traj = _run(name)
    assert_summary_matches(traj["summary"], g["meta"]["summary"], label=name, sqn_rtol=1e-12)
