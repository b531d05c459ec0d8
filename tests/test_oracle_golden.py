"""CPU (-m "not gpu"): pins the C oracle (oracle/fxenv_oracle.c) against
  * the reference's two known answers produced through REAL backtrader
    (examples/results/buy_hold_summary.json:3-4, flat_summary.json:3-4), and
  * every trajectory in tests/golden/ (the reference's own files run unmodified over oracle/bt_shim)."""
import numpy as np
import pytest

from common import assert_traj_matches, config_from_meta, golden_names, load_golden, replay
from oracle.c_oracle import OracleVec


def _run(name):
    g = load_golden(name)
    cfg = config_from_meta(g["meta"])
    env = OracleVec(cfg, [g["candles"]], [g["minutes"]])
    traj = replay(env, g, env.info)
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
