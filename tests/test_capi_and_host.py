"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/fxenv.h declares, the product fails
LOUDLY without a GPU (no CPU fallback), and the host-side logic (config lowering with the reference's per-plugin
precedence rules, plugin mirrors, plugin loader, spaces) behaves like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import scenarios as S
from gym_fx_b200 import _native
from gym_fx_b200.config import FxConfig, lower_config, obs_dim, obs_layout
from gym_fx_b200.plugin_base import KernelResident
from gym_fx_b200.plugin_loader import get_plugin_params, load_plugin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plugins(cfgd, **kw):
    return S.build_mirror_plugins(cfgd, {**S.DEFAULT_PLUGINS, **kw})


def _lower(cfgd, columns=S.OHLCV, **kw):
    pl = _plugins(cfgd, **kw)
    return lower_config(cfgd, broker_plugin=pl["broker"], strategy_plugin=pl["strategy"],
                        preprocessor_plugin=pl["preprocessor"], reward_plugin=pl["reward"], columns=columns)


def test_library_exports_every_header_symbol():
    _native.build()
    hdr = open(os.path.join(ROOT, "include", "fxenv.h")).read()
    declared = sorted(set(re.findall(r"\b(fxenv_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 16
    L = _native.load()
    for sym in declared:
        assert hasattr(L, sym), f"libfxenv.so does not export {sym}"
    assert L.fxenv_abi_version() == 2
    assert sorted(_native.EXPORTS) == declared


def test_create_validates_and_fails_loudly_without_gpu():
    L = _native.load()
    cfg = _lower({**S.DEFAULTS})
    cfg.struct_size = 8
    h = C.c_void_p()
    assert L.fxenv_create(C.byref(cfg), C.byref(h)) == -1
    assert b"ABI mismatch" in L.fxenv_last_error(None)
    cfg = _lower({**S.DEFAULTS, "slippage": 1.5})
    assert L.fxenv_create(C.byref(cfg), C.byref(h)) == -1 and b"slippage" in L.fxenv_last_error(None)
    if not torch.cuda.is_available():
        cfg = _lower({**S.DEFAULTS})
        rc = L.fxenv_create(C.byref(cfg), C.byref(h))
        assert rc == -2 and b"no CPU path" in L.fxenv_last_error(None)
        from gym_fx_b200.vec_env import VecFxEnv
        with pytest.raises(_native.FxEnvError, match="no CPU path"):
            VecFxEnv(cfg, [np.zeros((100, 5))])


def test_plugin_mirrors_have_no_host_compute():
    pl = _plugins({**S.DEFAULTS}, strategy="direct_atr_sltp", reward="sharpe_reward",
                  preprocessor="feature_window_preprocessor")
    with pytest.raises(KernelResident):
        pl["reward"].compute_reward(prev_equity=1.0, new_equity=2.0, step=1, config={})
    with pytest.raises(KernelResident):
        pl["preprocessor"].make_observation(data=None, step=1, bridge_state={}, config={})
    with pytest.raises(KernelResident):
        pl["strategy"].apply_action(None, 1, {})
    with pytest.raises(KernelResident):
        pl["broker"].build_bt_broker({})
    assert pl["strategy"].hparam_schema()[0] == ("atr_period", 7, 30, "int")


def test_lowering_follows_each_plugins_precedence_rule():
    # direct_*_sltp: plugin params overridden by NON-None config values of its own keys (direct_fixed_sltp.py:79-84)
    c = _lower({**S.DEFAULTS, "sl_pips": 7.0, "position_size": 3.0}, strategy="direct_fixed_sltp")
    assert (c.sl_pips, c.tp_pips, c.strat_position_size, c.position_size) == (7.0, 40.0, 3.0, 3.0)
    # broker: config.get(key, params[key]); 'slippage' legacy key (default_broker.py:39-44); reward initial_cash "or 1.0"
    c = _lower({**S.DEFAULTS, "commission": 2e-5, "leverage": 30.0, "initial_cash": 5000.0})
    assert (c.commission, c.leverage, c.reward_initial_cash, c.min_equity) == (2e-5, 30.0, 5000.0, 50.0)
    # sharpe window comes from plugin.params["window"] (set_params), annualization from config (sharpe_reward.py:24-32,57)
    c = _lower({**S.DEFAULTS, "window": 16, "annualization_factor": 100.0}, reward="sharpe_reward")
    assert (c.reward, c.sharpe_window, c.annualization_factor) == (1, 16, 100.0)
    # atr: rel_volume None disables sizing; None min/max frac disables the clamps (direct_atr_sltp.py:166-175,204-207)
    c = _lower({**S.DEFAULTS, "min_sltp_frac": None, "rel_volume": 0.1, "size_mode": "notional"}, strategy="direct_atr_sltp")
    assert (c.use_min_frac, c.use_max_frac, c.use_rel_volume, c.size_mode) == (0, 1, 1, 1)
    # default_strategy has no apply_action -> default market flow (app/bt_bridge.py:171-190)
    assert _lower({**S.DEFAULTS}).strategy == 0


def test_feature_window_errors_match_the_reference_tests():
    # tests/test_feature_window_preprocessor.py:109-128 of the reference
    with pytest.raises(ValueError, match="missing from dataframe"):
        _lower({**S.DEFAULTS, "feature_columns": ["CLOSE", "does_not_exist"], "feature_scaling": "none"},
               preprocessor="feature_window_preprocessor")
    with pytest.raises(ValueError, match="non-empty"):
        _lower({**S.DEFAULTS, "feature_columns": []}, preprocessor="feature_window_preprocessor")
    with pytest.raises(ValueError, match="feature_scaling must be one of"):
        _lower({**S.DEFAULTS, "feature_columns": ["CLOSE"], "feature_scaling": "bogus"},
               preprocessor="feature_window_preprocessor")
    with pytest.raises(ValueError, match="price_column 'NOPE' not found"):
        _lower({**S.DEFAULTS, "price_column": "NOPE"})


def test_obs_layout_matches_reference_shapes():
    # tests/test_feature_window_preprocessor.py:34-58: features (W, F), prices (W,), returns (W,), 4 x (1,)
    cols = S.OHLCV + ["FEAT_A", "BIN_FLAG"]
    c = _lower({**S.DEFAULTS, "window_size": 32, "feature_columns": cols, "feature_binary_columns": ["BIN_FLAG"]},
               columns=cols, preprocessor="feature_window_preprocessor")
    lay = obs_layout(c)
    assert lay["features"] == (0, (32, 7)) and lay["prices"] == (224, (32,)) and lay["returns"] == (256, (32,))
    assert [lay[k][0] for k in ("position", "equity_norm", "unrealized_pnl_norm", "steps_remaining_norm")] == [288, 289, 290, 291]
    assert obs_dim(c) == 292 and list(c.feature_binary)[:7] == [0, 0, 0, 0, 0, 0, 1]
    c = _lower({**S.DEFAULTS})
    assert obs_dim(c) == 68 and "features" not in obs_layout(c)


def test_plugin_loader_contract():
    for group, names in {"strategy.plugins": ["default_strategy", "direct_fixed_sltp", "direct_atr_sltp"],
                         "reward.plugins": ["pnl_reward", "sharpe_reward", "dd_penalized_reward"],
                         "preprocessor.plugins": ["default_preprocessor", "feature_window_preprocessor"],
                         "broker.plugins": ["default_broker"], "data_feed.plugins": ["default_data_feed"],
                         "metrics.plugins": ["default_metrics"]}.items():
        for name in names:
            cls, keys = load_plugin(group, name)
            assert cls.__name__ == "Plugin" and keys == list(cls.plugin_params.keys())
    assert get_plugin_params("reward.plugins", "sharpe_reward")["window"] == 64
    with pytest.raises(ImportError):
        load_plugin("reward.plugins", "nope")


def test_default_strategy_driver_modes():
    from gym_fx_b200.strategy_plugins.default_strategy import Plugin
    assert [Plugin({"driver_mode": "buy_hold"}).decide_action(None, None, k) for k in range(3)] == [1, 0, 0]
    assert Plugin({"driver_mode": "flat"}).decide_action(None, None, 0) == 0
    a = [Plugin({"driver_mode": "random", "seed": 5}).decide_action(None, None, k) for k in range(5)]
    import random
    r = random.Random(5)
    assert a[:1] == [r.choice([0, 1, 2])]
    # the whole stream, per step and as an up-front [steps, envs] table for VecFxEnv.step_many
    r = random.Random(9)
    want = [r.choice([0, 1, 2]) for _ in range(12)]
    p = Plugin({"driver_mode": "random", "seed": 9})
    assert [p.decide_action(None, None, k) for k in range(12)] == want
    tab = Plugin({"driver_mode": "random", "seed": 9}).action_table(12, 3)
    assert tab.shape == (12, 3) and tab.dtype.name == "int32" and tab[:, 0].tolist() == want and (tab[:, 2] == tab[:, 0]).all()
    import os, tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        fh.write("DATE_TIME,action\na,2\nb,0\nc,1\n")
    try:
        rp = Plugin({"driver_mode": "replay", "replay_actions_file": fh.name})
        assert [rp.decide_action(None, None, k) for k in range(5)] == [2, 0, 1, 0, 0]
    finally:
        os.unlink(fh.name)


def test_metrics_summary_matches_reference_golden_fields():
    # examples/results/buy_hold_summary.json of the reference
    from gym_fx_b200.metrics_plugins.default_metrics import Plugin
    s = Plugin().summarize(initial_cash=10000.0, final_equity=10000.095791583166, analyzers={}, config={})
    assert s["total_return"] == 9.579158316563863e-06 and s["trades_total"] == 0 and s["sharpe_ratio"] is None


def test_rollout_ticket_plan_covers_every_step_once(monkeypatch):
    """fx_rollout_plan (the rounds of a fxenv_step_many batch: one ticket = one env for the steps of one round) is pure host
    arithmetic: the rounds partition [0, n_steps) in order, are at most 64 steps long, a batch whose envs all have a
    resident warp is a single round (no hand-over), larger batches leave at least ~6 tickets per resident warp unless
    the rounds are already single steps, and FXENV_CHUNK forces the length (remainder as a shorter last round)."""
    import ctypes as C
    from gym_fx_b200 import _native
    L = _native.load()
    f = L.fxenv_debug_rollout_plan
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    f.restype = C.c_int

    def plan(n_envs, warps, steps):
        buf = (C.c_int * 8192)()
        n = f(n_envs, warps, steps, buf, 8192)
        assert n >= 1
        return list(buf[:n + 1])

    monkeypatch.delenv("FXENV_CHUNK", raising=False)
    rng = np.random.default_rng(0)
    cases = [(4096, 2368, 20), (4096, 2368, 500), (4096, 2368, 1), (64, 2368, 12), (16384, 2368, 300), (5000, 2368, 23),
             (2369, 2368, 1000), (1, 1, 7)]
    cases += [(int(rng.integers(1, 40000)), int(rng.integers(1, 4000)), int(rng.integers(1, 3000))) for _ in range(200)]
    for n_envs, warps, steps in cases:
        st = plan(n_envs, warps, steps)
        assert st[0] == 0 and st[-1] == steps and all(b > a for a, b in zip(st, st[1:])), (n_envs, warps, steps, st)
        lens = [b - a for a, b in zip(st, st[1:])]
        assert max(lens) <= max(64, steps if n_envs <= warps else 0)
        if n_envs <= warps:
            assert lens == [steps]                                  # every env has a warp of its own: one round
        else:
            assert len(set(lens[:-1])) <= 1 and lens[-1] <= lens[0]  # uniform rounds, the remainder last and shorter
            if lens[0] > 1:
                assert n_envs * len(lens) >= 6 * warps * 0.99 or lens[0] == 64, (n_envs, warps, steps, lens)
    monkeypatch.setenv("FXENV_CHUNK", "7")
    assert plan(5000, 2368, 23) == [0, 7, 14, 21, 23]
    assert plan(64, 2368, 5) == [0, 5]
