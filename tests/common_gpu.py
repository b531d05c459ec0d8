"""Adapters that give the GPU VecFxEnv the same numpy-facing surface as oracle.c_oracle.OracleVec."""
from __future__ import annotations

import numpy as np
import torch

from gym_fx_b200.vec_env import VecFxEnv

INFO_KEYS = ("equity", "prev_equity", "price", "cash", "position_size", "position_price", "commission_paid",
             "position", "bar_index", "total_bars", "trades", "n_orders", "flags")


class GpuVec:
    def __init__(self, cfg, candles, minutes=None):
        self.env = VecFxEnv(cfg, candles, minutes)
        self.N = self.env.num_envs

    def reset(self, start_bars=None, mask=None):
        sb = None if start_bars is None else torch.as_tensor(np.asarray(start_bars, np.int64))
        mk = None if mask is None else torch.as_tensor(np.asarray(mask, np.uint8))
        obs, _ = self.env.reset(sb, mk)
        return obs.cpu().numpy()

    def step(self, actions, want_obs=True):
        a = torch.as_tensor(np.ascontiguousarray(actions))
        obs, rew, term, _, _ = self.env.step(a)
        return (obs.cpu().numpy() if want_obs else None, rew.cpu().numpy(), self.env.reward64.cpu().numpy(),
                term.cpu().numpy().astype(np.uint8))

    def info(self):
        i = self.env.info()
        return {k: i[k].cpu().numpy() for k in INFO_KEYS}

    def close(self):
        self.env.close()


def compare_step(tag, go, oo, *, obs_rtol=1e-5, obs_atol=2e-6):
    """go / oo = (obs, rew32, rew64, term) from GPU and oracle.  Returns the number of non-identical obs floats."""
    gobs, grew, grew64, gterm = go
    oobs, orew, orew64, oterm = oo
    np.testing.assert_array_equal(gterm, oterm, err_msg=f"{tag}: terminated")
    np.testing.assert_allclose(grew64, orew64, rtol=1e-9, atol=1e-13, err_msg=f"{tag}: reward (fp64)")
    np.testing.assert_allclose(grew, orew, rtol=1e-5, atol=1e-12, err_msg=f"{tag}: reward (fp32, 1e-5 rel)")
    if gobs is not None:
        np.testing.assert_allclose(gobs, oobs, rtol=obs_rtol, atol=obs_atol, err_msg=f"{tag}: obs")
        return int(np.count_nonzero(gobs != oobs))
    return 0


def compare_info(tag, gi, oi):
    for k in ("position", "bar_index", "total_bars", "trades"):  # n_orders: oracle counts orders, GPU counts entries
        np.testing.assert_array_equal(gi[k], oi[k], err_msg=f"{tag}: {k}")
    # bit 32 (FX_FLAG_TRADE_PRICE_OWN) is bookkeeping of the product's trade statistics, not env status
    np.testing.assert_array_equal(gi["flags"].astype(np.uint32) & 31, oi["flags"].astype(np.uint32) & 31, err_msg=f"{tag}: flags")
    for k in ("equity", "prev_equity", "price", "cash", "position_size", "position_price", "commission_paid"):
        a, b = gi[k], oi[k]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"{tag}: {k} not bit-exact at env {bad[0]}: {a[bad[0]]!r} vs {b[bad[0]]!r}"


def gpu_summary(env, i=None):
    """VecFxEnv.summary() as numpy: dict of per-env arrays (or of env i's scalars) of the analyzer-derived fields."""
    sm = env.summary()
    out = {k: sm[k].cpu().numpy() for k in ("max_drawdown_pct", "max_drawdown_money", "trades_total", "trades_won",
                                            "trades_lost", "trades_closed", "avg_trade_pnl", "sqn")}
    return out if i is None else {k: v[i] for k, v in out.items()}


def compare_summary(tag, gs, osum):
    """GPU summary arrays vs OracleVec.summary(): drawdown / average pnl bit-exact, counters equal, sqn 1e-9."""
    for k in ("trades_total", "trades_won", "trades_lost", "trades_closed"):
        np.testing.assert_array_equal(gs[k].astype(np.int64), osum[k].astype(np.int64), err_msg=f"{tag}: {k}")
    for k in ("max_drawdown_pct", "max_drawdown_money", "avg_trade_pnl"):
        a, b = gs[k], osum[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{tag}: {k} NaN pattern"
        ok = ~np.isnan(a)
        bad = np.nonzero(a[ok] != b[ok])[0]
        assert bad.size == 0, f"{tag}: {k} not bit-exact at env {bad[0]}: {a[ok][bad[0]]!r} vs {b[ok][bad[0]]!r}"
    a, b = gs["sqn"], osum["sqn"]
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{tag}: sqn NaN pattern"
    ok = ~np.isnan(a)
    np.testing.assert_allclose(a[ok], b[ok], rtol=1e-9, atol=1e-12, err_msg=f"{tag}: sqn")
