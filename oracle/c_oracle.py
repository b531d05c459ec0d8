"""
oracle/c_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper around oracle/_build/libfxoracle.so (oracle/fxenv_oracle.c): N independent scalar CPU envs
stepped in lockstep, with the same array-in/array-out shape as the GPU VecFxEnv so parity tests can diff them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from gym_fx_b200.config import FxConfig, obs_dim

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfxoracle.so")
_lib = None


class FxoInfo(C.Structure):
    _fields_ = [
        ("equity", C.c_double), ("prev_equity", C.c_double), ("price", C.c_double), ("cash", C.c_double),
        ("position_size", C.c_double), ("position_price", C.c_double), ("commission_paid", C.c_double),
        ("position", C.c_int32), ("bar_index", C.c_int32), ("total_bars", C.c_int32), ("trades", C.c_int32),
        ("n_orders", C.c_int32), ("flags", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if the .so is missing or stale."""
    src = os.path.join(_HERE, "fxenv_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "fxenv.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.fxo_create.restype = C.c_void_p
        L.fxo_create.argtypes = [C.POINTER(FxConfig), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.fxo_destroy.argtypes = [C.c_void_p]
        L.fxo_reset.argtypes = [C.c_void_p, C.c_int64]
        L.fxo_step.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
        L.fxo_observe.argtypes = [C.c_void_p, C.c_void_p]
        L.fxo_info.argtypes = [C.c_void_p, C.POINTER(FxoInfo)]
        L.fxo_max_live_orders.argtypes = [C.c_void_p]
        L.fxo_summary.argtypes = [C.c_void_p, C.c_void_p]
        L.fxo_obs_dim.restype = C.c_int64
        L.fxo_obs_dim.argtypes = [C.POINTER(FxConfig)]
        L.fxo_step_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.fxo_step_batch_mt.argtypes = L.fxo_step_batch.argtypes + [C.c_int]
        L.fxo_run_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int]
        assert L.fxo_config_size() == C.sizeof(FxConfig), "FxConfig layout mismatch (python vs C)"
        assert L.fxo_info_size() == C.sizeof(FxoInfo)
        _lib = L
    return _lib


class OracleVec:
    """N scalar oracle envs. candles: list (one per pair) of float64 [T, n_cols]; env i uses pair i % num_pairs."""

    INFO_FIELDS = [f[0] for f in FxoInfo._fields_]

    def __init__(self, cfg: FxConfig, candles: Sequence[np.ndarray], minutes: Optional[Sequence[np.ndarray]] = None):
        self.L = lib()
        self.cfg = cfg
        self.N = int(cfg.num_envs)
        self.D = int(obs_dim(cfg))
        assert self.L.fxo_obs_dim(C.byref(cfg)) == self.D
        self.candles = [np.ascontiguousarray(c, dtype=np.float64) for c in candles]
        assert len(self.candles) == cfg.num_pairs
        for c in self.candles:
            assert c.ndim == 2 and c.shape[1] == cfg.n_cols
        self.minutes = None if minutes is None else [np.ascontiguousarray(m, dtype=np.int64) for m in minutes]
        self.envs = []
        for i in range(self.N):
            p = i % cfg.num_pairs
            mptr = None if self.minutes is None else self.minutes[p].ctypes.data
            h = self.L.fxo_create(C.byref(cfg), p, self.candles[p].ctypes.data, self.candles[p].shape[0], mptr)
            self.envs.append(h)
        self._arr = (C.c_void_p * self.N)(*self.envs)
        self.start = np.zeros(self.N, np.int64)

    def close(self):
        for h in self.envs:
            self.L.fxo_destroy(h)
        self.envs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, start_bars=None, mask=None):
        if start_bars is not None:
            self.start = np.asarray(start_bars, np.int64).copy()
        for i, h in enumerate(self.envs):
            if mask is None or mask[i]:
                self.L.fxo_reset(h, int(self.start[i]))
        return self.observe()

    def observe(self):
        obs = np.empty((self.N, self.D), np.float32)
        for i, h in enumerate(self.envs):
            self.L.fxo_observe(h, obs[i].ctypes.data)
        return obs

    def step(self, actions, want_obs=True):
        a = np.asarray(actions)
        obs = np.empty((self.N, self.D), np.float32) if want_obs else None
        rew = np.empty(self.N, np.float32)
        rew64 = np.empty(self.N, np.float64)
        term = np.empty(self.N, np.uint8)
        ai = af = None
        if self.cfg.action_mode == 1:
            af = np.ascontiguousarray(a, np.float32)
        else:
            ai = np.ascontiguousarray(a, np.int32)
        self.L.fxo_step_batch(self._arr, self.N, None if ai is None else ai.ctypes.data,
                              None if af is None else af.ctypes.data,
                              None if obs is None else obs.ctypes.data, self.D,
                              rew.ctypes.data, rew64.ctypes.data, term.ctypes.data)
        return obs, rew, rew64, term

    def info(self):
        out = {k: [] for k in self.INFO_FIELDS}
        inf = FxoInfo()
        for h in self.envs:
            self.L.fxo_info(h, C.byref(inf))
            for k in self.INFO_FIELDS:
                out[k].append(getattr(inf, k))
        dt = {"position": np.int32, "bar_index": np.int32, "total_bars": np.int32, "trades": np.int32,
              "n_orders": np.int32, "flags": np.uint32}
        return {k: np.asarray(v, dt.get(k, np.float64)) for k, v in out.items()}

    SUMMARY_FIELDS = ("max_drawdown_pct", "max_drawdown_money", "trades_total", "trades_won", "trades_lost",
                      "avg_trade_pnl", "sqn", "trades_closed", "open_trade_size", "open_trade_price", "open_trade_pnl",
                      "open_trade_commission")

    def summary(self):
        """Analyzer-derived fields of metrics_plugins/default_metrics.py:48-60, one array per field (NaN = None)."""
        out = np.zeros((self.N, len(self.SUMMARY_FIELDS)))
        for i, h in enumerate(self.envs):
            self.L.fxo_summary(h, out[i].ctypes.data)
        return {k: out[:, j].copy() for j, k in enumerate(self.SUMMARY_FIELDS)}

    def max_live_orders(self):
        return max(self.L.fxo_max_live_orders(h) for h in self.envs)


class ParallelStepper:
    """Steps an OracleVec with `threads` host threads (pthreads inside fxo_step_batch_mt, contiguous env slices).
    Used by bench.py's cpu_baseline / --impl reference legs only."""

    def __init__(self, vec: OracleVec, threads: int):
        self.vec = vec
        self.threads = max(1, min(int(threads), vec.N, 1024))
        self.obs = np.empty((vec.N, vec.D), np.float32)
        self.rew = np.empty(vec.N, np.float32)
        self.term = np.empty(vec.N, np.uint8)

    def step(self, actions: np.ndarray):
        v = self.vec
        is_f = v.cfg.action_mode == 1
        v.L.fxo_step_batch_mt(v._arr, v.N, None if is_f else actions.ctypes.data, actions.ctypes.data if is_f else None,
                              self.obs.ctypes.data, v.D, self.rew.ctypes.data, None, self.term.ctypes.data,
                              self.threads)
        return self.obs, self.rew, self.term

    def run(self, actions: np.ndarray):
        """K = actions.shape[0] steps of every env, no per-step barrier between threads (throughput baseline)."""
        v = self.vec
        is_f = v.cfg.action_mode == 1
        a = np.ascontiguousarray(actions, np.float32 if is_f else np.int32)
        v.L.fxo_run_mt(v._arr, v.N, a.shape[0], None if is_f else a.ctypes.data, a.ctypes.data if is_f else None,
                       self.obs.ctypes.data, v.D, self.rew.ctypes.data, self.term.ctypes.data, self.threads)
        return self.obs, self.rew, self.term

    def close(self):
        pass
