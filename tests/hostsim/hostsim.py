"""ctypes wrapper of tests/hostsim (g++ build of the product's fx_core.cuh) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gym_fx_b200.config import FxConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(os.path.join(_HERE, "_build", "libfxhostsim.so"))
        L.hs_create.restype = C.c_void_p
        L.hs_create.argtypes = [C.POINTER(FxConfig), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.hs_destroy.argtypes = [C.c_void_p]
        L.hs_reset.argtypes = [C.c_void_p, C.c_int64]
        L.hs_step.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
        L.hs_scalars.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class HostSimEnv:
    """Single env; OracleVec-like surface (reset/step/info) minus the observation windows."""

    def __init__(self, cfg: FxConfig, candles: np.ndarray, minutes=None):
        self.L = lib()
        self.candles = np.ascontiguousarray(candles, np.float64)
        self.minutes = None if minutes is None else np.ascontiguousarray(minutes, np.int64)
        self.h = self.L.hs_create(C.byref(cfg), 0, self.candles.ctypes.data, self.candles.shape[0],
                                  None if self.minutes is None else self.minutes.ctypes.data)

    def __del__(self):
        try:
            self.L.hs_destroy(self.h)
        except Exception:
            pass

    def reset(self, start=0):
        self.L.hs_reset(self.h, int(start))

    def step(self, action):
        r, t = C.c_double(), C.c_uint8()
        self.L.hs_step(self.h, float(action), C.byref(r), C.byref(t))
        return r.value, t.value

    def scalars(self):
        out = np.zeros(4, np.float32)
        self.L.hs_scalars(self.h, out.ctypes.data)
        return out

    def info(self):
        d, i, f = np.zeros(7), np.zeros(5, np.int32), np.zeros(1, np.uint32)
        self.L.hs_info(self.h, d.ctypes.data, i.ctypes.data, f.ctypes.data)
        return dict(equity=d[0], prev_equity=d[1], price=d[2], cash=d[3], position_size=d[4], position_price=d[5],
                    commission_paid=d[6], position=int(i[0]), bar_index=int(i[1]), total_bars=int(i[2]),
                    trades=int(i[3]), n_orders=int(i[4]), flags=int(f[0]))

    def stats(self):
        """The FX_RS_* record (fx_core.cuh): DrawDown / TradeAnalyzer / SQN state."""
        out = np.zeros(12)
        self.L.hs_stats(self.h, out.ctypes.data)
        return out
