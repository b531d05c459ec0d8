"""
VecFxEnv -- N lock-stepped gym-fx environments on one GPU, torch tensors in and out, no host round trip.

Semantics per env are those of the reference's GymFxEnv.reset/step (app/env.py:102-172); a step of all N envs is ONE
launch of the fused sm_100a kernel in libfxenv.so, a batch of K steps with known actions ONE persistent launch.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence

import numpy as np
import torch

from . import _native
from .config import FxConfig, obs_dim, obs_layout


class VecFxEnv:
    """
    cfg      FxConfig from gym_fx_b200.config.lower_config (num_envs, plugin selection and parameters)
    candles  one float64 [T, n_cols] array per currency pair (env i trades pair i % num_pairs)
    minutes  optional int64 [T] minutes-since-epoch per pair (ATR session filter)
    """

    def __init__(self, cfg: FxConfig, candles: Sequence[np.ndarray], minutes: Optional[Sequence[np.ndarray]] = None,
                 device: Optional[torch.device | str | int] = None):
        if not torch.cuda.is_available():
            raise _native.FxEnvError("VecFxEnv needs a CUDA device (libfxenv.so has no CPU path)")
        self.L = _native.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _native.FxEnvError("VecFxEnv device must be a CUDA device")
        self.cfg = cfg
        self.num_envs = int(cfg.num_envs)
        self.obs_dim = int(obs_dim(cfg))
        self.layout = obs_layout(cfg)
        if len(candles) != cfg.num_pairs:
            raise ValueError(f"expected {cfg.num_pairs} candle tables, got {len(candles)}")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.L.fxenv_create(C.byref(cfg), C.byref(self._h))
            _native.check(self.L, None, rc, "fxenv_create")
            assert self.L.fxenv_obs_dim(self._h) == self.obs_dim
            self.total_rows = []
            for p, tab in enumerate(candles):
                tab = np.ascontiguousarray(tab, dtype=np.float64)
                if tab.ndim != 2 or tab.shape[1] != cfg.n_cols:
                    raise ValueError(f"candle table {p} must be [T, {cfg.n_cols}]")
                m = None if minutes is None or minutes[p] is None else np.ascontiguousarray(minutes[p], np.int64)
                rc = self.L.fxenv_load_candles(self._h, p, tab.ctypes.data, tab.shape[0],
                                               None if m is None else m.ctypes.data)
                _native.check(self.L, self._h, rc, "fxenv_load_candles")
                self.total_rows.append(tab.shape[0])
        N, D, dev = self.num_envs, self.obs_dim, self.device
        self.obs = torch.empty((N, D), dtype=torch.float32, device=dev)
        self.reward = torch.empty(N, dtype=torch.float32, device=dev)
        self.reward64 = torch.empty(N, dtype=torch.float64, device=dev)
        self.terminated = torch.empty(N, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(N, dtype=torch.bool, device=dev)  # always False (app/env.py:158)
        self._info_ptrs = _native.FxInfoPtrs()
        self.L.fxenv_get_info(self._h, C.byref(self._info_ptrs))
        self._info_views: Dict[str, torch.Tensor] = {}
        self.action_dtype = torch.float32 if cfg.action_mode == 1 else torch.int32
        self._policies: list = []

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            # the info tensors are zero-copy views of the library's state slab, which fxenv_destroy frees: drop the cached
            # ones (views a caller still holds become invalid, like any tensor handed out by info() -- clone to keep)
            self._info_views.clear()
            for pol in list(getattr(self, "_policies", [])):   # policies hold device buffers tied to this handle
                pol.close()
            self.L.fxenv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ Gym-style API
    def reset(self, start_bars: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        """-> (obs [N, D] float32, info).  start_bars: int64 [N] first table row of each env's episode window."""
        sb = mk = None
        if start_bars is not None:
            sb = torch.as_tensor(start_bars, dtype=torch.int64, device=self.device).contiguous().reshape(-1)
            if sb.numel() != self.num_envs:
                raise ValueError(f"start_bars must have {self.num_envs} elements, got {sb.numel()}")
        if mask is not None:
            mk = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous().reshape(-1)
            if mk.numel() != self.num_envs:
                raise ValueError(f"mask must have {self.num_envs} elements, got {mk.numel()}")
        rc = self.L.fxenv_reset(self._h, None if sb is None else sb.data_ptr(), None if mk is None else mk.data_ptr(),
                                self._stream())
        _native.check(self.L, self._h, rc, "fxenv_reset")
        rc = self.L.fxenv_observe(self._h, self.obs.data_ptr(), self._stream())
        _native.check(self.L, self._h, rc, "fxenv_observe")
        return self.obs, self.info()

    def step(self, actions: torch.Tensor, out_obs: Optional[torch.Tensor] = None):
        """-> (obs, reward float32 [N], terminated bool [N], truncated bool [N], info).
        The returned tensors are views of buffers that the next step() overwrites."""
        a = actions
        if a.dtype != self.action_dtype or a.device != self.device or not a.is_contiguous():
            a = a.to(device=self.device, dtype=self.action_dtype).contiguous()
        a = a.reshape(-1)
        if a.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions")
        obs = self.obs
        if out_obs is not None:
            self._check("out_obs", out_obs, torch.float32, (self.num_envs, self.obs_dim))
            obs = out_obs
        rc = self.L.fxenv_step(self._h, a.data_ptr(), obs.data_ptr(), self.reward.data_ptr(),
                               self.terminated.data_ptr(), self.reward64.data_ptr(), self._stream())
        _native.check(self.L, self._h, rc, "fxenv_step")
        return obs, self.reward, self.terminated.view(torch.bool), self.truncated, self.info()

    # ---- argument checks of the raw-pointer calls: a wrong dtype / shape / device would be read as garbage or fault
    def _check(self, name: str, t: torch.Tensor, dtype: torch.dtype, shape, host: bool = False):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != dtype:
            raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        if host:
            if t.device.type != "cpu":
                raise ValueError(f"{name} must be a host (ideally pinned) tensor")
        elif t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")

    def _check_batch(self, actions, obs_ring, rewards, terminated):
        if actions.dim() != 2:
            raise ValueError("actions must be [K, num_envs]")
        K, N, D = int(actions.shape[0]), self.num_envs, self.obs_dim
        if K < 1:
            raise ValueError("step_many needs at least one step")
        self._check("actions", actions, self.action_dtype, (K, N))
        if obs_ring.dim() != 3 or obs_ring.shape[0] < 1:
            raise ValueError("obs_ring must be [slots, num_envs, obs_dim] with slots >= 1")
        self._check("obs_ring", obs_ring, torch.float32, (int(obs_ring.shape[0]), N, D))
        self._check("rewards", rewards, torch.float32, (K, N))
        self._check("terminated", terminated, torch.uint8, (K, N))
        return K

    def step_many(self, actions: torch.Tensor, obs_ring: torch.Tensor, rewards: torch.Tensor, terminated: torch.Tensor):
        """K consecutive steps with all actions supplied up front (replay / random / scripted drivers); identical results
        to K calls of step().  actions [K, N] (int32, or float32 in continuous mode); obs_ring float32 [slots, N, D]
        (step k writes slot k % slots); rewards float32 [K, N]; terminated uint8 [K, N].  The library runs the batch as one
        persistent launch or as a cached CUDA graph of single steps (`step_many_engine`)."""
        K = self._check_batch(actions, obs_ring, rewards, terminated)
        rc = self.L.fxenv_step_many(self._h, K, actions.data_ptr(), obs_ring.data_ptr(), int(obs_ring.shape[0]),
                                    rewards.data_ptr(), terminated.data_ptr(), self._stream())
        _native.check(self.L, self._h, rc, "fxenv_step_many")

    def plan_step_many(self, actions: torch.Tensor, obs_ring: torch.Tensor, rewards: torch.Tensor,
                       terminated: torch.Tensor):
        """Validate a step_many argument set ONCE and return a zero-argument callable that enqueues the batch on the
        current stream with a single C call (for loops that replay the same buffers: the checks and the Python attribute
        lookups stay out of the hot loop).  The tensors must stay alive and unchanged in place while the plan is used."""
        K = self._check_batch(actions, obs_ring, rewards, terminated)
        fn, h, check = self.L.fxenv_step_many, self._h, _native.check
        args = (K, actions.data_ptr(), obs_ring.data_ptr(), int(obs_ring.shape[0]), rewards.data_ptr(),
                terminated.data_ptr())
        keep = (actions, obs_ring, rewards, terminated)
        stream_of = self._stream

        def launch(_keep=keep):
            rc = fn(h, *args, stream_of())
            if rc:
                check(self.L, h, rc, "fxenv_step_many")

        return launch

    def step_host(self, actions_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  terminated_host: torch.Tensor):
        """Reference-facing call with (pinned) HOST tensors: H2D actions, one step, D2H results, synchronous."""
        N, D = self.num_envs, self.obs_dim
        self._check("actions_host", actions_host, self.action_dtype, (N,), host=True)
        self._check("obs_host", obs_host, torch.float32, (N, D), host=True)
        self._check("reward_host", reward_host, torch.float32, (N,), host=True)
        self._check("terminated_host", terminated_host, torch.uint8, (N,), host=True)
        rc = self.L.fxenv_step_host(self._h, actions_host.data_ptr(), obs_host.data_ptr(), reward_host.data_ptr(),
                                    terminated_host.data_ptr())
        _native.check(self.L, self._h, rc, "fxenv_step_host")

    # ------------------------------------------------------------------ info / state
    def _view(self, name: str) -> torch.Tensor:
        v = self._info_views.get(name)
        if v is None:
            ptr = getattr(self._info_ptrs, name)
            dt = getattr(torch, _native.INFO_DTYPES[name])
            v = _tensor_from_ptr(ptr, self.num_envs, dt, self.device)
            self._info_views[name] = v
        return v

    def info(self) -> Dict[str, torch.Tensor]:
        """Zero-copy views of the device-side info columns (app/env.py:244-254,162-166)."""
        return _LazyInfo(self)

    def obs_dict(self, obs: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Split flat rows into the reference's Dict observation (views)."""
        o = self.obs if obs is None else obs
        return {k: o[:, off:off + int(np.prod(shape))].reshape((o.shape[0],) + tuple(shape))
                for k, (off, shape) in self.layout.items()}

    def run_stats(self) -> torch.Tensor:
        """float64 [N, 12] view of the per-env analyzer state (include/fxenv.h FXENV_RS_*)."""
        v = self._info_views.get("run_stats")
        if v is None:
            v = _tensor_from_ptr(self._info_ptrs.run_stats, self.num_envs * _native.RUN_STATS, torch.float64, self.device)
            v = v.view(self.num_envs, _native.RUN_STATS)
            self._info_views["run_stats"] = v
        return v

    def summary(self) -> Dict[str, Any]:
        """Per-env end-of-run summary as device tensors, plus fleet aggregates: the fields of GymFxEnv.summary()
        (app/env.py:256-271 -> metrics_plugins/default_metrics.py:48-60).  The analyzer-derived ones (max drawdown,
        trades won / lost, average trade pnl, SQN) are tracked inside the step kernel -- what backtrader's DrawDown /
        TradeAnalyzer / SQN analyzers would report if the run ended now; NaN where the reference reports None.
        `sharpe_ratio` (SharpeRatio over calendar days) is not tracked and stays None."""
        i = self.info()
        R = _native.RS
        rs = self.run_stats()
        ic = float(self.cfg.initial_cash)
        eq = i["equity"]
        ret = eq / ic - 1.0 if ic else torch.zeros_like(eq)
        closed = i["trades"].double()
        nan = torch.full_like(closed, float("nan"))
        avg = torch.where(closed > 0, rs[:, R["pnl_net"]] / closed.clamp(min=1.0), nan)
        sqn = self._sqn(rs[:, R["pnl_net"]], rs[:, R["pnl_sq"]], closed)
        return {
            "initial_cash": ic, "final_equity": eq, "total_return": ret,
            "max_drawdown_pct": rs[:, R["dd_max_pct"]], "max_drawdown_money": rs[:, R["dd_max_money"]],
            "sharpe_ratio": None, "sqn": sqn,
            "trades_total": rs[:, R["opened"]].to(torch.int64), "trades_won": rs[:, R["won"]].to(torch.int64),
            "trades_lost": rs[:, R["lost"]].to(torch.int64), "trades_closed": i["trades"], "avg_trade_pnl": avg,
            "commission_paid": i["commission_paid"], "position": i["position"], "bar_index": i["bar_index"],
            "mean_total_return": float(ret.mean()), "min_total_return": float(ret.min()),
            "max_total_return": float(ret.max()), "mean_trades": float(closed.mean()),
            "worst_max_drawdown_pct": float(rs[:, R["dd_max_pct"]].max()),
            "fleet_win_rate": float(rs[:, R["won"]].sum() / closed.sum().clamp(min=1.0)),
            "order_overflow_envs": int((i["flags"] & 16).ne(0).sum()),
        }

    @staticmethod
    def _sqn(s1: torch.Tensor, s2: torch.Tensor, n: torch.Tensor) -> torch.Tensor:
        """backtrader's SQN = sqrt(n) * mean / population std of the closed trades' net pnl, from their sum and sum of
        squares; 0 for n <= 1 and NaN (the reference's None) when all trades had the same pnl (std == 0)."""
        nn = n.clamp(min=1.0)
        mean = s1 / nn
        ex2 = s2 / nn
        var = ex2 - mean * mean
        degenerate = var <= 1e-12 * ex2          # rounding noise of the subtraction, not a variance
        sd = torch.sqrt(var.clamp(min=0.0))
        sqn = torch.sqrt(nn) * mean / torch.where(degenerate, torch.ones_like(sd), sd)
        sqn = torch.where(degenerate, torch.full_like(sqn, float("nan")), sqn)
        return torch.where(n > 1, sqn, torch.zeros_like(sqn))

    def analyzers(self, env: int = 0) -> Dict[str, Any]:
        """The analyzer results of ONE env in the shape of backtrader's get_analysis() dicts, as GymFxEnv.summary()
        hands them to the metrics plugin (app/env.py:258-265): trades / drawdown / sqn / sharpe / time_return."""
        R = _native.RS
        rs = self.run_stats()[env].cpu().numpy()
        closed = int(self.info()["trades"][env])
        trades: Dict[str, Any] = {"total": {"total": int(rs[R["opened"]]), "open": int(rs[R["opened"]]) - closed,
                                            "closed": closed}}
        if closed:
            trades.update(won={"total": int(rs[R["won"]])}, lost={"total": int(rs[R["lost"]])},
                          pnl={"net": {"total": float(rs[R["pnl_net"]]), "average": float(rs[R["pnl_net"]]) / closed}})
        if closed > 1:
            v = float(self._sqn(torch.tensor([rs[R["pnl_net"]]], dtype=torch.float64), torch.tensor([rs[R["pnl_sq"]]], dtype=torch.float64),
                                torch.tensor([float(closed)], dtype=torch.float64))[0])
            sqn = None if v != v else v
        else:
            sqn = 0
        return {"trades": trades, "sqn": {"sqn": sqn, "trades": closed}, "sharpe": {}, "time_return": {},
                "drawdown": {"max": {"drawdown": float(rs[R["dd_max_pct"]]), "moneydown": float(rs[R["dd_max_money"]])}}}

    # ------------------------------------------------------------------ closed loop (policy on the device)
    def make_policy(self, weights=None) -> "FusedPolicy":
        """An actor-critic MLP(obs_dim, 256, 256) evaluated by the fused tensor-core kernel (fxenv.h: FxPolicy)."""
        return FusedPolicy(self, weights)

    def rollout(self, policy: "FusedPolicy", horizon: int, buffers: Optional[Dict[str, torch.Tensor]] = None,
                gumbel: Optional[torch.Tensor] = None, seed: int = 0) -> Dict[str, torch.Tensor]:
        """`horizon` closed-loop steps (policy -> sample -> env.step) from the current state, all on the device: the
        loop of the reference's driver (app/main.py:57-65) with a learned policy.  Returns / fills `buffers`:
        obs [H+1, N, D] (obs[t] is what the policy saw at step t), actions int32 [H, N], logp [H, N], value [H+1, N]
        (value[H] bootstraps), reward [H, N], done uint8 [H, N].  gumbel: optional float32 [H, N, 3] Gumbel(0,1) noise
        (reproducible sampling); default: the kernel's counter-based generator seeded with `seed`."""
        H, N, D, dev = int(horizon), self.num_envs, self.obs_dim, self.device
        b = buffers if buffers is not None else {}
        want = {"obs": ((H + 1, N, D), torch.float32), "actions": ((H, N), torch.int32), "logp": ((H, N), torch.float32),
                "value": ((H + 1, N), torch.float32), "reward": ((H, N), torch.float32), "done": ((H, N), torch.uint8)}
        for k, (shape, dt) in want.items():
            if k not in b:
                b[k] = torch.empty(shape, dtype=dt, device=dev)
            elif k == "obs":
                if b[k].dim() != 3 or b[k].shape[0] < 2:
                    raise ValueError("obs must be [slots >= 2, num_envs, obs_dim]")
                self._check("obs", b[k], dt, (int(b[k].shape[0]), N, D))
            else:
                self._check(k, b[k], dt, shape)
        if gumbel is not None:
            self._check("gumbel", gumbel, torch.float32, (H, N, 3))
        io = _native.FxRollout(H, int(b["obs"].shape[0]), b["obs"].data_ptr(), b["actions"].data_ptr(), b["logp"].data_ptr(),
                               b["value"].data_ptr(), b["reward"].data_ptr(), b["done"].data_ptr(),
                               0 if gumbel is None else gumbel.data_ptr(), int(seed) & (2**64 - 1))
        rc = self.L.fxenv_rollout(self._h, policy._p, C.byref(io), self._stream())
        _native.check(self.L, self._h, rc, "fxenv_rollout")
        policy._keep = (b, gumbel)
        return b

    def launch_count(self) -> int:
        return int(self.L.fxenv_launch_count(self._h))

    def step_many_engine(self, n_steps: int) -> str:
        """'persistent' (one launch, per-env dependencies) or 'graph' (CUDA graph of single steps) -- see fxenv.h."""
        return "persistent" if int(self.L.fxenv_step_many_engine(self._h, int(n_steps))) == 1 else "graph"

    def get_state(self) -> bytes:
        n = self.L.fxenv_state_bytes(self._h)
        buf = (C.c_char * n)()
        _native.check(self.L, self._h, self.L.fxenv_get_state(self._h, buf, n), "fxenv_get_state")
        return bytes(buf)

    def set_state(self, blob: bytes):
        _native.check(self.L, self._h, self.L.fxenv_set_state(self._h, blob, len(blob)), "fxenv_set_state")


class _LazyInfo(dict):
    """dict of info tensors materialised on first access (keeps step() free of Python overhead)."""

    KEYS = tuple(_native.INFO_DTYPES)

    def __init__(self, env: VecFxEnv):
        super().__init__()
        self._env = env

    def __missing__(self, key):
        if key == "pnl":
            v = self._env._view("equity") - self._env._view("prev_equity")
        elif key == "reward":
            v = self._env.reward
        elif key in self.KEYS:
            v = self._env._view(key)
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def keys(self):
        return list(self.KEYS) + ["pnl", "reward"]

    def __contains__(self, key):
        return key in self.KEYS or key in ("pnl", "reward")


class _CudaArrayView:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def _tensor_from_ptr(ptr: int, n: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArrayView(ptr, n, typestr), device=device)


class FusedPolicy:
    """Device-resident actor-critic MLP(obs_dim, 256, 256) -> 3 logits + value, evaluated between env steps by the fused
    tcgen05 kernel (gym_fx_b200/csrc/fx_policy.cu).  `set_weights` takes float32 CUDA tensors in torch.nn.Linear layout
    (or a module with .body[0], .body[2], .pi, .v like examples/ppo_rollout.py's ActorCritic)."""

    HIDDEN = 256

    def __init__(self, env: VecFxEnv, weights=None):
        self.env = env
        self._p = C.c_void_p()
        rc = env.L.fxenv_policy_create(env._h, C.byref(self._p))
        _native.check(env.L, env._h, rc, "fxenv_policy_create")
        self._keep = None
        env._policies.append(self)
        if weights is not None:
            self.set_weights(weights)

    def sync_timeouts(self) -> int:
        """Polls of the last rollout's tile hand-over that gave up (0 unless something is broken); synchronises."""
        rc = int(self.env.L.fxenv_policy_sync_timeouts(self._p))
        if rc < 0:
            _native.check(self.env.L, self.env._h, rc, "fxenv_policy_sync_timeouts")
        return rc

    def set_weights(self, weights):
        if not isinstance(weights, dict):
            m = weights
            weights = {"w1": m.body[0].weight, "b1": m.body[0].bias, "w2": m.body[2].weight, "b2": m.body[2].bias,
                       "w_pi": m.pi.weight, "b_pi": m.pi.bias, "w_v": m.v.weight, "b_v": m.v.bias}
        D, Hd = self.env.obs_dim, self.HIDDEN
        shapes = {"w1": (Hd, D), "b1": (Hd,), "w2": (Hd, Hd), "b2": (Hd,), "w_pi": (3, Hd), "b_pi": (3,), "w_v": (Hd,), "b_v": (1,)}
        ts = {}
        for k, shape in shapes.items():
            t = weights[k].detach()
            if k == "w_v":
                t = t.reshape(-1)
            t = t.to(device=self.env.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"{k} must have shape {shape}, got {tuple(t.shape)}")
            ts[k] = t
        w = _native.FxPolicyWeights(*[ts[k].data_ptr() for k in ("w1", "b1", "w2", "b2", "w_pi", "b_pi", "w_v", "b_v")])
        rc = self.env.L.fxenv_policy_set_weights(self._p, C.byref(w), self.env._stream())
        _native.check(self.env.L, self.env._h, rc, "fxenv_policy_set_weights")
        self._w = ts  # keep the sources alive until the stream-ordered copies have run

    def close(self):
        if self._p:
            self.env.L.fxenv_policy_destroy(self._p)
            self._p = C.c_void_p()
            if self in self.env._policies:
                self.env._policies.remove(self)

    def __del__(self):
        try:
            if self.env._h:
                self.close()
        except Exception:
            pass
