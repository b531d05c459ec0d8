"""Action / observation space containers.  Uses gymnasium's when it is installed (the reference depends on it,
app/env.py:20-26); otherwise minimal stand-ins with the same attributes (shape, dtype, low/high, n, sample,
contains) so that agent code reading `env.action_space` / `env.observation_space` keeps working."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installed in the build image
    import gymnasium as _gym
    from gymnasium.spaces import Box, Dict, Discrete  # noqa: F401

    EnvBase = _gym.Env
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class EnvBase:  # noqa: D401
        metadata = {"render_modes": []}

        def reset(self, *, seed=None, options=None):
            return None

    class _Space:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = shape, (np.dtype(dtype) if dtype is not None else None)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n
            except Exception:
                return False

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__(tuple(shape if shape is not None else np.shape(low)), dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Dict(_Space):
        def __init__(self, spaces):
            super().__init__(None, None)
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())
