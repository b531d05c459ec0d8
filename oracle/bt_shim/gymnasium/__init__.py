"""
oracle/bt_shim/gymnasium -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Stand-in for the (uninstalled, unpinned) ``gymnasium`` dependency of
/root/reference/app/env.py:20-26, which uses it only for ``gym.Env`` and the
``spaces`` containers (no arithmetic).  Lets the reference env import here.
"""
from __future__ import annotations

import numpy as np

from . import spaces  # noqa: F401


class Env:
    metadata = {"render_modes": []}
    action_space = None
    observation_space = None
    np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)
        return None

    def step(self, action):  # pragma: no cover
        raise NotImplementedError

    def render(self):  # pragma: no cover
        return None

    def close(self):
        pass
