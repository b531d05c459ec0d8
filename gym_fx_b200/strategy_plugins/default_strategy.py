"""default_strategy mirror -- the diagnostic action DRIVER of the reference (strategy_plugins/default_strategy.py:19-54).

It defines no `apply_action`, so the env runs the default market-order flow (app/bt_bridge.py:171-190) in the kernel.
Choosing actions is host logic and lives here.  Besides the reference's per-step `decide_action`, `action_table`
materialises a whole [steps, num_envs] action block up front -- the natural input of `VecFxEnv.step_many`, which then
runs the batch without coming back to the host."""
from __future__ import annotations

import csv
import random
from typing import Callable, Dict, List

import numpy as np

from ..plugin_base import PluginBase

HOLD, LONG, SHORT = 0, 1, 2


def _load_action_column(path: str) -> List[int]:
    """CSV with an `action` column (missing cells read as hold), one row per step."""
    with open(path, "r", encoding="utf-8", newline="") as fh:
        return [int(row.get("action", HOLD)) for row in csv.DictReader(fh)]


class Plugin(PluginBase):
    plugin_kind = "default_strategy"
    plugin_params = {"driver_mode": "buy_hold", "replay_actions_file": None, "seed": None}

    def __init__(self, config=None):
        self._script: List[int] = []
        self._rng = random.Random()
        # mode -> rule(step); anything unknown behaves like buy_hold, as in the reference
        self._rules: Dict[str, Callable[[int], int]] = {
            "flat": lambda step: HOLD,
            "random": lambda step: self._rng.choice([HOLD, LONG, SHORT]),
            "replay": lambda step: self._script[step] if 0 <= step < len(self._script) else HOLD,
            "buy_hold": lambda step: LONG if step == 0 else HOLD,
        }
        super().__init__(config)

    def set_params(self, **kwargs):
        super().set_params(**kwargs)
        seed = self.params.get("seed")
        if seed is not None:
            self._rng = random.Random(seed)   # same generator and draw as the reference => same action stream
        source = self.params.get("replay_actions_file")
        if source:
            self._script = _load_action_column(source)

    def decide_action(self, obs, info, step: int) -> int:
        rule = self._rules.get(self.params.get("driver_mode", "buy_hold"), self._rules["buy_hold"])
        return int(rule(int(step)))

    def action_table(self, steps: int, num_envs: int = 1) -> np.ndarray:
        """int32 [steps, num_envs]: what `decide_action` would return step by step, the same stream for every env
        (the random driver consumes its generator exactly as `steps` calls of decide_action would)."""
        column = np.fromiter((self.decide_action(None, None, k) for k in range(int(steps))), dtype=np.int32, count=int(steps))
        return np.ascontiguousarray(np.repeat(column[:, None], int(num_envs), axis=1))
