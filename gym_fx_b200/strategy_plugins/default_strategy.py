"""default_strategy mirror (reference: strategy_plugins/default_strategy.py:19-54): the diagnostic DRIVER
(buy_hold | random | flat | replay).  It has no `apply_action`, so the env uses the default market-order flow
(app/bt_bridge.py:171-190).  `decide_action` is host logic and is implemented here."""
from __future__ import annotations

import csv
import random

from ..plugin_base import PluginBase


class Plugin(PluginBase):
    plugin_kind = "default_strategy"
    plugin_params = {"driver_mode": "buy_hold", "replay_actions_file": None, "seed": None}

    def __init__(self, config=None):
        self._replay = []
        self._rng = random.Random()
        super().__init__(config)

    def set_params(self, **kwargs):
        super().set_params(**kwargs)
        if self.params.get("seed") is not None:
            self._rng = random.Random(self.params["seed"])
        path = self.params.get("replay_actions_file")
        if path:
            with open(path, "r", encoding="utf-8") as fh:
                self._replay = [int(r.get("action", 0)) for r in csv.DictReader(fh)]

    def decide_action(self, obs, info, step: int) -> int:
        mode = self.params.get("driver_mode", "buy_hold")
        if mode == "flat":
            return 0
        if mode == "random":
            return self._rng.choice([0, 1, 2])
        if mode == "replay":
            return self._replay[step] if step < len(self._replay) else 0
        return 1 if step == 0 else 0
