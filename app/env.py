"""`app.env.GymFxEnv` (reference: app/env.py:31) -> gym_fx_b200.env.GymFxEnv."""
from gym_fx_b200.env import GymFxEnv  # noqa: F401

__all__ = ["GymFxEnv"]
