"""`gym_fx.env.GymFxEnv` import path of the reference, served by gym_fx_b200.env."""
from gym_fx_b200.env import GymFxEnv  # noqa: F401
