"""gym_fx_b200 -- B200-native vectorised gym-fx environment (drop-in for the reference's env.step() hot path).

    from gym_fx_b200 import GymFxEnv      # the reference's single-env Gym API (app/env.py)
    from gym_fx_b200 import VecFxEnv      # N envs per GPU, torch tensors, one fused kernel launch per step

The CUDA library (gym_fx_b200/libfxenv.so, C-ABI in include/fxenv.h) is built by `__graft_entry__.build()`.
"""
from .config import FxConfig, lower_config, obs_dim, obs_layout  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA
    if name == "VecFxEnv":
        from .vec_env import VecFxEnv
        return VecFxEnv
    if name == "GymFxEnv":
        from .env import GymFxEnv
        return GymFxEnv
    raise AttributeError(name)
