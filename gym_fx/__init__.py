"""`gym_fx` import alias kept for callers of the reference (`from gym_fx import GymFxEnv`): resolves to the
B200-native implementation in gym_fx_b200 and additionally exports the vectorised env."""


def __getattr__(name):
    if name in ("GymFxEnv", "VecFxEnv"):
        import gym_fx_b200
        return getattr(gym_fx_b200, name)
    raise AttributeError(name)


__all__ = ["GymFxEnv", "VecFxEnv"]
