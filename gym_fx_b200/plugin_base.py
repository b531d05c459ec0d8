"""
Common base of the plugin mirrors.

The reference's plugins are duck-typed (no base class): class attribute `plugin_params`, `__init__(config=None)`,
`set_params(**kw)` (app/main.py:20-24 calls `klass(config)` then `set_params(**config)`; app/plugin_loader.py:40
reads `plugin_class.plugin_params`).  The mirrors keep that surface and the same parameter names/defaults, but the
per-tick methods the reference env calls (`apply_action`, `compute_reward`, `make_observation`) are NOT host code
here: their arithmetic runs inside the fused CUDA step kernel, selected by `gym_fx_b200.config.lower_config` from
the plugin's kind + params.  Calling them on the host raises `KernelResident`.
"""
from __future__ import annotations

from typing import Any, Dict


class KernelResident(NotImplementedError):
    """The requested per-tick computation exists only inside the CUDA step kernel (no host fallback)."""


def kernel_resident(what: str):
    raise KernelResident(
        f"{what} is computed inside the fused sm_100a step kernel of libfxenv.so; "
        "drive it through gym_fx_b200.GymFxEnv / VecFxEnv (there is deliberately no CPU fallback)"
    )


class PluginBase:
    plugin_params: Dict[str, Any] = {}
    #: True -> set_params only accepts keys already in plugin_params (direct_*_sltp behaviour)
    strict_keys = False

    def __init__(self, config: Dict[str, Any] | None = None):
        self.params = dict(self.plugin_params)
        if config:
            self.set_params(**config)

    def set_params(self, **kwargs: Any) -> None:
        if self.strict_keys:
            kwargs = {k: v for k, v in kwargs.items() if k in self.plugin_params}
        self.params.update(kwargs)
