"""`app.plugin_loader` (reference: app/plugin_loader.py:12-83): same two functions, same return contract."""
from gym_fx_b200.plugin_loader import get_plugin_params, load_plugin  # noqa: F401

__all__ = ["load_plugin", "get_plugin_params"]
