"""CPU (-m "not gpu") pre-flight of the PRODUCT's scalar core (gym_fx_b200/csrc/fx_core.cuh compiled with g++ by
tests/hostsim) against every golden trajectory: integer state and fp64 equity/commission bit-exact, rewards to
1e-12.  The CUDA kernel's warp-parallel parts are covered only by the -m gpu tests."""
import numpy as np
import pytest

from common import config_from_meta, golden_names, load_golden
from hostsim.hostsim import HostSimEnv


@pytest.mark.parametrize("name", golden_names())
def test_core_matches_golden(name):
    g = load_golden(name)
    cfg = config_from_meta(g["meta"], order_capacity=512)
    env = HostSimEnv(cfg, g["candles"], g["minutes"])
    env.reset(0)
    n = g["reward"].shape[0]
    lay_scal = g["obs"].shape[1] - 4
    has_agent = not (g["meta"]["config"].get("include_agent_state") is False)
    rows = {int(r): i for i, r in enumerate(g["obs_rows"])}
    for k in range(n):
        if k > 0:
            r, t = env.step(g["actions"][k - 1])
            assert t == g["terminated"][k], (name, k)
            np.testing.assert_allclose(r, g["reward"][k], rtol=1e-12, atol=1e-15, err_msg=f"{name} reward row {k}")
        inf = env.info()
        assert not (inf["flags"] & 16), f"{name}: order table overflow at row {k}"
        for key in ("position", "bar_index", "trades"):
            assert inf[key] == g[key][k], (name, key, k, inf[key], g[key][k])
        for key in ("equity", "price", "commission_paid"):
            assert inf[key] == g[key][k], (name, key, k, repr(inf[key]), repr(g[key][k]))
        if has_agent and k in rows:
            np.testing.assert_allclose(env.scalars(), g["obs"][rows[k], lay_scal:], rtol=1e-6, atol=1e-7,
                                       err_msg=f"{name} agent scalars row {k}")


def test_entry_fill_select_form_matches_branchy_rules():
    """fx_entry_fill (what every lane of the CUDA kernel evaluates) == fx_entry_hits + fx_match_limit / fx_match_stop."""
    import ctypes as C
    from hostsim.hostsim import lib
    L = lib()
    L.hs_check_entry_fill.argtypes = [C.c_int, C.c_uint]
    L.hs_check_entry_fill.restype = C.c_int
    for seed in (1, 2, 3):
        assert L.hs_check_entry_fill(200000, seed) == 0
