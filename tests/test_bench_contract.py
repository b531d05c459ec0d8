"""CPU (-m "not gpu"): the bench.py contract that can be checked without a GPU -- the reference arm prints exactly ONE
JSON line on stdout with the agreed keys, and the workload table / algorithmic-bytes formula match SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["workload"].startswith("cfg2")
    # both arms print the SAME config object (bench.common_config), including the episode phase the steps are taken from
    sys.path.insert(0, ROOT)
    import bench
    cfg, _, _, envs, D, _, desc = bench.build_workload("cfg2")
    pre = bench.preroll_steps(cfg)
    assert pre > max(cfg.window_size, cfg.scaling_window)
    assert d["config"] == bench.common_config(desc, envs, D, 1, pre) and "steady state" in d["config"]["episode_phase"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    want = {"cfg2": 3609, "cfg3": 7209, "cfg4": 4129, "cfg5": 14881}
    for name, (envs, W, strat, rew, pairs, R) in bench.WORKLOADS.items():
        assert 4 * (W * 5 + 2 * W + 4) + 4 + 1 + 4 + R == want[name], name
