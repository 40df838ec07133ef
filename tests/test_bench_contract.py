"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--envs-per-gpu", "2048",
                           "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 3 and line["warmup"] == 1
    assert line["config"]["env"] == "MiniGrid-DoorKey-8x8-v0" and line["config"]["envs_per_gpu"] == 2048
    # both arms describe the workload with the same dict (the driver compares them: same_config)
    import bench

    assert line["config"] == bench.make_config("MiniGrid-DoorKey-8x8-v0", 2048, 1)
    assert "desynchronised" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29555"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""
