"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores -- the one place outside tests/
where oracle/ is executed) honours the JSON contract: same metric / unit / config as the product arm, `impl`, `cpu_baseline`, `e2e`;
under torchrun only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, NTTB200_REF_STEP_BUDGET_S="0.2", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], env=env, capture_output=True,
                          text=True, timeout=300)


def test_reference_arm_prints_the_contract_line():
    o = _run({}, "--steps", "2", "--warmup", "1")
    assert o.returncode == 0, o.stderr[-500:]
    line = json.loads(o.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "60-bit NTT/s at N=2^15 batched" and line["unit"] == "NTT/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "NTT/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "forward NTTs" in cb["sample"]
    # the product arm describes its workload with the same function
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config(1)


def test_reference_arm_other_ranks_do_nothing():
    o = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert o.returncode == 0 and o.stdout.strip() == ""
