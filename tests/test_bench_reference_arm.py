"""bench.py --impl reference runs on the host cores only, so its contract is checked on the CPU box: one JSON line on
stdout with the same metric / unit / higher_is_better / config.workload as the GPU arm prints, the cpu_baseline and
e2e blocks of the tier contract, and silence (exit 0) on every rank but 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(env_extra=None, *args):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--ks", "3,64", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_line(oracle_mod):
    sys.path.insert(0, ROOT)
    import bench
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["higher_is_better"] is True
    assert d["metric"] == bench.metric_name(28, [3, 64])                       # the GPU arm's metric, verbatim
    assert d["config"]["workload"] == bench.workload_name(28, 1, [3, 64])      # and its workload
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "i16" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29991"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
