"""bench.py's JSON contract, checked on the CPU arm (the GPU arm prints the same keys plus roofline / clocks / gpu_launches)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "30000", "--steps", "1",
                          "--warmup", "1", "--ref-seconds", "0.5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vectors/s" and d["higher_is_better"] is True
    assert d["metric"] and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert set(d["config"]) >= {"workload", "rows", "cols", "k", "metric"} and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "query rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None
    # both arms print the same `config` object (the driver compares them): bench.config_dict is the only producer
    import bench
    wl = dict(bench.WORKLOADS["c2"], rows=30000)
    wl["name"] += " (rows overridden to 30000)"
    assert d["config"] == bench.config_dict(wl)


def test_reference_arm_runs_small_configs_in_full(oracle):
    """Config c1 is the one the CPU runs whole: nothing sampled, nothing extrapolated."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--rows", "1500",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert "WHOLE build" in d["cpu_baseline"]["sample"] and d["config"]["rows"] == 1500
    # value is rows / wall time of the step: no extrapolation
    assert abs(d["value"] - 1500 / (d["ms_per_step"] * 1e-3)) / d["value"] < 0.05


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "20000"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
