"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) prints
one JSON line with the keys the driver reads, and rank > 0 of a torchrun launch prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1", "--cpu-sample-elements", "1500", "--cal-elements", "1500",
                           "--elements", "4000", "--ref-budget-s", "120"], env=env, capture_output=True,
                          text=True, timeout=300)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "eigensolves/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["steps_timed"] == 1 and d["vs_baseline"] is None
    # the arm runs the SAME configuration as ours when it fits the budget and says so either way
    assert d["same_config"] is True and d["config"]["same_config"] is True
    assert abs(d["config"]["elements"] - 4000) <= 0.02 * 4000 and d["config"]["calibration"]["seconds"] > 0
    assert d["cpu_baseline"]["host"]["cpu_count"] == os.cpu_count() and "OMP_NUM_THREADS" in d["cpu_baseline"]["host"]
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-9 * d["value"]
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_falls_back_to_a_bounded_sample_and_says_so():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample-elements", "1500", "--cal-elements", "1500", "--elements", "1000000",
                        "--ref-budget-s", "0.5"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["same_config"] is False and d["config"]["elements"] < 100000
    assert d["extrapolated_full_size"]["seconds"] > d["ms_per_step"] / 1e3 and "NOT a measurement" in d["extrapolated_full_size"]["basis"]
