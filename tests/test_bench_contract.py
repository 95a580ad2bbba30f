"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints exactly ONE JSON line on stdout
with the keys of the contract -- the GPU arm shares `emit()` and `workload_config()` with it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports; the arm must still use every core
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert "workload" in d["config"] and "n=10" in d["config"]["workload"] and "N=6" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_episode_legs_warm_up_at_full_size_and_report_the_median():
    """bench._episodes: one untimed run first (idle clocks, cold allocator), then the median of the timed runs, each
    bracketed by the caller's sync and reduced over the ranks by the caller's reduction."""
    import importlib
    import time as _time
    bench = importlib.import_module("bench")
    calls, syncs = [], []
    durations = iter([0.0, 0.03, 0.01, 0.02])               # warm-up, then three timed runs

    def run():
        calls.append(len(calls))
        _time.sleep(next(durations))
        return len(calls)

    out, med, runs = bench._episodes(run, reps=3, sync=lambda: syncs.append(1), reduce=lambda d: round(d, 2))
    assert calls == [0, 1, 2, 3] and out == 4               # warm-up + 3 timed; the last result is returned
    assert len(syncs) == 3 and len(runs) == 3
    assert med == sorted(runs)[1] and min(runs) <= med <= max(runs)
