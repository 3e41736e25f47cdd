"""bench.py on the CPU: the algorithmic-byte model of SURVEY.md §8d, the synthetic generators, and the JSON line of
the reference arm's CPU fallback (the GPU arms need a device; their lines are checked by the driver)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_follow_the_survey_model():
    for name, expect in (("C2", 70004), ("C1", 11684), ("C3", 88484), ("C5", 66436)):
        w = dict(bench.WORKLOADS[name], update_method="sgd")                               # without the dense Adam terms
        n, R, dw, dd = w["n"], w["z"] + 1, w["dw"], w["dd"]
        assert 12 * (n * dw + R * dd) + 8 * (n + R) + 4 * (n + 1) == expect
        assert sum(bench.algorithmic_bytes_per_ngram(w).values()) == expect
    w = bench.WORKLOADS["C2"]
    dense = sum(bench.algorithmic_bytes_per_ngram(w).values()) - 70004
    assert abs(dense - 24.0 * (w["V"] * w["dw"] + w["D"] * w["dd"]) / w["B"]) < 1e-6   # full_adam: theta, m, v read + written


def test_synthetic_generators():
    w = dict(bench.WORKLOADS["C1"])
    a = bench.make_batches(w, 64, 7, 3)
    b = bench.make_batches(w, 64, 7, 3)
    assert len(a) == 3 and all((x[0] == y[0]).all() and (x[1] == y[1]).all() for x, y in zip(a, b))      # seeded
    assert a[0][0].shape == (64, w["n"]) and a[0][0].max() < w["V"] and a[0][1].max() < w["D"]
    wz = dict(w, word_zipf=1.0)
    f = np.concatenate([x[0].ravel() for x in bench.make_batches(wz, 4096, 7, 2)])
    counts = np.bincount(f, minlength=w["V"])
    assert f.max() < w["V"] and counts[0] > 5 * counts[50] > 0                           # Zipf(1): rank 1 vs rank 51


def test_reference_arm_cpu_fallback_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reference_kind", "cpu",
                          "--workload", "C1", "--steps", "1", "--warmup", "1", "--cpu_sample", "256"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "n-grams/sec" and line["unit"] == "n-grams/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("LSE tanh")


def test_reference_arm_runs_on_rank_zero_only():
    """Under torchrun the reference arm runs and prints on rank 0; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
