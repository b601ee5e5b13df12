"""bench.py contract that can be checked without a GPU: the reference arm prints exactly one JSON line on stdout
with the keys the driver reads; the GPU arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    import time
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0",
                        "--cpu-genes", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "genes/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # ms_per_step is the wall time a step really took here (the driver holds steps x ms_per_step against its own clock);
    # the time the whole workload would take at `value` is a separate key
    assert 0 < d["steps"] * d["ms_per_step"] / 1000.0 <= wall
    assert d["ms_full_workload"] == pytest.approx(1000.0 * d["config"]["genes"] / d["value"])


def test_gpu_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--genes", "8"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_committed_traffic_record_matches_the_default_workload():
    """roofline.traffic is quoted from the ncu launch list of the headline command (profiles/r02_traffic_c3.json):
    it only applies while bench.py's default C3 sample is the workload that list was taken on."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    rec = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_c3.json")))
    assert rec["config"] == "c3" and rec["genes"] == bench.DEFAULT_GENES["c3"]
    assert rec["outer_iterations"] == bench.RUN_KW["degnorm_iter"]
    # DRAM traffic of the fused launches within a few per cent of the algorithmic bytes (nothing re-read)
    assert 0.9 < rec["dram_bytes_per_launch_group"] / 13.49e12 < 1.05
