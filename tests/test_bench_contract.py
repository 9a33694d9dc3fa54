"""The one-JSON-line contract of bench.py: the reference arm is run here (CPU only, one short step per core); the
CUDA arm is checked on the lines that the GPU runs of this round left in profiles/."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "bo_iterations_per_sec" and d["unit"] == "it/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_cuda_arm_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_5*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*gpu_1*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*gpu_2*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*gpu_4*.json")))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d), f
        assert d["metric"] == "bo_iterations_per_sec" and d["scaling"] == "weak" and d["dtype"] == "f64"
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == [] and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
