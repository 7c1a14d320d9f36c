"""The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the bench
contract; the reference arm's line is produced live here on a tiny CPU sample and checked likewise."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r1k_bench_read_21M_n1.json", "r1i_bench_read_21M_n1.json",
                                  "r1i_bench_c2_1M_n1.json", "r1m_bench_read_21M_n8.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = _line(name)
    for k in BASE_KEYS:
        assert k in d, k
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["scaling"] in ("weak", "strong") and d["data"] == "synthetic" and d["steps"] >= 1 and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"].get("global_batch", 64) / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not (set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    assert d["gpu_launches"] > 0
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    else:
        assert d["cpu_baseline"] is None


def test_reference_arm_line_on_a_tiny_cpu_sample():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--rows", "20000", "--cpu-sample-rows", "10000", "--layers", "1",
                          "--retrieve-only"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
