"""The recorded bench lines under profiles/ carry every key of the bench.py contract (the files are the outputs of
`python bench.py` on the B200 box; this guards the reporting code against silently dropping a field)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOP = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
       "config", "e2e", "gpu_launches", "clocks", "roofline"]


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        pytest.skip("no recorded bench line")
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _latest("r01_bench_v9.json")
    for k in TOP + ["cpu_baseline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["data"] == "synthetic"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] < d["value"]  # host-to-host can not beat the device-resident number
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    c = d["clocks"]
    assert c["samples"] >= 1 and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == 2 * d["steps"]  # prologue + blind rotation per step
    for leg in ("tfhe_pbs", "ckks_mul", "ntt", "next_rows"):
        assert leg in d, leg


def test_multi_gpu_lines_scale_the_batch():
    for pattern, n in (("r01_bench_v9_2gpu.json", 2), ("r01_bench_v7_8gpu.json", 8)):
        d = _latest(pattern)
        for k in TOP:
            assert k in d, (pattern, k)
        assert d["n_gpus"] == n and d["config"]["global_batch"] == n * d["config"]["batch_per_gpu"]
