"""bench.py contract guards that need no GPU: the committed bench lines carry every key the driver reads, and every kernel family
that ever dominated a step has an algorithmic work model (so `roofline` is never silently empty)."""
import importlib.util
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def load_bench():
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def lines():
    for p in sorted((ROOT / "profiles").glob("r01z_bench_*.json")):
        yield p.name, json.loads(p.read_text().strip().splitlines()[-1])


@pytest.mark.parametrize("name,line", list(lines()))
def test_committed_bench_lines_have_the_contract_keys(name, line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in line, (name, k)
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] != line["value"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["roofline"]["bound"] in ("hbm", "tensor") and 0 < line["roofline"]["frac"] <= 1.0
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_every_dominant_kernel_family_has_a_work_model():
    bench = load_bench()
    peaks = bench.load_peaks()
    for name, line in lines():
        B = 44 if "c3" in name else 64
        top = sorted(line["kernel_breakdown"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:5]
        for label, v in top:
            assert bench.kernel_work(label, B) is not None, (name, label)
            r = bench.make_roofline(label, {label: v["ms_per_step"]}, {label: v["launches"]}, line["ms_per_step"], peaks, B)
            assert r["bound"] in ("hbm", "tensor") and r["achieved"] > 0
