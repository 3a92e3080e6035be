"""bench.py's output contract (one JSON line with the keys the driver reads), on a short run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(args, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference needs no GPU: the reference's own CPU path on a small bounded sample"""
    d = _line(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", str(1 << 18)])
    assert d["impl"] == "reference" and d["unit"] == "Gk-mer/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("cfg2")


@pytest.mark.gpu
def test_bench_line_has_the_contract_keys():
    d = _line(["--steps", "3", "--warmup", "3", "--cpu-sample", str(1 << 20)])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["value"] > 1 and d["e2e"]["value"] > 1 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    assert d["roofline_build"]["frac"] > 0 and d["roofline_step"]["frac"] > 0 and d["query_miss_set"]["gkmers_s"] > 0
