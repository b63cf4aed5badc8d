"""bench.py's contract, the parts that run without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with the
keys the driver reads, on the CPU restatement only; the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_line_has_the_contract_keys():
    out = _run("--impl", "reference", "--config", "C1", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "aligned CCS reads/sec pileup+call+phase" and d["unit"] == "reads/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert abs(d["value"] - d["config"]["total_reads"] / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    assert d["config"]["workload"].startswith("C1") and d["config"]["total_reads"] == 5000 and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = _run("--config", "C1", "--steps", "1", "--warmup", "0")
    assert out.returncode != 0 and "no CPU fallback" in out.stderr and not out.stdout.strip()
