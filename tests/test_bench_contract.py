"""bench.py's contract on a machine without a GPU: the reference arm runs (the reference's CPU path on a bounded sample) and prints
ONE JSON line with the contract's keys; the product arm refuses to run without CUDA (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "slices/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_product_arm_refuses_to_run_without_cuda():
    p = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]          # no bench line from a CPU fallback
