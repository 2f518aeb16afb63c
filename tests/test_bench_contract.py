"""bench.py contract on a CPU-only box: the product arm refuses to run without CUDA (no CPU
fallback), the reference arm prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=600)


def test_product_arm_fails_loudly_without_cuda():
    p = _run("--config", "c1", "--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "CUDA" in p.stderr and p.stdout.strip() == ""


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lgconv_gedges_per_s" and d["unit"] == "GEdges/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
