"""bench.py's reference arm (CPU-runnable) prints ONE JSON line with the keys the driver's contract names; the arms and
flags the driver passes parse. The GPU arm shares the line builder's key set (checked on the GPU box by the bench run)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], capture_output=True, text=True, cwd=ROOT,
                       env=e, timeout=600)
    return r


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--batch", "1", "--cpu-sample-hw", "64x96")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == dict(value=d["value"], unit="img/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env=dict(RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
