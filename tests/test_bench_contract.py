"""bench.py's reference arm runs without a GPU: check its JSON line against the contract (one line, the device arm's
metric / unit / config keys, `impl`, `cpu_baseline`, `e2e`) on a tiny grid, and that the device arm refuses to run
without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--size", "32", "--cpu-sub", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "parsdmm_iterations_per_second" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["grid"] == [32, 32, 32] and d["config"]["ran_on_grid"] == [32, 32, 16]
    assert d["config"]["extrapolated"] is True and d["config"]["extrapolation_factor"] == 2.0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sub-volume" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    r = _run("--impl", "reference", "--gpus", "2", "--size", "32", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_device_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--size", "32", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "CUDA device" in (r.stderr + r.stdout)
