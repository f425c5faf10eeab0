"""bench.py's contract on the CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints ONE
JSON line with the keys the driver reads, and the GPU arm fails loudly -- no CPU fallback -- when there is no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-rows", "200000")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "filtered_scan_rows_per_s" and d["unit"] == "rows/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return                                            # on a GPU box the arm runs; covered by the driver
    p = _run("--steps", "1", "--warmup", "1", "--rows", "100000", "--no-cpu-baseline", env={"CUDA_VISIBLE_DEVICES": ""})
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]      # no number is ever printed from a CPU path
