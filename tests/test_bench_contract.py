"""CPU tier: bench.py's command-line contract that can be checked without a GPU.

* `--impl reference` (the CPU arm: the restated IPM of oracle/, the one place besides tests/smoke where oracle/ may run)
  prints ONE JSON line on rank 0 with the same metric/unit/config as the GPU arm; other ranks exit 0 without work.
* the GPU arm has no CPU fallback: without a CUDA device it exits non-zero and prints no JSON line.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **env):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update({k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=240)


def test_reference_arm_prints_one_contract_line(oracle):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", "16"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "converged SE(3) MPC solves/sec" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("configs[1]")


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], RANK=1, LOCAL_RANK=1, WORLD_SIZE=2)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "{" not in r.stdout
