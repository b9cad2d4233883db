"""bench.py's contract as far as it can be checked without a GPU: the reference arm (the oracle port on the host cores)
prints ONE JSON line with BASELINE.json's metric and the keys the driver reads, non-zero ranks of a torchrun launch print
nothing, and the product arm refuses to run without a CUDA device instead of measuring some CPU path."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=300, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    line = json.loads(lines[0])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference"
    assert line["metric"] == baseline["metric"] and line["unit"] == "trajectory-steps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 1e6 and line["ms_per_step"] > 0
    assert "2000x56" in line["config"]["workload"] and line["config"]["batch_size"] == 2000 and line["config"]["time_steps"] == 56
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == 1 and cpu["value"] == line["value"] and cpu["sample"]
    # the steps are bounded samples of the workload: value = units of all samples / their total time
    units = 2000 * 56 * line["config"]["iteration_count"] * line["cycles_per_step"]
    assert abs(line["value"] - units / (line["ms_per_step"] * 1e-3)) / line["value"] < 1e-6


def test_reference_arm_is_silent_on_other_ranks():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"], RANK=1, LOCAL_RANK=1, WORLD_SIZE=2)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    out = _run(["--steps", "1", "--warmup", "1"])
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr
