"""bench.py contract, CPU side: the reference arm (the oracle port timed on the host cores) prints ONE JSON line with the
keys the driver reads; without a CUDA device the GPU arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_line():
    # torchrun exports OMP_NUM_THREADS=1 to every rank (the N >= 2 reference arm ran single-threaded): the arm
    # must take every usable core anyway
    res = run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "images_per_sec_rpn_roialign_maskpaste" and d["unit"] == "images/s" and d["value"] > 0
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] == d["value"] and cb["sample"]
    rp = d["cpu_baseline_reference_python"]          # the unmodified reference functions from baseline/_ref, when staged
    assert ("unavailable" in rp) or (rp["kind"] == "reference" and rp["value"] > 0 and rp["cores"] >= 1 and rp["value"] < d["value"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
              env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.strip().startswith("{")]


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA present")
    res = run(["--steps", "1", "--warmup", "0"])
    assert res.returncode != 0
    assert not [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
