"""CPU: the reference arm of bench.py (the only bench leg that runs without a GPU) prints the contract's JSON
line; the GPU arm must refuse to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == d["unit"] == "walker-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] >= 3 and d["value"] > 0
    cb = d["cpu_baseline"]
    # the unmodified Python reference when baseline/_ref travelled with the snapshot (oracle/fetch_ref.py), else the
    # C port; with the reference, the port's number rides along
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "ndsimulator"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["value"] == d["value"]
    if have_ref:
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))          # not the single thread OMP_NUM_THREADS=1 asks for
    assert d["e2e"] == {"value": d["value"], "unit": "walker-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_port_reference_arm_on_request():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--port-reference",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 1e6


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
