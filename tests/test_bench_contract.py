"""CPU: the reference arm of ``bench.py`` prints one JSON line with the driver's contract (impl, metric, unit, value,
ms_per_step, cpu_baseline describing the run itself, e2e with zero copy bytes) -- exercised on the rendering workload,
whose CPU leg is cv2 itself (`kind: reference`) and takes a few seconds.  The GPU arm's line is checked by the
driver's own run; here only the fields both arms share are pinned, plus the failure mode of the GPU arm on a box
without a GPU (it must raise, never fall back to the CPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_line_render():
    r = _run("--impl", "reference", "--workload", "render", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "rendered views/s" and d["unit"] == "views/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u8"
    assert d["config"]["device"] == "cpu" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the GPU arm runs (covered by the driver's bench)")
    r = _run("--workload", "render", "--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0                           # no silent CPU fallback
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]


def test_reference_arm_under_torchrun_prints_one_line():
    """N > 1: the driver launches the reference arm under torchrun like the GPU arm; rank 0 alone measures and prints,
    the other rank exits 0 without work."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--gpus", "2", "--impl", "reference", "--workload", "render", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
