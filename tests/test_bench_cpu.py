"""bench.py contract pieces that run without a GPU: the reference arm (the CPU oracle port timed on the host cores) prints exactly
one JSON line with the driver's keys; the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "yolov5n", "--size", "128",
                        "--ref-frames", "1", "--rois", "2", "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["vs_baseline"] is None


def test_cuda_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_rank_core_binding_partitions_the_allowed_cores():
    """bench.pin_to_gpu_local_cores: ranks that share a core set split it into disjoint, equally sized shares (NVML is absent here,
    so the process' own affinity mask is the set); the binding is undone afterwards."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    before = os.sched_getaffinity(0)
    try:
        if len(before) < 2:
            import pytest
            pytest.skip("one core")
        world = 2
        shares = []
        for r in range(world):
            os.sched_setaffinity(0, before)
            shares.append(set(bench.pin_to_gpu_local_cores(0, r, world)))
            assert os.sched_getaffinity(0) == shares[-1]
        assert shares[0] and shares[1] and not (shares[0] & shares[1])
        assert len(shares[0]) == len(shares[1]) == len(before) // world
        assert (shares[0] | shares[1]) <= set(before)
    finally:
        os.sched_setaffinity(0, before)
