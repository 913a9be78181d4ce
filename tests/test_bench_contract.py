"""bench.py's reference arm (the reference's own CPU path, oracle/_ref) on a tiny sample: the JSON line the driver
parses.  No GPU.  (The GPU arm needs a B200; its line is checked by the driver at round end.)"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "ecg_dump_ref")


def _run(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-n", "12"] + extra,
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref is not built (needs /root/reference)")
def test_reference_arm_line():
    lines = _run(["--gpus", "1", "--steps", "4", "--warmup", "3"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ecg_t8_bjacobi_iterations_per_s" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] == pytest.approx(1000.0 / d["value"])
    # --ref-n 12 is NOT the headline configuration, and the line says so: config names the grid that ran (with the default
    # --ref-n the reference runs the same 128^3 as the GPU arm; that takes minutes and is left to the driver)
    assert d["config"]["workload"].startswith("synthetic 3D Poisson 7-point 12^3 (1728 rows), ECG t=8 + block Jacobi")
    assert d["config"]["n"] == 12 and d["cpu_baseline"]["sample_iterations"] == 7
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "12^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref is not built (needs /root/reference)")
def test_reference_arm_other_ranks_are_silent():
    assert _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_gpu_arm_fails_loudly_without_a_device():
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in out.stderr and out.stdout.strip() == ""
