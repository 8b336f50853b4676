"""bench.py --impl reference (the CPU arm) prints ONE JSON line with the contract's keys. Runs the CPU
path oracle on a bounded crop, no GPU needed. The GPU arm's line is checked on the GPU box (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline")


def _line(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _line(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference"
    for k in COMMON:
        assert k in d, k
    assert d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_loads_the_product_library():
    """The arm's scene comes from oracle/oracle_scenes.c; the timed call is the oracle's. The product .so must not
    even be mapped into the process (the driver records which .so files each arm loaded)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libg19oracle.so' in maps\n"
            "assert 'lib2019global_b200.so' not in maps, 'product library loaded by the reference arm'\n"
            "print('CLEAN')\n") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "CLEAN" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]


def test_reference_arm_other_ranks_print_nothing():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _line(["--steps", "2", "--warmup", "3", "--no-cpu", "--no-others"])
    for k in COMMON + ("gpu_launches", "clocks", "roofline", "parity", "other_configs"):
        assert k in d, k
    # the timed path's own correctness, checked inside bench.py after the timed region
    par = d["parity"]
    assert par["ok"] is True and par["windows_relrmse_max"] <= par["tolerance"] == 1e-2 and len(par["windows"]) >= 4
    assert d["metric"] == "Msamples/s" and d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["value"] > 100
    assert d["e2e"]["d2h_bytes_per_step"] == 1920 * 1080 * 3 and d["e2e"]["value"] > 100
    ro = d["roofline"]
    assert ro["bound"] in ("hbm", "tensor") and 0 < ro["frac"] < 1.5 and ro["unit"] == "GB/s" and ro["peak"] > 1000
