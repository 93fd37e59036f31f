"""bench.py's command line contract on CPU: the reference arm runs visor's own rasterizer (oracle/_ref) and
prints ONE JSON line with the keys the driver reads; without a GPU the product arm must fail loudly, not
fall back."""
import json
import os
import subprocess
import sys

import pytest

from harness import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, env=e)


@pytest.mark.skipif(not abi.available("vref"), reason="oracle/_ref/libvisor_ref.so not built")
def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--workload", "c2", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "triangle throughput" and d["unit"] == "Mtri/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2")
    # same config keys as the product arm prints (bench.py: workload_config), --steps/--warmup honoured as given
    assert set(d["config"]) == {"workload", "triangles", "resolution", "l2", "parallelism", "raster_path"}
    assert d["steps"] == 2 and d["warmup"] == 1
    # the shader stage runs as native code (the reference JITs it); the interpreted figure is reported next to it
    assert "vkQueueSubmit" in d["reference_arm"]["timing"] and d["reference_arm"]["shader_stage"].startswith("native")
    assert d["reference_arm"]["modes"]["serial"]["frames"] == 2
    assert d["reference_arm"]["serial_with_interpreted_shaders"]["mtri_s"] > 0


@pytest.mark.skipif(not abi.available("vref"), reason="oracle/_ref/libvisor_ref.so not built")
def test_reference_arm_other_ranks_stay_silent():
    """under torchrun only rank 0 runs the reference; the other ranks exit 0 without output"""
    r = _run("--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present")
    r = _run("--workload", "c1", "--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
